function run_bundle(bundle_file, result_file, reference_root)
% RUN_BUNDLE  Replay a twoace parity bundle through the UNMODIFIED reference solvers and save the results.
%
%   run_bundle('bundle.mat', 'bundle_ref.mat', '/path/to/2ACE-mmWave-Channel-Estimation')
%
% bundle.mat is written by twoace_b200.harness.export_bundle (Python): per instance the sensing matrix A,
% the RSS amplitudes B, the antenna counts, the solver name and the randsample draws (1-based, drawn order).
% The result file is read back by tools/pin_parity.py, which compares it with the CUDA library / the NumPy
% oracle.  This is the procedure that turns "parity unpinned" (DESIGN.md section 2) into pinned parity: it
% needs a MATLAB licence, which the build environment of this repository does not have.
  S = load(bundle_file);
  here = fileparts(mfilename('fullpath'));
  addpath(fullfile(reference_root, 'main', 'src', 'my_recovery_algorithms', 'ADMM_v2'));
  addpath(fullfile(reference_root, 'main', 'src', 'my_recovery_algorithms'));
  addpath(genpath(fullfile(reference_root, 'main', '3rd_software_component', 'sparsepr')));
  addpath(here, '-begin');                       % randsample shadow first
  global TWOACE_DRAWS TWOACE_DRAW_POS
  nb = numel(S.A);
  X = cell(nb, 1); Y = cell(nb, 1); quality = nan(nb, 1);
  for b = 1:nb
    TWOACE_DRAWS = S.train_idx{b};                % cell of draws (one, or three for _multi)
    if ~iscell(TWOACE_DRAWS), TWOACE_DRAWS = {TWOACE_DRAWS}; end
    TWOACE_DRAW_POS = 0;
    solver = S.solver{b};
    p = S.params(b, :);                           % [lambda r mu0 rho cc_frac tol_rel tol_abs maxiter]
    switch solver
      case {'inferLowRankV4', 'inferLowRankV4_multi', 'inferLowRank_Nuclear', 'inferLowRankV3'}
        f = str2func(solver);
        [X{b}, Y{b}, quality(b)] = f(S.A{b}, S.B{b}(:), double(S.tx(b)), double(S.rx(b)), ...
                                     p(1), p(2), p(3), p(4), p(5), p(6), p(7), p(8));
      case {'inferLowRankV2', 'inferLowRank'}
        f = str2func(solver);
        [X{b}, Y{b}, quality(b)] = f(S.A{b}, S.B{b}(:), double(S.tx(b)), double(S.rx(b)), p(1), p(2), p(6), p(7), p(8));
      case 'MyPhaseLift'
        X{b} = MyPhaseLift(S.B{b}(:), S.A{b});     % B holds intensities for this solver
        Y{b} = [];
      otherwise
        error('twoace:bundle', 'unknown solver %s', solver);
    end
  end
  rmpath(here);
  TWOACE_DRAWS = {}; TWOACE_DRAW_POS = 0;
  matlab_version = version;
  save(result_file, 'X', 'Y', 'quality', 'matlab_version', '-v7');
end
