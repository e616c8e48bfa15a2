function y = randsample(n, k)
% RANDSAMPLE shadow used ONLY while replaying a twoace parity bundle (tools/matlab/run_bundle.m):
% returns the next stored draw instead of consuming MATLAB's RNG, so that inferLowRankV4.m:37
% (inferLowRankV4_multi.m:48: three calls) splits the rows exactly as the GPU / oracle run did.
% Put this directory FIRST on the path for the replay and remove it afterwards.
  global TWOACE_DRAWS TWOACE_DRAW_POS
  if isempty(TWOACE_DRAWS)
    error('twoace:randsample', 'no stored draws: this shadow is only valid inside run_bundle.m');
  end
  TWOACE_DRAW_POS = TWOACE_DRAW_POS + 1;
  y = TWOACE_DRAWS{TWOACE_DRAW_POS};
  y = y(:);
  if numel(y) ~= k || any(y > n)
    error('twoace:randsample', 'stored draw %d does not match randsample(%d,%d)', TWOACE_DRAW_POS, n, k);
  end
end
