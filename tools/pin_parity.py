#!/usr/bin/env python
"""Pin parity against the real reference (needs MATLAB on some machine).

  1.  python tools/pin_parity.py export bundle.mat [--trials 8]      # seeded config-0/1/3 style instances
  2.  (MATLAB)  run_bundle('bundle.mat', 'bundle_ref.mat', '<reference checkout>')     # tools/matlab/
  3.  python tools/pin_parity.py compare bundle.mat bundle_ref.mat [--impl gpu|oracle]

`compare` re-solves the bundle with the CUDA library (default) or the NumPy oracle and reports, per instance,
the relative error of the recovered CSI after global-phase alignment (Evaluation_H.m:81-82) against MATLAB's
result, the quality difference, and the summary the north star asks for (fraction within 1e-4, NMSE delta)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

_MAP = {"inferLowRankV4": "V4", "inferLowRankV4_multi": "V4_MULTI", "inferLowRank_Nuclear": "NUCLEAR",
        "inferLowRankV3": "V3", "inferLowRankV2": "V2", "inferLowRank": "V1"}


def do_export(args):
    import twoace_b200 as tw
    hz = tw.harness
    cb = hz.load_codebook()
    A, B, tr, solver = [], [], [], []
    for name, M, T in (("inferLowRankV4", 64, 1), ("inferLowRankV4_multi", 64, 3), ("inferLowRank_Nuclear", 128, 1),
                       ("inferLowRankV3", 64, 1)):
        for ins in hz.make_batch(args.trials, cb, M, 20.0):
            A.append(ins.A); B.append(ins.B); tr.append(ins.train_idx[:T]); solver.append(name)
    for ins in hz.make_batch(max(1, args.trials // 4), cb, 128, 20.0):
        A.append(ins.A); B.append((ins.B / 2.0) ** 2); tr.append(np.zeros((0, 0), np.int32)); solver.append("MyPhaseLift")
    hz.export_bundle(args.bundle, A, B, 16, 16, solver, tr)
    print(f"wrote {len(A)} instances to {args.bundle}")


def solve_bundle(S, impl):
    """Re-solve every instance of an imported bundle; returns (X list, quality array)."""
    nb = len(S["A"])
    X, q = [None] * nb, np.full(nb, np.nan)
    if impl == "gpu":
        import twoace_b200 as tw
        for b in range(nb):
            name = S["solver"][b]
            if name == "MyPhaseLift":
                X[b] = tw.MyPhaseLift(S["B"][b], S["A"][b])
            else:
                pv = S["params"][b]
                p = tw.Params.default(lam=pv[0], r=int(pv[1]), mu0=pv[2], rho=pv[3], cc_frac=pv[4], tol_rel=pv[5],
                                      tol_abs=pv[6], maxiter=int(pv[7]))
                res = tw.solve_batch(getattr(tw, _MAP[name]), [S["A"][b]], [S["B"][b]], int(S["tx"][b]),
                                     int(S["rx"][b]), [S["train_idx"][b]], p)
                X[b], q[b] = res.X[0], res.quality[0]
    else:
        from oracle import admm, phaselift
        fns = {"V4": admm.infer_low_rank_v4, "V4_MULTI": admm.infer_low_rank_v4_multi,
               "NUCLEAR": admm.infer_low_rank_nuclear, "V3": admm.infer_low_rank_v3, "V2": admm.infer_low_rank_v2,
               "V1": admm.infer_low_rank_v1}
        for b in range(nb):
            name = S["solver"][b]
            if name == "MyPhaseLift":
                X[b] = phaselift.my_phase_lift(S["B"][b], S["A"][b])
            else:
                pv = S["params"][b]
                p = admm.Params(lam=pv[0], r=int(pv[1]), mu0=pv[2], rho=pv[3], cc_frac=pv[4], tol_rel=pv[5],
                                tol_abs=pv[6], maxiter=int(pv[7]))
                tr = S["train_idx"][b]
                X[b], _, q[b] = fns[_MAP[name]](S["A"][b], S["B"][b], int(S["tx"][b]), int(S["rx"][b]), p,
                                                train_idx=tr if _MAP[name] == "V4_MULTI" else tr[0])
    return X, q


def compare(S, ref, X, q):
    import twoace_b200 as tw
    errs = np.array([tw.harness.aligned_rel_err(X[b], ref["X"][b]) for b in range(len(X))])
    return errs, np.abs(q - ref["quality"])


def do_compare(args):
    import twoace_b200 as tw
    S = tw.harness.import_bundle(args.bundle)
    ref = tw.harness.import_reference_results(args.results)
    X, q = solve_bundle(S, args.impl)
    errs, dq = compare(S, ref, X, q)
    for b in range(len(X)):
        print(f"{b:4d} {S['solver'][b]:22s} m={len(S['B'][b]):4d}  rel err {errs[b]:.3e}  |dquality| {dq[b]:.2e}")
    ok = errs <= 1e-4
    print(f"\n{ok.mean() * 100:.1f} % of {len(X)} instances within 1e-4 of MATLAB {ref['matlab_version']} "
          f"(bar: >= 95 %); median {np.median(errs):.2e}, max {errs.max():.2e}")
    return 0 if ok.mean() >= 0.95 else 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    e = sub.add_parser("export"); e.add_argument("bundle"); e.add_argument("--trials", type=int, default=8)
    c = sub.add_parser("compare"); c.add_argument("bundle"); c.add_argument("results")
    c.add_argument("--impl", default="gpu", choices=["gpu", "oracle"])
    a = ap.parse_args()
    sys.exit(do_export(a) if a.cmd == "export" else do_compare(a))
