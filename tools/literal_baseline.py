"""CPU baseline of SURVEY.md 8(d): the reference's own formulation of proposed_algorithm.m - dense K1, K2 = kron(B.', A),
R = K2'K2 (proposed_algorithm.m:14-25) - restated in NumPy fp64 (oracle.estimators.proposed_algorithm_literal), timed on the
host cores next to the Kronecker-free restatement that bench.py uses as its `cpu_baseline`.  No MATLAB / Octave exists in the
image, so this is the closest thing to "the reference's CPU path" that can run; it is a reported baseline, not a target.

    python tools/literal_baseline.py [--metric-trials 1] [--config0-trials 10] > profiles/r01_cpu_literal.json

Config 0 = plot_errorVSsnr.m defaults (32 x 140, BASELINE.json configs[0]: OMP + proposed ADMM, 1 SNR point, 10 trials);
metric shape = Nt 64, Nr 16, K 16, L 4 (16 x 1024; the literal operators take 2 GiB + 1 GiB + 256 MiB per trial)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import estimators as est       # noqa: E402
from oracle import fixtures as fx          # noqa: E402


def run(shape, trials, literal, imax=100, with_omp=False):
    t_tot, nm = 0.0, []
    for k in range(trials):
        t = fx.make_trial(shape, 5.0, 900 + k)
        args = (t["subY"], t["Omega"], t["A"], t["B"], imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
        t0 = time.perf_counter()
        if literal:
            S, _, _ = est.proposed_algorithm_literal(*args)
        else:
            S, _, _ = est.proposed_algorithm_structured(*args, want_conv=False)
        if with_omp:                                   # the conventional-HBF OMP of plot_errorVSsnr.m:79-80 (OMP.m semantics, numOfnz atoms)
            c = fx.conventional_problem(t)
            est.omp_literal(c["Phi"], c["y"], 100)
        t_tot += time.perf_counter() - t0
        nm.append(est.nmse(S, t["Zbar"]))
    return dict(trials=trials, seconds=t_tot, per_s=trials / t_tot, mean_nmse=float(np.mean(nm)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--metric-trials", type=int, default=1)
    ap.add_argument("--config0-trials", type=int, default=10)
    a = ap.parse_args()
    out = dict(cores=os.cpu_count(), numpy=np.__version__, unit="estimates/s", imax=100,
               note="NumPy fp64 restatement of the MATLAB path (no MATLAB/Octave in the image); literal = dense K1/K2/R exactly as proposed_algorithm.m:14-25")
    out["config0_literal_admm_plus_omp"] = run(fx.CONFIG0, a.config0_trials, True, with_omp=True)
    out["config0_literal_admm"] = run(fx.CONFIG0, a.config0_trials, True)
    out["config0_structured_admm"] = run(fx.CONFIG0, a.config0_trials, False)
    out["metric_structured_admm"] = run(fx.METRIC, 12, False)
    if a.metric_trials > 0:
        out["metric_literal_admm"] = run(fx.METRIC, a.metric_trials, True)
    print(json.dumps(out))
