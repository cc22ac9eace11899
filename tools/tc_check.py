"""Developer check (GPU): error of the fp32 paths against the fp64 oracle at the metric shape, with and without the
tensor-core path (JSTSP_DISABLE_TC=1 selects the FFMA kernels)."""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import estimators as est, fixtures as fx

def run():
    import jstsp19_b200 as jb
    trials = [fx.make_trial(fx.METRIC, snr, 100 + k) for k, snr in enumerate([-15.0, 0.0, 15.0])]
    st = lambda k: np.stack([t[k] for t in trials])
    S1, Y1 = jb.proposed_algorithm(st("subY"), st("Omega"), st("A"), st("B"), 100, [t["tau_Y"] for t in trials], [t["tau_Z"] for t in trials],
                                   [t["rho"] for t in trials], "approximate", precision="f32", nargout=2)
    for k, t in enumerate(trials):
        S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 100, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
        n0, n1 = est.nmse(S0, t["Zbar"]), est.nmse(S1[k].astype(np.complex128), t["Zbar"])
        rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
        print(f"  trial {k}: relS {rel(S1[k], S0):.3e} relY {rel(Y1[k], Y0):.3e} rel_nmse {abs(n1 - n0) / n0:.3e}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for env in ({"JSTSP_TC": "1"}, {"JSTSP_TC": "0"}):
            print("env", env, flush=True)
            subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env})
