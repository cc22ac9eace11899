"""Developer check (GPU): Psi-domain tensor-core path vs the fp64 oracle at the metric shape (3 trials), then speed of both entries."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import estimators as est, fixtures as fx

import jstsp19_b200 as jb
from jstsp19_b200._lib import default_handle
trials = [fx.make_trial(fx.METRIC, snr, 100 + k) for k, snr in enumerate([-15.0, 0.0, 15.0])]
st = lambda k: np.stack([t[k] for t in trials])
for imax in (1, 2, 3, 100):
    S1, Y1 = jb.proposed_algorithm_psi(st("subY"), st("Omega"), st("A"), trials[0]["Dt"], st("Psi_bar"), imax, [t["tau_Y"] for t in trials],
                                       [t["tau_Z"] for t in trials], [t["rho"] for t in trials], "approximate", precision="f32", nargout=2)
    print("imax", imax, "path", default_handle().last_path, flush=True)
    for k, t in enumerate(trials):
        S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
        rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
        n0, n1 = est.nmse(S0, t["Zbar"]), est.nmse(S1[k].astype(np.complex128), t["Zbar"])
        print(f"  trial {k}: relS {rel(S1[k], S0):.3e} relY {rel(Y1[k], Y0):.3e} rel_nmse {abs(n1 - n0) / n0:.3e}", flush=True)
# generic-rotation route (JSTSP_PSI_NOFFT=1) must agree with the FFT route
if os.environ.get("JSTSP_PSI_NOFFT") is None:
    import subprocess
    print("--- with JSTSP_PSI_NOFFT=1 ---", flush=True)
    subprocess.run([sys.executable, __file__], env={**os.environ, "JSTSP_PSI_NOFFT": "1"})
