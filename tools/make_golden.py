"""Regenerates tests/golden/*.npz from the oracle (the reference is MATLAB and cannot run in this
image, so these are ORACLE outputs frozen as regression fixtures, not reference outputs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import estimators as est  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
t = fx.make_trial(fx.TINY, 5.0, 777)
Imax = 20
S, Y, c = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], Imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
Sa, Ya, ca = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], Imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate",
                                               indx_S=t["indx_S"])
np.savez_compressed(os.path.join(out, "admm_tiny.npz"), subY=t["subY"], Omega=t["Omega"], A=t["A"], B=t["B"], Imax=Imax,
                    tau_Y=t["tau_Y"], tau_S=t["tau_Z"], rho=t["rho"], S=S, Y=Y, conv=c, indx_S=t["indx_S"], S_angles=Sa, Zbar=t["Zbar"])
print("wrote", os.path.join(out, "admm_tiny.npz"))
