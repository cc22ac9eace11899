"""Developer check of the persistent solve kernel (csrc/admm_mega.cuh) on a B200: one persistent kernel per pass against the
four-kernel form of the same structured path and against the fp64 oracle, at a few iteration counts.
usage: python tools/mega_check.py [imax ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jstsp19_b200 as jb  # noqa: E402
from jstsp19_b200._lib import default_handle  # noqa: E402
from oracle import estimators as est  # noqa: E402
from oracle import fixtures as fx  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    imaxs = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 10, 100]
    trials = [fx.make_trial(fx.METRIC, snr, 100 + k) for k, snr in enumerate([-15.0, 0.0, 15.0])]
    st = lambda k: np.stack([t[k] for t in trials])
    args = (st("subY"), st("Omega"), st("A"), trials[0]["Dt"], st("Psi_bar"))
    par = ([t["tau_Y"] for t in trials], [t["tau_Z"] for t in trials], [t["rho"] for t in trials], "approximate")
    for imax in imaxs:
        out = {}
        for mega in ("0", "1"):
            os.environ["JSTSP_MEGA"] = mega
            t0 = time.time()
            S, Y = jb.proposed_algorithm_psi(*args, imax, *par, precision="f32", nargout=2)
            out[mega] = (S, Y, default_handle().last_path, default_handle().last_variant, time.time() - t0)
        S0 = []
        for t in trials:
            s0, y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
            S0.append((s0, y0))
        for k in range(len(trials)):
            e4 = rel(out["0"][0][k], S0[k][0]); em = rel(out["1"][0][k], S0[k][0]); ey = rel(out["1"][1][k], S0[k][1]); d = rel(out["1"][0][k], out["0"][0][k])
            print(f"imax {imax:3d} trial {k}: four-kernel vs oracle {e4:.2e} | mega vs oracle S {em:.2e} Y {ey:.2e} | mega vs four-kernel {d:.2e}"
                  f" | path/variant {out['0'][2]}/{out['0'][3]} {out['1'][2]}/{out['1'][3]}", flush=True)


if __name__ == "__main__":
    main()
