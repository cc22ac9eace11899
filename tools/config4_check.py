"""BASELINE config 4 at its real size (Nt=256, Nr=64, K=128 frames, L=8: subY 64 x 32768, dictionary 2048 x 32768 never sent by the host) on a B200:
the pilots entry against the fp64 structured oracle, and the time per call.
usage: python tools/config4_check.py [imax] [trials] [precision] [Nt Nr L T]    (other shapes of the large-array route)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jstsp19_b200 as jb  # noqa: E402
from jstsp19_b200._lib import default_handle  # noqa: E402
from oracle import estimators as est  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

CONFIG4 = fx.Shape(Nt=256, Nr=64, L=8, Mr=4, T=128)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    imax = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ntr = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    prec = sys.argv[3] if len(sys.argv) > 3 else "f32"
    global CONFIG4
    if len(sys.argv) > 7:
        CONFIG4 = fx.Shape(Nt=int(sys.argv[4]), Nr=int(sys.argv[5]), L=int(sys.argv[6]), Mr=4, T=int(sys.argv[7]))
    trials = []
    for k in range(ntr):
        t0 = time.time()
        trials.append(fx.make_trial(CONFIG4, 5.0, 4001 + k, rho_rule="sigma1"))      # plot_errorVSdelays.m:22,127: 5 dB, rho from the largest eigenvalue
        print(f"trial {k} drawn in {time.time() - t0:.1f} s", flush=True)
    st = lambda key: np.stack([t[key] for t in trials])
    par = ([t["tau_Y"] for t in trials], [t["tau_Z"] for t in trials], [t["rho"] for t in trials], "approximate")
    for rep in range(2):
        t0 = time.time()
        S, Y = jb.proposed_algorithm_pilots(st("subY"), st("Omega"), st("A"), trials[0]["Dt"], st("pilots"), CONFIG4.L, imax, *par, precision=prec, nargout=2)
        dt = time.time() - t0
        print(f"GPU call {rep}: {dt:.3f} s for {ntr} trials x {imax} iterations, path/variant {default_handle().last_path}/{default_handle().last_variant}", flush=True)
    if os.environ.get("CONFIG4_PROFILE"):
        hd = default_handle(); hd.profile(1)
        jb.proposed_algorithm_pilots(st("subY"), st("Omega"), st("A"), trials[0]["Dt"], st("pilots"), CONFIG4.L, imax, *par, precision=prec, nargout=2)
        pr = hd.profile_read(); hd.profile(0)
        tot = sum(v[0] for v in pr.values())
        for k, v in pr.items():
            if v[1]:
                print(f"  {k:12s} {v[0]:9.3f} ms over {v[1]:6d} launches ({100 * v[0] / tot:5.1f} %)")
        print(f"  device total {tot:.2f} ms -> {tot / ntr / max(imax, 1):.3f} ms per trial-iteration")
    if os.environ.get("CONFIG4_NO_ORACLE"):
        return
    for k, t in enumerate(trials):
        t0 = time.time()
        s0, y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
        print(f"trial {k}: oracle {time.time() - t0:.1f} s | S rel {rel(S[k], s0):.2e} Y rel {rel(Y[k], y0):.2e} | NMSE gpu {est.nmse(S[k], t['Zbar']):.4e} oracle {est.nmse(s0, t['Zbar']):.4e}", flush=True)


if __name__ == "__main__":
    main()
