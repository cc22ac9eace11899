"""Frozen fp64 oracle outputs for BASELINE config 4 at its real size (Nt=256, Nr=64, 128 frames, L=8; plot_errorVSdelays.m:22,45-49,127-128:
5 dB, rho from the largest eigenvalue), 100 iterations of proposed_algorithm('approximate') on two seeded trials.
The oracle needs ~3.5 minutes per trial on 8 cores, too slow for the GPU suite, so its outputs are committed:
tests/golden/config4_full.npz holds S (64 x 2048, exact fp64), the NMSE, |Y|_F and the first 64 columns of Y per trial;
tests/test_gpu_large.py regenerates the inputs from the same seeds (fixtures.make_trial) and compares the CUDA path with them.
usage: python tools/make_golden_config4.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import estimators as est  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

CONFIG4 = fx.Shape(Nt=256, Nr=64, L=8, Mr=4, T=128)
SEEDS = (4001, 4002)
SNR_DB = 5.0
IMAX = 100


def main():
    out = {"seeds": np.array(SEEDS), "snr_db": SNR_DB, "imax": IMAX}
    for k, seed in enumerate(SEEDS):
        t = fx.make_trial(CONFIG4, SNR_DB, seed, rho_rule="sigma1")
        t0 = time.time()
        S, Y, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], IMAX, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
        print(f"seed {seed}: {time.time() - t0:.0f} s, NMSE {est.nmse(S, t['Zbar']):.6e}", flush=True)
        out[f"S{k}"] = S
        out[f"nmse{k}"] = est.nmse(S, t["Zbar"])
        out[f"Ynorm{k}"] = np.linalg.norm(Y)
        out[f"Yhead{k}"] = Y[:, :64].copy()
        out[f"params{k}"] = np.array([t["tau_Y"], t["tau_Z"], t["rho"]])
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "config4_full.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
