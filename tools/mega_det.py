"""Determinism scan of the persistent solve kernel (developer tool): same inputs twice per iteration count, bitwise comparison, and the
difference to the four-kernel form.  usage: python tools/mega_det.py [ntrials] [imax ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import synth  # noqa: E402
from jstsp19_b200.engine import AdmmEngine  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
imaxs = [int(a) for a in sys.argv[2:]] or [1, 2, 3, 4, 5, 6, 8, 12, 20]
dev = torch.device("cuda", 0)
data = synth.make_batch(synth.METRIC, nb, torch.zeros(nb, dtype=torch.float64), seed=1, device=dev)
eng = AdmmEngine(0, "f32")


def solve(imax, mega):
    os.environ["JSTSP_MEGA"] = mega
    S = eng.proposed_algorithm_psi(data["subY"], data["Omega"], data["A"], data["Dt"], data["Psi"], imax, data["tau_Y"], data["tau_Z"], data["rho"])
    torch.cuda.synchronize()
    return S.clone()


for imax in imaxs:
    ref = solve(imax, "0")
    runs = [solve(imax, "1") for _ in range(3)]
    var = eng.h.last_variant
    same = [bool(torch.equal(runs[0], r)) for r in runs[1:]]
    err = [float((r - ref).norm() / ref.norm()) for r in runs]
    per = ((runs[0] - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).cpu().numpy()
    for k, r in enumerate(runs):
        e = ((r - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).cpu().numpy()
        bad = np.nonzero(e > 1e-5)[0]
        if len(bad):
            print(f"   run {k}: trials off by > 1e-5: {[(int(i), int(i) % 148, int(i) // 148, float('%.1e' % e[i])) for i in bad]}  (trial, CTA, position in the CTA's list, error)")
    print(f"imax {imax:3d} variant {var}: bitwise repeatable {same}; rel diff to four-kernel {['%.2e' % e for e in err]}; per trial (run 0) {np.array2string(per, precision=1)}", flush=True)
