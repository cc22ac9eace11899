"""Race triage for the persistent solve kernel: dump checksums of everything iteration 1 consumes (state tiles, pass-1 result, W, alpha)
for every trial, twice, and report which quantity differs between two runs in the trials whose result differs.  usage: python tools/mega_dump.py [ntrials]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["JSTSP_DBG_KERNEL"] = "14"
from jstsp19_b200 import _lib, synth  # noqa: E402
from jstsp19_b200.engine import AdmmEngine  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
dev = torch.device("cuda", 0)
data = synth.make_batch(synth.METRIC, nb, torch.zeros(nb, dtype=torch.float64), seed=1, device=dev)
eng = AdmmEngine(0, "f32")
names = ["X", "V1", "pass-1 Xs", "V2", "subY", "XV", "G"]
runs = []
for r in range(4):
    buf = torch.zeros(nb * 1024, dtype=torch.float32, device=dev)
    _lib.lib.jstsp_debug_buffer(eng.h.ptr, C.c_void_p(buf.data_ptr()))
    S = eng.proposed_algorithm_psi(data["subY"], data["Omega"], data["A"], data["Dt"], data["Psi"], 2, data["tau_Y"], data["tau_Z"], data["rho"])
    torch.cuda.synchronize()
    runs.append((S.clone().cpu().numpy(), buf.cpu().numpy().reshape(nb, 1024)))
_lib.lib.jstsp_debug_buffer(eng.h.ptr, None)
S0, D0 = runs[0]
for r in range(1, 4):
    S1, D1 = runs[r]
    e = np.linalg.norm((S1 - S0).reshape(nb, -1), axis=1) / np.linalg.norm(S0.reshape(nb, -1), axis=1)
    bad = np.nonzero(e > 1e-6)[0]
    print(f"run {r} vs run 0: {len(bad)} trials differ")
    for b in bad[:12]:
        out = []
        for k, n in enumerate(names):
            a, c = D0[b, 8 * k: 8 * k + 8], D1[b, 8 * k: 8 * k + 8]
            d = np.abs(a - c) / np.maximum(np.abs(a), 1e-30)
            if d.max() > 1e-4:
                out.append(f"{n} chunks {np.nonzero(d > 1e-4)[0].tolist()} ({d.max():.1e})")
        dw = np.linalg.norm(D0[b, 512:] - D1[b, 512:]) / max(np.linalg.norm(D0[b, 512:]), 1e-30)
        if dw > 1e-6:
            out.append(f"W ({dw:.1e})")
        if abs(D0[b, 100] - D1[b, 100]) > 1e-6 * abs(D0[b, 100]):
            out.append("alpha_prev")
        print(f"   trial {b} (slot {b // 148 % 2}), S differs by {e[b]:.1e}: " + ("; ".join(out) if out else "no dumped input differs"))
