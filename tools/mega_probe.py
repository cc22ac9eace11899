"""In-kernel phase timing of the persistent solve kernel (csrc/admm_mega.cuh); developer tool, needs a B200.
usage: python tools/mega_probe.py [trials]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("JSTSP_DBG_KERNEL", "7")
os.environ.setdefault("JSTSP_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "debug", "libjstsp_b200.so"))      # the instrumented build (make debug-lib)
from jstsp19_b200 import _lib, synth  # noqa: E402
from jstsp19_b200.engine import AdmmEngine  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
imax = 16
s = synth.METRIC
dev = torch.device("cuda", 0)
data = synth.make_batch(s, nb, torch.zeros(nb, dtype=torch.float64), seed=1, device=dev)
eng = AdmmEngine(0, "f32")
ncta = min(nb, 148)
buf = torch.zeros(ncta * 16 * 8, dtype=torch.int64, device=dev)


def solve():
    eng.proposed_algorithm_psi(data["subY"], data["Omega"], data["A"], data["Dt"], data["Psi"], imax, data["tau_Y"], data["tau_Z"], data["rho"])


solve()
torch.cuda.synchronize()
_lib.lib.jstsp_debug_buffer(eng.h.ptr, C.c_void_p(buf.data_ptr()))
solve()
torch.cuda.synchronize()
_lib.lib.jstsp_debug_buffer(eng.h.ptr, None)
t = buf.cpu().numpy().reshape(ncta, 16, 8).astype(np.float64)
if os.environ["JSTSP_DBG_KERNEL"] == "8":
    nm = ["wait X,V1", "Z, W Z", "wait+read pass 1", "wait V2,subY, update, stores", "wait XV,G", "K operand (wait kop_empty)", "Gram"]
    print(f"variant {eng.h.last_variant}; chunk 3 of phase F, cycles per stage, median over {ncta} CTAs")
    for it in (0, 1, 2, 5, 10, 15):
        d = np.diff(t[:, it, :8], axis=1)
        print(f"it {it:2d}: " + " | ".join(f"{n} {np.median(d[:, i]):7.0f}" for i, n in enumerate(nm)) + f" | total {np.median(t[:, it, 7] - t[:, it, 0]):8.0f}")
    sys.exit(0)
if os.environ["JSTSP_DBG_KERNEL"] in ("11", "12"):
    nm = {"11": ["T1' from TMEM", "A', A A'", "operand image", "FFT", "Res store, |Res|^2"], "12": ["alpha, V, S", "A S", "inverse FFT", "operand image -> global"]}[os.environ["JSTSP_DBG_KERNEL"]]
    print(f"variant {eng.h.last_variant}; tap 1 of phase {'R' if os.environ['JSTSP_DBG_KERNEL'] == '11' else 'S'}, cycles per stage, median over {ncta} CTAs")
    for it in (0, 1, 2, 5, 10):
        d = np.diff(t[:, it, :len(nm) + 1], axis=1)
        print(f"it {it:2d}: " + " | ".join(f"{n} {np.median(d[:, i]):7.0f}" for i, n in enumerate(nm)) + f" | total {np.median(t[:, it, len(nm)] - t[:, it, 0]):8.0f}")
    sys.exit(0)
if os.environ["JSTSP_DBG_KERNEL"] == "15":
    w = t.reshape(ncta, 128)[:, :24].reshape(ncta, 8, 3)
    print(f"variant {eng.h.last_variant}; cycles thread 0 waits for state tiles in phase F of iteration 5, median over {ncta} CTAs")
    for c in range(8):
        print(f"chunk {c}: X,V1 {np.median(w[:, c, 0]):7.0f} | V2,subY {np.median(w[:, c, 1]):7.0f} | XV,G {np.median(w[:, c, 2]):7.0f}")
    print(f"sum per item: {np.median(w.sum(axis=(1, 2))):.0f}")
    sys.exit(0)
if os.environ["JSTSP_DBG_KERNEL"] == "13":
    print(f"variant {eng.h.last_variant}; MMA warp in phase G, cycles, median over {ncta} CTAs")
    for it in (0, 1, 2, 5, 10):
        print(f"it {it:2d}: wait image {np.median(t[:, it, 0]):8.0f} | wait pilot tiles {np.median(t[:, it, 1]):8.0f} | wait accumulators {np.median(t[:, it, 2]):8.0f} | issue {np.median(t[:, it, 3]):8.0f} | phase total {np.median(t[:, it, 4]):8.0f}")
    sys.exit(0)
if os.environ["JSTSP_DBG_KERNEL"] == "9":
    print(f"variant {eng.h.last_variant}; Jacobi warps, cycles, median over {ncta} CTAs")
    for it in (0, 1, 2, 5, 10, 14):
        d = np.diff(t[:, it, :4], axis=1)
        print(f"it {it:2d}: similarity {np.median(d[:, 0]):8.0f} | sweeps {np.median(d[:, 1]):8.0f} | weights {np.median(d[:, 2]):8.0f} | sweep count histogram {np.bincount(t[:, it, 4].astype(int))}")
    sys.exit(0)
names = ["F (8 chunks)", "wait T1' (last pass 2)", "R (Res, image of G)", "G (8 chunks)", "S (alpha, V, S, image)"]
print(f"variant {eng.h.last_variant}; cycles per phase, median over {ncta} CTAs (first trial of each)")
for it in (0, 1, 2, 5, 10, 15):
    d = np.diff(t[:, it, :6], axis=1)
    line = " | ".join(f"{n} {np.median(d[:, i]):8.0f}" for i, n in enumerate(names))
    tot = np.median(t[:, it, 5] - t[:, it, 0])
    jac = np.median(t[:, it, 7] - t[:, it, 6]) if it + 1 < imax else 0
    nxt = np.median(t[:, it + 1, 0] - t[:, it, 0]) if it + 1 < imax else 0
    print(f"it {it:2d}: {line} | total {tot:8.0f} | Jacobi {jac:8.0f} | start of this F to start of the trial's next F {nxt:8.0f} ({min(2, nb // ncta)} trials interleaved)")
