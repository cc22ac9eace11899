"""In-kernel phase timing of the ADMM fast-path kernels (developer tool, needs a B200).
usage: JSTSP_DBG_KERNEL={0|2|3} python tools/phase_probe.py   (0 = k_xupd_t1_fast, 2 = k_xs_fast, 3 = k_fused_tc; 3..6 need JSTSP_TC=1)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import _lib, synth  # noqa: E402
from jstsp19_b200.engine import AdmmEngine  # noqa: E402

nb = 592
s = synth.METRIC
dev = torch.device("cuda", 0)
data = synth.make_batch(s, nb, torch.zeros(nb, dtype=torch.float64), seed=1, device=dev)
eng = AdmmEngine(0, "f32")
kid = int(os.environ.get("JSTSP_DBG_KERNEL", "0"))
ncta = nb * 8
buf = torch.zeros(ncta * 8, dtype=torch.int64, device=dev)
psi = os.environ.get("PSI", "0") == "1"
def solve():
    if psi:
        eng.proposed_algorithm_psi(data["subY"], data["Omega"], data["A"], data["Dt"], data["Psi"], 3, data["tau_Y"], data["tau_Z"], data["rho"])
    else:
        eng.proposed_algorithm(data["subY"], data["Omega"], data["A"], data["B"], 3, data["tau_Y"], data["tau_Z"], data["rho"])
for it in range(2):
    solve()
torch.cuda.synchronize()
_lib.lib.jstsp_debug_buffer(eng.h.ptr, C.c_void_p(buf.data_ptr()))
solve()
torch.cuda.synchronize()
_lib.lib.jstsp_debug_buffer(eng.h.ptr, None)
t = buf.cpu().numpy().reshape(ncta, 8)
t = t[t[:, 0] > 0]
if kid in (4, 5, 6) and not psi:
    cols = {4: ["MMA thread total", "wait full", "wait lo_ready", "wait kop_ready", "wait d2_empty", "issue (64 x 4 MMAs)", "commits"], 5: ["TMA thread total", "wait empty"],
            6: ["wait hi_done (32 stages)", "rewrite body", "fence+arrive", "epilogues 0..2"]}[kid]
    print(f"kernel 3 role stats (dbg {kid}), cycles per CTA: median / mean / p90")
    for i, n in enumerate(cols):
        v = t[:, i + 1].astype(np.float64)
        print(f"  {n:26s} {np.median(v):10.0f} {v.mean():10.0f} {np.percentile(v, 90):10.0f}")
    sys.exit(0)
if psi and kid == 6:
    t = t[: nb * 4]
    pn = ["load A/tw + T1' partial sums", "A'., AA'., scale", "FFT", "Res store + rr", "image build + store"]
    d = np.diff(t[:, :6], axis=1).astype(np.float64)
    print(f"k_psi_res: {len(t)} CTAs, cycles per phase (median / mean / p90)")
    for i, n in enumerate(pn):
        print(f"  {n:30s} {np.median(d[:, i]):9.0f} {d[:, i].mean():9.0f} {np.percentile(d[:, i], 90):9.0f}")
    tot = (t[:, 5] - t[:, 0]).astype(np.float64)
    print(f"  {'total':30s} {np.median(tot):9.0f} {tot.mean():9.0f} {np.percentile(tot, 90):9.0f}")
    sys.exit(0)
if psi and kid == 5:
    t = t[:nb]
    jac = (t[:, 1] - t[:, 0]).astype(np.float64); tot = (t[:, 3] - t[:, 0]).astype(np.float64)
    print(f"k_svt_weights (last launch): Jacobi cycles median {np.median(jac):.0f} p90 {np.percentile(jac, 90):.0f}; total median {np.median(tot):.0f}; sweeps histogram {np.bincount(t[:, 2].astype(int))}")
    sys.exit(0)
names = {3: ["X,V1 tiles, Z, W Z, pass-1 wait", "V2,subY,XV tiles, update, K operand", "pass-2 epilogues", "V2 store + gram"] if psi else ["pass-1 rewrites", "wait D1 + tmem ld", "element-wise", "pass-2 rewrites+epi", "last epilogue", "gram"], 0: ["init+ring zero", "Z staging", "W Z + element-wise", "T1 main loop", "gram"], 2: ["init+ring zero", "AS staging", "main loop", "epilogue"]}[kid]
d = np.diff(t[:, : len(names) + 1], axis=1).astype(np.float64)
print(f"kernel {kid}: {len(t)} CTAs, clock cycles per phase (median / mean / p90)")
for i, n in enumerate(names):
    print(f"  {n:22s} {np.median(d[:, i]):10.0f} {d[:, i].mean():10.0f} {np.percentile(d[:, i], 90):10.0f}")
tot = (t[:, len(names)] - t[:, 0]).astype(np.float64)
print(f"  {'total':22s} {np.median(tot):10.0f} {tot.mean():10.0f} {np.percentile(tot, 90):10.0f}")
sm = t[:, 7]
for sid in (0, 1):
    sel = t[sm == sid]
    order = np.argsort(sel[:, 0])
    base = sel[order[0], 0]
    print(f"SM {sid}: CTA (start, end) cycles:", [(int(r[0] - base), int(r[len(names)] - base)) for r in sel[order][:8]])
