"""Rate of jstsp_vamp at BASELINE config 3's operator (256 x 1024 per trial, 100 iterations, device-resident).  Developer tool, needs a B200."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
mm, nn, nbv = 256, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator(device=dev); g.manual_seed(3)
h = _lib.Handle(0); h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
L = _lib.lib
p = lambda t: C.c_void_p(t.data_ptr())
crandn = lambda *sh: (torch.randn(*sh, generator=g, device=dev) + 1j * torch.randn(*sh, generator=g, device=dev)).to(torch.complex64)
Av = (crandn(nbv, nn, mm) / mm ** 0.5).contiguous()
Uh, sh = [], []
for k in range(nbv):
    U_, s_, _ = np.linalg.svd(Av[k].cpu().numpy().T.astype(np.complex128), full_matrices=True)
    Uh.append(np.ascontiguousarray(U_.T)); sh.append(s_ ** 2)
Ud = torch.tensor(np.stack(Uh), dtype=torch.complex64, device=dev).contiguous(); dd = torch.tensor(np.stack(sh), dtype=torch.float32, device=dev).contiguous()
yv = crandn(nbv, mm); xv = torch.empty(nbv, nn, dtype=torch.complex64, device=dev)
sg, Ln = torch.full((nbv,), 1.0, dtype=torch.float64, device=dev), torch.full((nbv,), 50.0, dtype=torch.float64, device=dev)
fn = lambda: h.check(L.jstsp_vamp(h.ptr, _lib.F32, _lib.DEVICE, mm, nn, nbv, 100, 0.85, p(yv), mm, p(Av), mm * nn, p(sg), p(Ln), p(Ud), mm * mm, p(dd), mm, p(xv), nn))
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"vamp {mm} x {nn}, {nbv} trials, 100 iterations: {ms:.2f} ms per call, {nbv / ms * 1e3:.0f} estimates/s, finite {bool(torch.isfinite(torch.view_as_real(xv)).all())}")
