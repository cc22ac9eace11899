"""Rates of mc_svt / mc_admm at BASELINE config 3's longest frame (32 x 280, per-trial operands, 100 iterations, device-resident), with the
persistent one-launch kernel and with the two-kernels-per-iteration form (JSTSP_SVT_PERSIST_OFF=1).  Developer tool, needs a B200."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
nbs, Mr, Mt, IM = int(sys.argv[1]) if len(sys.argv) > 1 else 592, 32, 280, 100
g = torch.Generator(device=dev); g.manual_seed(3)
h = _lib.Handle(0); h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
L = _lib.lib
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
OH = (torch.randn(nbs, Mt, Mr, generator=g, device=dev) + 1j * torch.randn(nbs, Mt, Mr, generator=g, device=dev)).to(torch.complex64)
Om = (torch.rand(nbs, Mt, Mr, generator=g, device=dev) < 0.125).float().contiguous(); OH = (OH * Om).contiguous()
tau = torch.full((nbs,), 0.02, dtype=torch.float64, device=dev); rho = torch.full((nbs,), 0.1, dtype=torch.float64, device=dev)
res = {}
for off in ("1", ""):
    if off:
        os.environ["JSTSP_SVT_PERSIST_OFF"] = off
    else:
        os.environ.pop("JSTSP_SVT_PERSIST_OFF", None)
    for name in ("mc_svt", "mc_admm"):
        X = torch.empty_like(OH)
        if name == "mc_svt":
            fn = lambda: h.check(L.jstsp_mc_svt(h.ptr, _lib.F32, _lib.DEVICE, Mr, Mt, nbs, IM, p(OH), Mr * Mt, p(Om), Mr * Mt, p(tau), p(rho), p(X), Mr * Mt))
        else:
            fn = lambda: h.check(L.jstsp_mc_admm(h.ptr, _lib.F32, _lib.DEVICE, Mr, Mt, nbs, IM, None, 0, p(OH), Mr * Mt, p(Om), Mr * Mt, p(tau), p(rho), p(X), Mr * Mt, None, 0))
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        h.profile(1); fn(); pr = h.profile_read(); h.profile(0)
        print("   ", {k: (round(v[0], 2), v[1]) for k, v in pr.items() if v[1]})
        res[(name, off)] = X.clone()
        print(f"{name:8s} {'two kernels per iteration' if off else 'persistent               '}: {ms:8.2f} ms per call, {nbs / ms * 1e3:9.0f} estimates/s", flush=True)
for name in ("mc_svt", "mc_admm"):
    print(name, "bitwise equal between the two forms:", bool(torch.equal(res[(name, "1")], res[(name, "")])))
