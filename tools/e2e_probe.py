"""Developer probe (GPU): wall time of the HOST-buffer C-ABI call vs batch size and pass size (copy/compute overlap)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import _lib, synth
from jstsp19_b200.engine import AdmmEngine
s = synth.METRIC; N, M, G, P = s.Nr, s.M, s.Nr, s.P
dev = torch.device("cuda", 0)
nbmax = 1184
data = synth.make_batch(s, nbmax, torch.zeros(nbmax, dtype=torch.float64), seed=1, device=dev)
eng = AdmmEngine(0, "f32")
pin = lambda t: t.cpu().contiguous().pin_memory()
hsubY, hOm, hB, hA = pin(data["subY"]), pin(data["Omega"]), pin(data["B"]), pin(data["A"])
hty, hts, hrho = (pin(data[k]) for k in ("tau_Y", "tau_Z", "rho"))
hS = torch.empty(nbmax, P, G, dtype=torch.complex64).pin_memory()
vp = lambda t: C.c_void_p(t.data_ptr())
def call(ne, chunk):
    _lib.lib.jstsp_set_chunk(eng.h.ptr, chunk)
    d = _lib.AdmmDesc(); d.N, d.M, d.G, d.P, d.imax, d.type, d.batch = N, M, G, P, 100, _lib.APPROXIMATE, ne
    d.ld_subY, d.ld_omega, d.ld_A, d.ld_B, d.ld_S, d.ld_Y = N * M, N * M, 0, P * M, G * P, N * M
    for rep in range(3):
        t0 = time.perf_counter()
        rc = _lib.lib.jstsp_proposed_algorithm(eng.h.ptr, C.byref(d), _lib.F32, _lib.HOST, vp(hsubY), vp(hOm), vp(hA), vp(hB), vp(hty), vp(hts), vp(hrho), vp(hS), None, None)
        dt = time.perf_counter() - t0
    print(f"trials {ne:5d} chunk {chunk:4d}: {dt*1e3:8.1f} ms  -> {ne/dt:7.0f} est/s", flush=True)
for ne, chunk in [(296, 0), (592, 592), (592, 0), (592, 296), (1184, 1184), (1184, 0), (1184, 592)]:
    call(ne, chunk)
