"""Device-resident throughput of the benchmark solvers at the shapes of BASELINE
configs 0 / 3, scored against the HBM roofline with SURVEY 8(d)'s algorithmic bytes, with the NumPy oracle timed on the
host cores on a bounded sample.  One JSON line per solver.  (The ADMM headline is bench.py; Kronecker OMP is
tools/omp_kron_bench.py.)    python tools/solver_bench.py [--batch 1184] [--no-cpu]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jstsp19_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1184)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
# under torchrun every rank runs its own batch (weak scaling over independent trials); times are the max over ranks
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(7 + rank)
h = _lib.Handle(local)
h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
L = _lib.lib
F32, DEV = _lib.F32, _lib.DEVICE
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6557.8))
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def crandn(*shape):
    return (torch.randn(*shape, generator=g, device=dev) + 1j * torch.randn(*shape, generator=g, device=dev)).to(torch.complex64)


LAST_KERNELS = {}


def timed(fn):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    L.jstsp_profile(h.ptr, 2)                       # one more call with per-kernel-class CUDA events (not part of the timing above)
    fn(); torch.cuda.synchronize()
    LAST_KERNELS.clear()
    slot = 0
    while True:
        t, n, name = C.c_double(), C.c_longlong(), C.c_char_p()
        if L.jstsp_profile_read(h.ptr, slot, C.byref(t), C.byref(n), C.byref(name)) != 0:
            break
        if n.value:
            LAST_KERNELS[name.value.decode()] = dict(ms_total=round(t.value, 4), launches=n.value)
        slot += 1
    L.jstsp_profile(h.ptr, 0)
    return ms


def cpu_rate(fn, n):
    if a.no_cpu or rank != 0 or world > 1:
        return None
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return n / (time.perf_counter() - t0)


def emit(name, shape, batch, ms, bytes_per_unit, cpu, note, flops_per_unit=None):
    rate = world * batch / ms * 1e3
    gbs = bytes_per_unit * rate / 1e9 / world                 # per-GPU bandwidth for the roofline
    if rank != 0:
        return
    line = dict(solver=name, shape=shape, n_gpus=world, batch=batch, ms_per_call=ms, units_per_s=rate, algorithmic_bytes_per_unit=bytes_per_unit,
                roofline=dict(bound="hbm", achieved=gbs, peak=HBM, unit="GB/s", frac=gbs / HBM), note=note, kernels=dict(LAST_KERNELS),
                cpu_baseline=dict(value=cpu, unit="units/s", kind="port", cores=os.cpu_count()) if cpu else None)
    if flops_per_unit:
        line["algorithmic_tflops"] = flops_per_unit * rate / 1e12
    print(json.dumps(line), flush=True)


from oracle import estimators as est  # noqa: E402  (CPU baseline leg only)
from oracle import vamp as ovamp  # noqa: E402

B = a.batch
one = lambda v: torch.full((B,), v, dtype=torch.float64, device=dev)

# ---- svt.m:1-15 at the metric shape and at config 3's longest frame ----
for (Mr, Mt) in ((16, 1024), (32, 280)):
    Y = crandn(B, Mt, Mr); X = torch.empty_like(Y); tau = one(0.5)
    ms = timed(lambda: h.check(L.jstsp_svt(h.ptr, F32, DEV, Mr, Mt, B, p(Y), Mr * Mt, p(tau), p(X), Mr * Mt)))
    y0 = Y[0].cpu().numpy().T
    emit("svt", [Mr, Mt], B, ms, 2 * 8 * Mr * Mt, cpu_rate(lambda: est.svt_structured(y0, 0.5), 20), "read Y, write X; Gram + Jacobi + apply in one kernel")

# ---- mc_svt.m / mc_admm.m, Imax = 100, 32 x 280 (plot_errorVSframelength.m T = 35) ----
Mr, Mt, IMAX = 32, 280, 100
OH = crandn(B, Mt, Mr); Om = (torch.rand(B, Mt, Mr, generator=g, device=dev) < 0.125).float().contiguous(); OH = (OH * Om).contiguous()
X = torch.empty_like(OH); tau, rho = one(0.02), one(0.1)
ms = timed(lambda: h.check(L.jstsp_mc_svt(h.ptr, F32, DEV, Mr, Mt, B, IMAX, p(OH), Mr * Mt, p(Om), Mr * Mt, p(tau), p(rho), p(X), Mr * Mt)))
oh0, om0 = OH[0].cpu().numpy().T, Om[0].cpu().numpy().T
emit("mc_svt", [Mr, Mt, IMAX], B, ms, IMAX * (3 * 8 * Mr * Mt + Mr * Mt // 8) + 8 * Mr * Mt, cpu_rate(lambda: est.mc_svt(oh0, om0, IMAX, 0.02, 0.1), 2),
     "SURVEY 8(d): Imax (read Y, read OH, write Y, mask bits) + X")
ms = timed(lambda: h.check(L.jstsp_mc_admm(h.ptr, F32, DEV, Mr, Mt, B, IMAX, None, 0, p(OH), Mr * Mt, p(Om), Mr * Mt, p(tau), p(rho), p(X), Mr * Mt, None, 0)))
emit("mc_admm", [Mr, Mt, IMAX], B, ms, IMAX * (5 * 8 * Mr * Mt + Mr * Mt // 8), cpu_rate(lambda: est.mc_admm_structured(oh0, oh0, om0, IMAX, 0.02, 0.1), 2),
     "SURVEY 8(d): Imax (Y, Z read + write, OH read, mask bits)")

# ---- sparse_admm.m, Mr = 32, Mt = 8 (config 3 antennas), Imax = 100 ----
Mr, Mt = 32, 8
Dr = (torch.exp(-2j * torch.pi * torch.outer(torch.arange(Mr, device=dev), torch.arange(Mr, device=dev)) / Mr) / Mr ** 0.5).to(torch.complex64).contiguous()
Dt = (torch.exp(-2j * torch.pi * torch.outer(torch.arange(Mt, device=dev), torch.arange(Mt, device=dev)) / Mt) / Mt ** 0.5).to(torch.complex64).contiguous()
OHs = crandn(B, Mt, Mr); S = torch.empty_like(OHs)
ms = timed(lambda: h.check(L.jstsp_sparse_admm(h.ptr, F32, DEV, Mr, Mt, B, IMAX, None, 0, p(OHs), Mr * Mt, p(Dr), 0, p(Dt), 0, p(S), Mr * Mt, None, 0)))
ohs0, dr0, dt0 = OHs[0].cpu().numpy().T, Dr.cpu().numpy().T, Dt.cpu().numpy().T
emit("sparse_admm", [Mr, Mt, IMAX], B, ms, IMAX * 4 * 8 * Mr * Mt + 8 * Mr * Mt, cpu_rate(lambda: est.sparse_admm_structured(ohs0, ohs0, dr0, dt0, IMAX), 3),
     "SURVEY 8(d): Imax x 4 x 8 Gr Gt + 8 Mr Mt; the state of a 32 x 8 problem lives on chip, so this is latency-, not HBM-bound")

# ---- OMP.m on the dense conventional system (plot_errorVSsnr.m:79-80: Phi 512 x 512, m = 100), shared dictionary ----
ME, SD, m = 512, 512, 100
Phi = (crandn(SD, ME) / ME ** 0.5).contiguous(); v = crandn(B, ME)
xh = torch.empty(B, SD, dtype=torch.complex64, device=dev); idx = torch.empty(B, m, dtype=torch.int32, device=dev); amb = torch.zeros(B, dtype=torch.int32, device=dev)
ms = timed(lambda: h.check(L.jstsp_omp(h.ptr, F32, DEV, ME, SD, m, B, p(Phi), 0, p(v), ME, p(xh), SD, p(idx), None, p(amb), 1e-4)))
phi0, v0 = Phi.cpu().numpy().T, v[0].cpu().numpy()
emit("OMP", [ME, SD, m], B, ms, m * 8 * ME * SD, cpu_rate(lambda: est.omp_literal(phi0, v0, m), 1),
     "bytes = the dictionary streamed once per iteration (shared by all trials, so it is served from L2: frac > 1 of HBM is expected)", flops_per_unit=m * 8 * ME * SD)

# ---- vamp.m, 256 x 1024 (config 3, T = 5), 100 iterations, shared operator ----
mm, nn, NIT = 256, 1024, 100
A = (crandn(nn, mm) / mm ** 0.5).contiguous()                     # stored (n, m) = column-major m x n
An = A.cpu().numpy().T.astype(np.complex128)
U, s, _ = np.linalg.svd(An, full_matrices=True)
Ud = torch.tensor(np.ascontiguousarray(U.T), dtype=torch.complex64, device=dev); dd = torch.tensor(s ** 2, dtype=torch.float32, device=dev)
yv = crandn(B, mm); xv = torch.empty(B, nn, dtype=torch.complex64, device=dev); sg, Ln = one(1.0), one(50.0)
ms = timed(lambda: h.check(L.jstsp_vamp(h.ptr, F32, DEV, mm, nn, B, NIT, 0.85, p(yv), mm, p(A), 0, p(sg), p(Ln), p(Ud), 0, p(dd), 0, p(xv), nn)))
y0 = yv[0].cpu().numpy()
emit("vamp", [mm, nn, NIT], B, ms, NIT * 8 * (2 * mm * nn + 2 * mm * mm), cpu_rate(lambda: ovamp.vamp_literal(y0, An, 1.0, 50), 1),
     "bytes = A, A^H, U, U^H streamed once per iteration (shared operator, L2-resident)", flops_per_unit=NIT * 8 * (2 * mm * nn + 2 * mm * mm))

# ---- driver-side metric / parameters at the metric shape (plot_errorVSsnr.m:127-130,138) and the rate metric ----
N_, M_, G_, P_ = 16, 1024, 16, 256
Yp = crandn(B, M_, N_); Zb = crandn(B, P_, G_); Se = (Zb + 0.1 * crandn(B, P_, G_)).contiguous()
tY, tZ, rh, nm = one(0.0), one(0.0), one(0.0), one(0.0)
ms = timed(lambda: h.check(L.jstsp_admm_parameters(h.ptr, F32, DEV, N_, M_, G_, P_, B, 6, p(Yp), N_ * M_, p(Zb), G_ * P_, p(tY), p(tZ), p(rh))))
yp0, zb0 = Yp[0].cpu().numpy().T, Zb[0].cpu().numpy().T
emit("admm_parameters", [N_, M_, G_, P_], B, ms, 8 * (N_ * M_ + G_ * P_), cpu_rate(lambda: est.admm_parameters(yp0, zb0, "sigma6"), 20), "read Y and Zbar once")
ms = timed(lambda: h.check(L.jstsp_nmse(h.ptr, F32, DEV, G_, P_, B, p(Se), G_ * P_, p(Zb), G_ * P_, p(nm))))
se0 = Se[0].cpu().numpy().T
emit("nmse", [G_, P_], B, ms, 2 * 8 * G_ * P_, cpu_rate(lambda: est.nmse(se0, zb0), 50), "read S and Zbar once; two spectral norms")
sc, rt = one(0.3), one(0.0)
ms = timed(lambda: h.check(L.jstsp_log2det_rate(h.ptr, F32, DEV, G_, P_, B, p(Zb), G_ * P_, p(sc), p(rt))))
emit("log2det_rate", [G_, P_], B, ms, 8 * G_ * P_, cpu_rate(lambda: est.log2det_rate(zb0, 0.3), 50), "read X once; Gram + fp64 Jacobi eigenvalues")

if world > 1:
    dist.destroy_process_group()
