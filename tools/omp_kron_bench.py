"""Times jstsp_omp_kron on device-resident operands at BASELINE config 2 (A 64 x 256, B 1024 x 128 per trial,
Y 64 x 128; Phi = kron(B.', A) would be 8192 x 262144) and prints one JSON line with the per-kernel split.
Usage: python tools/omp_kron_bench.py [--batch 296] [--m 50] [--steps 3]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=296)
ap.add_argument("--m", "--picks", dest="m", type=int, default=50, help="OMP iterations (--picks avoids the clash with torchrun's own --m* options)")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dbg", action="store_true")
ap.add_argument("--shape", type=int, nargs=4, default=[64, 128, 256, 1024])
a = ap.parse_args()
N, M, G, P = a.shape
# under torchrun every rank solves its own a.batch trials (weak scaling; trials are independent, one all-reduce at the end)
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(1 + rank)
A = (torch.exp(-2j * torch.pi * torch.outer(torch.arange(G, device=dev), torch.arange(N, device=dev)) / G) / N ** 0.5).to(torch.complex64).contiguous()   # (G,N) = col-major N x G
B = ((torch.randint(0, 2, (a.batch, M, P), generator=g, device=dev) * 2 - 1) + 1j * (torch.randint(0, 2, (a.batch, M, P), generator=g, device=dev) * 2 - 1)).to(torch.complex64) / (2 * M) ** 0.5
S = torch.zeros(a.batch, P, G, dtype=torch.complex64, device=dev)
idx = torch.randint(0, G * P, (a.batch, 12), generator=g, device=dev)
gains = (torch.randn(a.batch, 12, generator=g, device=dev) + 1j * torch.randn(a.batch, 12, generator=g, device=dev)).to(torch.complex64) * 2   # CN(0,.) path gains (wideband_mmwave_channel.m:19)
S.view(a.batch, -1).scatter_(1, idx, gains)
# Y' (M,N) = B' S' A'  in the stored (transposed) layout
Y = torch.matmul(torch.matmul(B, S), A.unsqueeze(0).expand(a.batch, -1, -1)).contiguous()
Y += 0.02 * torch.randn(Y.shape, generator=g, device=dev, dtype=torch.float32).to(torch.complex64)
h = _lib.Handle(local)
h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
index_set = torch.zeros(a.batch, a.m, dtype=torch.int32, device=dev)
xsel = torch.zeros(a.batch, a.m, dtype=torch.complex64, device=dev)
amb = torch.zeros(a.batch, dtype=torch.int32, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())


def run():
    h.check(_lib.lib.jstsp_omp_kron(h.ptr, _lib.F32, _lib.DEVICE, N, M, G, P, a.m, a.batch, p(A), 0, p(B), P * M, p(Y), N * M,
                                    None, 0, p(index_set), p(xsel), None, 0, p(amb), 1e-4))


run(); torch.cuda.synchronize()
if a.dbg:
    dbg = torch.zeros(2048, dtype=torch.int64, device=dev)
    _lib.lib.jstsp_debug_buffer(h.ptr, C.c_void_p(dbg.data_ptr()))
    run(); torch.cuda.synchronize()
    _lib.lib.jstsp_debug_buffer(h.ptr, None)
    d = dbg[:1184].view(148, 8).double().mean(0).tolist()
    print(json.dumps(dict(screen_updates=int(dbg[1200]), failed=int(dbg[1201]), candidates=int(dbg[1202]), whole_rows=int(dbg[1203]))), file=sys.stderr)
    print(json.dumps(dict(dbg_last_launch=dict(mma_total_cyc=d[0], mma_wait_full=d[1], mma_wait_tready=d[2], mma_wait_d2empty=d[3], items=d[4],
                                               prod_wait_empty=d[5], prod_total_cyc=d[6]))), file=sys.stderr)
_lib.lib.jstsp_profile(h.ptr, 2)
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                  # device time, max over ranks
    ms = float(t.item())
kern = {}
slot = 0
while True:
    t, n, name = C.c_double(), C.c_longlong(), C.c_char_p()
    if _lib.lib.jstsp_profile_read(h.ptr, slot, C.byref(t), C.byref(n), C.byref(name)) != 0:
        break
    if n.value:
        kern[name.value.decode()] = dict(ms_total=t.value, launches=n.value, avg_ms=t.value / n.value)
    slot += 1
_lib.lib.jstsp_profile(h.ptr, 0)
cmac = N * M * P + G * N * P                      # per iteration per trial: R B^H then A^H T
corr = (kern.get("omp_kron_corr_tc") or kern.get("omp_kron_corr", {})).get("avg_ms")
if rank == 0:
  print(json.dumps(dict(shape=a.shape, n_gpus=world, batch_per_gpu=a.batch, m=a.m, ms_per_call=ms, trials_per_s=world * a.batch / ms * 1e3,
                      corr_tflops=(8 * cmac * a.batch / (corr * 1e-3) / 1e12) if corr else None,
                        ambiguous_trials=int((amb > 0).sum()), kernels=kern)))
if world > 1:
    dist.destroy_process_group()
