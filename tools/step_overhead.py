"""Fixed cost of one device-resident call of the structured entry (setup kernels, memsets, host gaps) against the solve itself:
times the call at Imax = 0, 1, 2 and 100.  Developer tool."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import synth
from jstsp19_b200.engine import AdmmEngine
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 888
dev = torch.device("cuda", 0)
data = synth.make_batch(synth.METRIC, nb, torch.zeros(nb, dtype=torch.float64), seed=1, device=dev)
eng = AdmmEngine(0, "f32")
S = torch.empty(nb, synth.METRIC.P, synth.METRIC.Nr, dtype=torch.complex64, device=dev)
for imax in (0, 1, 2, 100):
    f = lambda: eng.proposed_algorithm_psi(data["subY"], data["Omega"], data["A"], data["Dt"], data["Psi"], imax, data["tau_Y"], data["tau_Z"], data["rho"], "approximate", S_out=S)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(4):
        f()
    e1.record(); th = time.perf_counter() - t0; torch.cuda.synchronize()
    print(f"Imax {imax:3d}: {e0.elapsed_time(e1) / 4:8.2f} ms per call on the device, host returned after {th / 4 * 1e3:8.2f} ms per call")
