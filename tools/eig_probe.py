"""Phase timing of the Jacobi eigen-solve kernel (developer tool).  JSTSP_DBG_KERNEL=5 JSTSP_NO_OVERLAP=1"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import _lib, synth
from jstsp19_b200.engine import AdmmEngine
nb = 592
dev = torch.device("cuda", 0)
data = synth.make_batch(synth.METRIC, nb, torch.full((nb,), 5.0, dtype=torch.float64), seed=1, device=dev)
eng = AdmmEngine(0, "f32")
buf = torch.zeros(nb * 8 * 8, dtype=torch.int64, device=dev)
eng.proposed_algorithm(data["subY"], data["Omega"], data["A"], data["B"], 3, data["tau_Y"], data["tau_Z"], data["rho"])
torch.cuda.synchronize()
_lib.lib.jstsp_debug_buffer(eng.h.ptr, C.c_void_p(buf.data_ptr()))
eng.proposed_algorithm(data["subY"], data["Omega"], data["A"], data["B"], 20, data["tau_Y"], data["tau_Z"], data["rho"])
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 8)[:nb]
print("jacobi cycles median", np.median(t[:, 1] - t[:, 0]), "weights cycles", np.median(t[:, 3] - t[:, 1]), "sweeps", np.bincount(t[:, 2].astype(int)))
