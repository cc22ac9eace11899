# host-path ping-pong check: a 600-trial HOST call (3 passes of 200 via set_chunk... or automatic 2x300) must equal per-trial results
import numpy as np, sys
sys.path.insert(0, "/root/repo")
import jstsp19_b200 as jb
from oracle import fixtures as fx
t = fx.make_trial(fx.TINY, 5.0, 3)
nbt = 700
rng = np.random.default_rng(0)
scale = rng.uniform(0.5, 1.5, nbt)
subY = np.stack([t["subY"] * s for s in scale]); Om = np.stack([t["Omega"]] * nbt)
args = (30, [t["tau_Y"]] * nbt, [t["tau_Z"]] * nbt, [t["rho"]] * nbt, "approximate")
S, Y = jb.proposed_algorithm(subY, Om, t["A"], t["B"], *args, precision="f64", nargout=2)
for k in (0, 299, 300, 349, 350, 699):
    Sk, Yk = jb.proposed_algorithm(subY[k], Om[k], t["A"], t["B"], 30, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", precision="f64", nargout=2)
    print(k, np.abs(S[k] - Sk).max(), np.abs(Y[k] - Yk).max())
