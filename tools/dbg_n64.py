import numpy as np, sys
sys.path.insert(0, "/root/repo")
from oracle import fixtures as fx
import jstsp19_b200 as jb
shape = fx.Shape(Nt=16, Nr=64, L=8, Mr=8, T=8)
t = fx.make_trial(shape, 5.0, 640)
args = (t["subY"], t["Omega"], t["A"], t["B"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
for prec in ("f64", "f32"):
    try:
        S1, Y1, _ = jb.proposed_algorithm(*args, precision=prec)
        print(prec, "ok")
    except Exception as e:
        print(prec, "ERR", e)
