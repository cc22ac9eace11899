"""Regenerates tests/golden/*.npz from the oracle (the reference is MATLAB and cannot run in this
image, so these are ORACLE outputs frozen as regression fixtures, not reference outputs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import estimators as est  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
t = fx.make_trial(fx.TINY, 5.0, 777)
Imax = 20
S, Y, c = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], Imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
Sa, Ya, ca = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], Imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate",
                                               indx_S=t["indx_S"])
np.savez_compressed(os.path.join(out, "admm_tiny.npz"), subY=t["subY"], Omega=t["Omega"], A=t["A"], B=t["B"], Imax=Imax,
                    tau_Y=t["tau_Y"], tau_S=t["tau_Z"], rho=t["rho"], S=S, Y=Y, conv=c, indx_S=t["indx_S"], S_angles=Sa, Zbar=t["Zbar"])
print("wrote", os.path.join(out, "admm_tiny.npz"))


# ---- one small trial through every solver and the measurement model (solvers_small.npz) ------------------------------------------
from oracle import matlab_compat as mc  # noqa: E402
from oracle import system_model as sm  # noqa: E402
from oracle import vamp as ovamp  # noqa: E402


class Rec(mc.RefRandom):
    """RefRandom that records its draws so that they can be replayed into the library."""

    def __init__(self, seed):
        super().__init__(seed)
        self.normals, self.uniforms, self.perms = [], [], []

    def randn(self, *shape):
        v = super().randn(*shape); self.normals.append(np.array(v)); return v

    def rand(self, *shape):
        v = super().rand(*shape); self.uniforms.append(np.array(v)); return v

    def randperm(self, n):
        v = super().randperm(n); self.perms.append(np.array(v)); return v


def solvers_small():
    s = fx.Shape(Nt=4, Nr=16, L=2, Mr=8, T=10)
    t = fx.make_trial(s, 8.0, 4711)
    g = dict(subY=t["subY"], Omega=t["Omega"], A=t["A"], B=t["B"], Zbar=t["Zbar"], tau_Y=t["tau_Y"], tau_Z=t["tau_Z"], rho=t["rho"])
    # channel and measurement model from recorded draws (wideband_mmwave_channel.m, proposed_hbf.m)
    rng = Rec(99)
    H, Zb, Ar, At, Dr, Dt = sm.wideband_mmwave_channel(2, 16, 4, 2, 3, 16, 4, rng)
    g.update(ch_normals=np.array(rng.normals).reshape(-1, 2), ch_uniforms=np.array(rng.uniforms).reshape(-1, 2), ch_H=H, ch_Zbar=Zb)
    rng = Rec(100)
    Psi_i = sm.psi_i_literal(t["pilots"], s.M)
    Y, We, Pb, Om, Yn = sm.proposed_hbf(t["H"], t["N"], Psi_i, s.M, s.Mr_e, s.Mr, t["W"], rng)
    g.update(hbf_H=t["H"], hbf_N=t["N"], hbf_pilots=t["pilots"], hbf_W=t["W"], hbf_perm=np.stack(rng.perms), hbf_Y=Y, hbf_Omega=Om, hbf_dims=np.array([s.M, s.Mr_e, s.Mr]))
    # estimators
    args = (t["subY"], t["Omega"], t["A"], t["B"])
    g["S_std"], g["Y_std"], _ = est.proposed_algorithm_structured(*args, 8, t["tau_Y"], t["tau_Z"], t["rho"], "std")
    g["S_apx"], g["Y_apx"], g["conv_apx"] = est.proposed_algorithm_structured(*args, 15, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
    g["svt_tau"] = 0.3 * np.linalg.norm(t["subY"], 2)
    g["X_svt"] = est.svt_literal(t["subY"], g["svt_tau"])
    g["X_mc_svt"] = est.mc_svt(t["subY"], t["Omega"], 15, t["tau_Y"], 0.1)
    g["Htrue"] = t["W_e"].conj().T @ t["Ynoiseless"]
    g["X_mc_admm"], g["conv_mc_admm"] = est.mc_admm_structured(g["Htrue"], t["subY"], t["Omega"], 15, t["tau_Y"], t["rho"])
    r = np.random.default_rng(5)
    g["sp_H"] = t["H"][:, :, 0]
    g["sp_OH"] = g["sp_H"] + 0.05 * (r.standard_normal(g["sp_H"].shape) + 1j * r.standard_normal(g["sp_H"].shape))
    g["sp_Dr"], g["sp_Dt"] = t["Dr"], t["Dt"]
    g["S_sparse"], g["conv_sparse"] = est.sparse_admm_structured(g["sp_H"], g["sp_OH"], t["Dr"], t["Dt"], 15)
    Ao = (r.standard_normal((48, 120)) + 1j * r.standard_normal((48, 120))) / np.sqrt(48)
    xo = np.zeros(120, complex); xo[r.choice(120, 6, replace=False)] = r.standard_normal(6) + 1j * r.standard_normal(6) + 2
    vo = Ao @ xo + 0.01 * (r.standard_normal(48) + 1j * r.standard_normal(48))
    xh, idx, _, _ = est.omp_literal(Ao, vo, 8)
    g.update(omp_A=Ao, omp_v=vo, omp_x=xh, omp_idx=np.array(idx))
    for name, (m, n) in (("wide", (40, 80)), ("tall", (80, 40))):
        Av = (r.standard_normal((m, n)) + 1j * r.standard_normal((m, n))) / np.sqrt(2 * m)
        xv = np.zeros(n, complex); xv[r.choice(n, 5, replace=False)] = 3 * (r.standard_normal(5) + 1j * r.standard_normal(5))
        yv = Av @ xv + 0.01 * (r.standard_normal(m) + 1j * r.standard_normal(m))
        g.update({f"vamp_{name}_A": Av, f"vamp_{name}_y": yv, f"vamp_{name}_x": ovamp.vamp_literal(yv, Av, 1e-4, 10, nit=20)})
    g["params"] = np.array(est.admm_parameters(t["subY"], t["Zbar"], "sigma6"))
    g["nmse_apx"] = est.nmse(g["S_apx"], t["Zbar"])
    np.savez_compressed(os.path.join(out, "solvers_small.npz"), **g)
    print("wrote", os.path.join(out, "solvers_small.npz"))


solvers_small()
