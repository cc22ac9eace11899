"""Precision study (CPU, NumPy): how many TF32 terms do the two big contractions of
proposed_algorithm ('approximate') need?  Emulates the fused tensor-core path: everything in fp32
except T1 = Kt B^H and Xs = (A S) B, which use tf32-truncated operands in 1, 2 or 3 product terms
(fp32 accumulation is emulated in fp64: the accumulation error is not what is studied here).
Prints the relative error of S and of the NMSE against the fp64 oracle.  Developer tool, not shipped."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import estimators as est, fixtures as fx


def trunc_tf32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = trunc_tf32(x)
    lo = trunc_tf32((x.astype(np.float32) - hi))
    return hi.astype(np.float64), lo.astype(np.float64)


def csplit(Z):
    rh, rl = split(Z.real); ih, il = split(Z.imag)
    return rh + 1j * ih, rl + 1j * il


def mm(a, b, mode):
    """a small operand, b big operand (the dictionary); mode: 'f32' exact, '1x', 'big_hi' (small split only), '3x'."""
    if mode == "f32":
        return a @ b
    ah, al = csplit(a); bh, bl = csplit(b)
    if mode == "1x":
        return ah @ bh
    if mode == "big_hi":
        return ah @ bh + al @ bh
    return ah @ bh + al @ bh + ah @ bl


def run(t, Imax, mode):
    subY, Om, A, B = t["subY"], t["Omega"], t["A"], t["B"]
    tau_Y, tau_S, rho = t["tau_Y"], t["tau_Z"], t["rho"]
    f = lambda z: z.astype(np.complex64).astype(np.complex128)
    N, M = subY.shape; G, P = A.shape[1], B.shape[0]
    X = np.zeros((N, M), complex); V1 = X.copy(); V2 = X.copy(); C = X.copy(); Xs = X.copy()
    V = np.zeros((G, P), complex)
    D = 1.0 / (Om + 2 * rho); AH = A.conj().T; BH = B.conj().T; AHA = AH @ A; BBH = f(B @ BH)
    for i in range(Imax):
        Y = f(est.svt_structured(X - V1 / rho, tau_Y / rho))
        X = f((V1 + rho * Y + subY + V2 + rho * C + rho * Xs) * D)
        Kt = f(X - V2 / rho - C)
        T1 = f(mm(Kt, BH, mode))
        Res = f(AH @ T1 - AHA @ f(V @ BBH))
        Q = f(AHA @ f(Res @ BBH))
        alpha = np.vdot(Res, Res).real / np.vdot(Res, Q).real
        V = f(V + alpha * Res)
        S = f(est.soft_complex(V, tau_S / rho))
        Xs = f(mm(f(A @ S), B, mode))
        C = f(rho / (rho + 1) * (X - Xs - V2 / rho))
        V1 = f(V1 + rho * (Y - X)); V2 = f(V2 + rho * (C - X + Xs))
    return S


if __name__ == "__main__":
    for snr, seed in [(-15.0, 100), (0.0, 101), (15.0, 102)]:
        t = fx.make_trial(fx.METRIC, snr, seed)
        S0, _, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 100, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
        n0 = est.nmse(S0, t["Zbar"])
        for mode in ["f32", "1x", "big_hi", "3x"]:
            S = run(t, 100, mode)
            n1 = est.nmse(S, t["Zbar"])
            print(f"snr {snr:6.1f} mode {mode:7s} relS {np.linalg.norm(S - S0) / np.linalg.norm(S0):.3e} rel_nmse {abs(n1 - n0) / n0:.3e}", flush=True)
