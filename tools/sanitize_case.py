"""A short tour of every solver entry point at small iteration counts, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py dense      # shared-memory hazards of the FMA-pipe kernels

Each call is also compared with the fp64 oracle (test infrastructure), so a run that is clean but wrong still fails.
Groups: dense (proposed_algorithm*, svt family, sparse_admm, OMP, SOMP, VAMP, parameters), tc (the tcgen05 / TMA kernels:
Psi-domain ADMM at the metric shape, 3xTF32 dense ADMM, Kronecker OMP screen), large (the large-array route of csrc/admm_large.cuh at a
small shape, the on-device draws, the tiled measurement kernel)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jstsp19_b200 as jb                                   # noqa: E402
from oracle import estimators as est                        # noqa: E402
from oracle import fixtures as fx                           # noqa: E402
from oracle import vamp as ovamp                            # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def dense():
    t = fx.make_trial(fx.CONFIG0, 5.0, 7)
    for type_, it in (("approximate", 4), ("std", 3)):
        args = (t["subY"], t["Omega"], t["A"], t["B"], it, t["tau_Y"], t["tau_Z"], t["rho"], type_)
        S0, Y0, _ = est.proposed_algorithm_structured(*args)
        for prec, tol in (("f64", 1e-9), ("f32", 5e-4)):
            S1, Y1, _ = jb.proposed_algorithm(*args, precision=prec)
            assert rel(S1, S0) < tol and rel(Y1, Y0) < tol, (type_, prec, rel(S1, S0))
    S0, _, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 4, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", indx_S=t["indx_S"])
    S1, _, _ = jb.proposed_algorithm_angles(t["subY"], t["Omega"], t["indx_S"], t["A"], t["B"], 4, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", 20, precision="f32")
    assert rel(S1, S0) < 1e-4
    # 64 fp64 rows: k_res with its small operands read through L2
    t64 = fx.make_trial(fx.Shape(Nt=16, Nr=64, L=8, Mr=8, T=8), 5.0, 640)
    a64 = (t64["subY"], t64["Omega"], t64["A"], t64["B"], 2, t64["tau_Y"], t64["tau_Z"], t64["rho"], "approximate")
    assert rel(jb.proposed_algorithm(*a64, precision="f64", nargout=1), est.proposed_algorithm_structured(*a64)[0]) < 1e-9
    tau = 0.3 * np.linalg.norm(t["subY"], 2)
    assert rel(jb.svt(t["subY"], tau), est.svt_literal(t["subY"], tau)) < 1e-9
    assert rel(jb.mc_svt(t["subY"], t["Omega"], 3, t["tau_Y"], 0.1), est.mc_svt(t["subY"], t["Omega"], 3, t["tau_Y"], 0.1)) < 1e-9
    c = fx.conventional_problem(t)
    x0, i0 = est.omp_literal(c["Phi"], c["y"], 6)[:2]
    x1, i1 = jb.OMP(c["Phi"], c["y"], 6)[:2]
    assert [int(v) for v in i1] == [int(v) for v in i0] and rel(x1, x0) < 1e-9
    assert rel(jb.vamp(c["y"], c["Phi"], 1.0, 100, nit=3), ovamp.vamp_literal(c["y"], c["Phi"], 1.0, 100, nit=3)) < 1e-9
    At = c["Phi"][:, :200]
    assert rel(jb.vamp(c["y"], At, 1.0, 50, nit=3), ovamp.vamp_literal(c["y"], At, 1.0, 50, nit=3)) < 1e-9
    a = jb.admm_parameters(t["subY"], t["Zbar"], "sigma6")
    b = est.admm_parameters(t["subY"], t["Zbar"], "sigma6")
    assert np.allclose(a, b, rtol=1e-9)
    assert abs(jb.nmse(t["Zbar"] * 1.1, t["Zbar"]) - est.nmse(t["Zbar"] * 1.1, t["Zbar"])) < 1e-10
    print("dense group ok")


def tc():
    tm = fx.make_trial(fx.METRIC, 5.0, 4243)
    args = (tm["subY"], tm["Omega"], tm["A"], tm["B"], 3, tm["tau_Y"], tm["tau_Z"], tm["rho"], "approximate")
    S0, _, _ = est.proposed_algorithm_structured(*args, want_conv=False)
    S1 = jb.proposed_algorithm_pilots(tm["subY"], tm["Omega"], tm["A"], tm["Dt"], tm["pilots"], fx.METRIC.L, 3, tm["tau_Y"], tm["tau_Z"], tm["rho"],
                                      "approximate", precision="f32", nargout=1)
    from jstsp19_b200._lib import default_handle
    assert default_handle().last_path == 2 and rel(S1, S0) < 2e-5, (default_handle().last_path, rel(S1, S0))
    os.environ["JSTSP_TC"] = "1"
    S2 = jb.proposed_algorithm(*args, precision="f32", nargout=1)
    os.environ["JSTSP_TC"] = "0"
    assert rel(S2, S0) < 5e-5, rel(S2, S0)
    rng = np.random.default_rng(7)
    N, M, G, P = 64, 32, 64, 128
    A = np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(G)) / G) / np.sqrt(N)
    B = (rng.choice([-1, 1], (P, M)) + 1j * rng.choice([-1, 1], (P, M))) / np.sqrt(2 * M)
    Sx = np.zeros((G, P), complex)
    Sx.flat[rng.choice(G * P, 3, replace=False)] = rng.standard_normal(3) + 1j * rng.standard_normal(3) + 2
    Yk = A @ Sx @ B + 0.01 * (rng.standard_normal((N, M)) + 1j * rng.standard_normal((N, M)))
    _, i0, _, _ = est.omp_kron_structured(A, B, Yk, 3)
    _, i1, _, _ = jb.OMP_kron(A, B, Yk, 3, precision="f32", want_x_hat=False)
    assert i1 == i0, (i1, i0)
    print("tc group ok")


def large():
    import torch
    from jstsp19_b200 import synth
    from jstsp19_b200._lib import default_handle
    from jstsp19_b200.engine import TrialPipeline
    sh = fx.Shape(Nt=64, Nr=32, L=3, Mr=4, T=8)
    t = fx.make_trial(sh, 5.0, 99)
    S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 3, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
    S1, Y1 = jb.proposed_algorithm_pilots(t["subY"], t["Omega"], t["A"], t["Dt"], t["pilots"], sh.L, 3, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", precision="f32", nargout=2)
    assert default_handle().last_path == 3 and rel(S1, S0) < 5e-5 and rel(Y1, Y0) < 5e-5, (default_handle().last_path, rel(S1, S0), rel(Y1, Y0))
    pipe = TrialPipeline(synth.Shape(Nt=64, Nr=16, L=2, Mr=4, T=2), 0, "f32")
    nm = pipe.run(3, 5.0, seed=11, first_trial=2, imax=2)
    assert bool(torch.isfinite(nm).all())
    print("large group ok")


if __name__ == "__main__":
    groups = sys.argv[1:] or ["dense", "tc", "large"]
    for g in groups:
        {"dense": dense, "tc": tc, "large": large}[g]()
