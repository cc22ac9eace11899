"""GPU parity against the committed fixtures of tests/golden/solvers_small.npz (oracle outputs frozen by tools/make_golden.py): one small
trial through every entry point of the C ABI in fp64.  Needs neither /root/reference nor a run of the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "solvers_small.npz")


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def test_measurement_model(g):
    import jstsp19_b200 as jb
    H, Z, *_ = jb.wideband_mmwave_channel(2, 16, 4, 2, 3, 16, 4, normals=g["ch_normals"], uniforms=g["ch_uniforms"])
    assert _rel(H, g["ch_H"]) < 1e-12 and _rel(Z, g["ch_Zbar"]) < 1e-12
    M, Mr_e, Mr = (int(v) for v in g["hbf_dims"])
    Y, _, _, Om, _ = jb.proposed_hbf(g["hbf_H"], g["hbf_N"], None, M, Mr_e, Mr, g["hbf_W"], perm=g["hbf_perm"], pilots=g["hbf_pilots"])
    assert np.array_equal(Om, g["hbf_Omega"]) and _rel(Y, g["hbf_Y"]) < 1e-12


def test_proposed_algorithm(g):
    import jstsp19_b200 as jb
    a = (g["subY"], g["Omega"], g["A"], g["B"])
    p = (float(g["tau_Y"]), float(g["tau_Z"]), float(g["rho"]))
    S, Y, c = jb.proposed_algorithm(*a, 15, *p, "approximate")
    assert _rel(S, g["S_apx"]) < 1e-9 and _rel(Y, g["Y_apx"]) < 1e-9
    np.testing.assert_allclose(np.asarray(c, dtype=np.float64)[:, :2], g["conv_apx"][:, :2], rtol=1e-6)
    S, Y, _ = jb.proposed_algorithm(*a, 8, *p, "std")
    assert _rel(S, g["S_std"]) < 1e-7 and _rel(Y, g["Y_std"]) < 1e-7
    assert abs(jb.nmse(g["S_apx"], g["Zbar"]) - float(g["nmse_apx"])) < 1e-10
    np.testing.assert_allclose(jb.admm_parameters(g["subY"], g["Zbar"], "sigma6"), g["params"], rtol=1e-9)


def test_svt_family(g):
    import jstsp19_b200 as jb
    assert _rel(jb.svt(g["subY"], float(g["svt_tau"])), g["X_svt"]) < 1e-9
    assert _rel(jb.mc_svt(g["subY"], g["Omega"], 15, float(g["tau_Y"]), 0.1), g["X_mc_svt"]) < 1e-8
    X, c = jb.mc_admm(g["Htrue"], g["subY"], g["Omega"], 15, float(g["tau_Y"]), float(g["rho"]))
    assert _rel(X, g["X_mc_admm"]) < 1e-8
    np.testing.assert_allclose(np.asarray(c).reshape(-1), g["conv_mc_admm"].reshape(-1), rtol=1e-6)
    S, c = jb.sparse_admm(g["sp_H"], g["sp_OH"], g["sp_Dr"], g["sp_Dt"], 15)
    assert _rel(S, g["S_sparse"]) < 1e-9
    np.testing.assert_allclose(np.asarray(c).reshape(-1), g["conv_sparse"].reshape(-1), rtol=1e-5)


def test_omp_and_vamp(g):
    import jstsp19_b200 as jb
    x, idx, _, _ = jb.OMP(g["omp_A"], g["omp_v"], 8)
    assert [int(k) for k in idx] == [int(k) for k in g["omp_idx"]] and _rel(x, g["omp_x"]) < 1e-9
    for name in ("wide", "tall"):
        x = jb.vamp(g[f"vamp_{name}_y"], g[f"vamp_{name}_A"], 1e-4, 10, nit=20)
        assert _rel(x, g[f"vamp_{name}_x"]) < 1e-8, name
