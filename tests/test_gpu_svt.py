"""GPU parity: svt / mc_svt / mc_admm through the C ABI against the oracle."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-9, "f32": 2e-5}


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("shape", [(8, 12), (32, 140), (16, 1024), (5, 7), (64, 300)])
def test_svt(shape, precision):
    import jstsp19_b200 as jb
    rng = np.random.default_rng(7)
    Y = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    Y[0] *= 30.0
    tau = 0.4 * np.linalg.norm(Y, 2)
    X0 = est.svt_literal(Y, tau)
    X1 = jb.svt(Y, tau, precision=precision)
    assert _rel(X1, X0) < TOL[precision]
    # svt of the zero matrix is zero (svt.m:7-13), threshold above sigma_max gives zero
    assert not np.any(jb.svt(np.zeros(shape, complex), 0.1, precision=precision))
    assert np.abs(jb.svt(Y, 2.0 * np.linalg.norm(Y, 2), precision=precision)).max() < 1e-6 * np.abs(Y).max()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_mc_svt_and_mc_admm(precision):
    import jstsp19_b200 as jb
    t = fx.make_trial(fx.CONFIG0, 10.0, 31)
    OH, Om = t["subY"], t["Omega"]
    X0 = est.mc_svt(OH, Om, 40, t["tau_Y"], 0.1)
    X1 = jb.mc_svt(OH, Om, 40, t["tau_Y"], 0.1, precision=precision)
    assert _rel(X1, X0) < 10 * TOL[precision]
    Htrue = t["W_e"].conj().T @ t["Ynoiseless"]
    X0, c0 = est.mc_admm_structured(Htrue, OH, Om, 40, t["tau_Y"], t["rho"])
    X1, c1 = jb.mc_admm(Htrue, OH, Om, 40, t["tau_Y"], t["rho"], precision=precision)
    assert _rel(X1, X0) < 10 * TOL[precision]
    np.testing.assert_allclose(c1, c0, rtol=1e-6 if precision == "f64" else 2e-3)


def test_batched_svt_matches_single():
    import jstsp19_b200 as jb
    rng = np.random.default_rng(9)
    Y = rng.standard_normal((5, 16, 200)) + 1j * rng.standard_normal((5, 16, 200))
    taus = np.linspace(1.0, 9.0, 5)
    Xb = jb.svt(Y, taus)
    for k in range(5):
        assert _rel(Xb[k], est.svt_literal(Y[k], taus[k])) < 1e-9
