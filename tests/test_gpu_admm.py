"""GPU parity: proposed_algorithm / proposed_algorithm_angles through the C ABI (HOST
buffers) against the fp64 oracle on identical seeded inputs."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu

# stated tolerances (relative Frobenius error of S and Y; relative error of the NMSE)
TOL = {"f64": dict(S=1e-9, nmse=1e-9), "f32": dict(S=2e-5, nmse=1e-4)}


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _run_case(shape, snr, seed, Imax, type_, precision, angles=False):
    import jstsp19_b200 as jb
    t = fx.make_trial(shape, snr, seed)
    args = (t["subY"], t["Omega"], t["A"], t["B"], Imax, t["tau_Y"], t["tau_Z"], t["rho"], type_)
    if angles:
        S0, Y0, c0 = est.proposed_algorithm_structured(*args, indx_S=t["indx_S"])
        S1, Y1, c1 = jb.proposed_algorithm_angles(t["subY"], t["Omega"], t["indx_S"], t["A"], t["B"], Imax,
                                                  t["tau_Y"], t["tau_Z"], t["rho"], type_, 20, precision=precision)
    else:
        S0, Y0, c0 = est.proposed_algorithm_structured(*args)
        S1, Y1, c1 = jb.proposed_algorithm(*args, precision=precision)
    n0, n1 = est.nmse(S0, t["Zbar"]), est.nmse(S1.astype(np.complex128), t["Zbar"])
    return dict(S=_rel(S1, S0), Y=_rel(Y1, Y0), nmse=abs(n1 - n0) / n0, c0=c0, c1=np.asarray(c1, dtype=np.float64))


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("shape,Imax", [(fx.TINY, 30), (fx.CONFIG0, 100)])
def test_approximate_matches_oracle(shape, Imax, precision):
    r = _run_case(shape, 5.0, 11, Imax, "approximate", precision)
    tol = TOL[precision]
    assert r["S"] < tol["S"] and r["Y"] < tol["S"], r
    assert r["nmse"] < tol["nmse"], r
    # convergence diagnostics (spectral norms, proposed_algorithm.m:51,67-69); column 3 is Inf at i=1
    c0, c1 = r["c0"], r["c1"]
    assert np.isinf(c1[0, 2]) and np.isinf(c0[0, 2])
    ctol = 1e-7 if precision == "f64" else 5e-3
    np.testing.assert_allclose(c1[:, :2], c0[:, :2], rtol=ctol)
    np.testing.assert_allclose(c1[1:, 2], c0[1:, 2], rtol=ctol)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_angles_matches_oracle(precision):
    r = _run_case(fx.CONFIG0, 5.0, 12, 60, "approximate", precision, angles=True)
    tol = TOL[precision]
    assert r["S"] < tol["S"] and r["nmse"] < tol["nmse"], r


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_std_branch_matches_oracle(precision):
    shape = fx.Shape(Nt=4, Nr=16, L=2, Mr=12, T=20)
    r = _run_case(shape, 5.0, 13, 30, "std", precision)
    tol = dict(f64=1e-8, f32=5e-4)[precision]
    assert r["S"] < tol and r["Y"] < tol, r


@pytest.mark.parametrize("precision", ["f64", "f32", "f32-tcgen05"])
def test_metric_shape_batch(precision, monkeypatch):
    """Nt=64, Nr=16, K=16, L=4 (BASELINE.json configs[1]); 3 trials in one batched call, per-trial B.
    "f32-tcgen05" runs the tensor-core kernel (admm_tc.cuh: 3xTF32 products, fp32 storage)."""
    import jstsp19_b200 as jb
    monkeypatch.setenv("JSTSP_TC", "1" if precision.endswith("tcgen05") else "0")
    precision = precision.split("-")[0]
    trials = [fx.make_trial(fx.METRIC, snr, 100 + k) for k, snr in enumerate([-15.0, 0.0, 15.0])]
    st = lambda k: np.stack([t[k] for t in trials])
    S1, Y1 = jb.proposed_algorithm(st("subY"), st("Omega"), st("A"), st("B"), 100,
                                   [t["tau_Y"] for t in trials], [t["tau_Z"] for t in trials], [t["rho"] for t in trials],
                                   "approximate", precision=precision, nargout=2)
    tol = TOL[precision]
    for k, t in enumerate(trials):
        S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 100, t["tau_Y"], t["tau_Z"],
                                                       t["rho"], "approximate", want_conv=False)
        n0, n1 = est.nmse(S0, t["Zbar"]), est.nmse(S1[k].astype(np.complex128), t["Zbar"])
        assert _rel(S1[k], S0) < tol["S"], (k, _rel(S1[k], S0))
        assert _rel(Y1[k], Y0) < tol["S"], (k, _rel(Y1[k], Y0))
        assert abs(n1 - n0) / n0 < tol["nmse"], (k, n0, n1)


def test_shared_dictionary_equals_per_trial():
    """ld_B = 0 (one pilot matrix for all trials) must give the same result as replicating B."""
    import jstsp19_b200 as jb
    t = fx.make_trial(fx.CONFIG0, 5.0, 21)
    subY = np.stack([t["subY"], 0.5 * t["subY"]])
    Om = np.stack([t["Omega"]] * 2)
    args = (100, [t["tau_Y"]] * 2, [t["tau_Z"]] * 2, [t["rho"]] * 2, "approximate")
    Sa = jb.proposed_algorithm(subY, Om, t["A"], t["B"], *args, nargout=1)
    Sb = jb.proposed_algorithm(subY, Om, np.stack([t["A"]] * 2), np.stack([t["B"]] * 2), *args, nargout=1)
    assert _rel(Sa, Sb) < 1e-12
