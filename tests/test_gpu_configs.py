"""GPU parity at the shapes of the other BASELINE.json configs, at parity-test scale:
config 3 = frame-length sweep (plot_errorVSframelength.m:8-23: Nt=8, Nr=32, L=4, Mr=4, T in {5,15,25,35}, fft combiner),
config 4 = large array / delay sweep (plot_errorVSdelays.m: L in {2..10}; N = 64 receive rows), through the C ABI."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle import fixtures as fx
from oracle import vamp as ovamp
from oracle.matlab_compat import vec

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("T", [5, 15, 25, 35])
def test_framelength_sweep_proposed_and_mc(T):
    import jstsp19_b200 as jb
    shape = fx.Shape(Nt=8, Nr=32, L=4, Mr=4, T=T, combiner="fft")
    t = fx.make_trial(shape, 15.0, 300 + T)
    args = (t["subY"], t["Omega"], t["A"], t["B"], 100, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
    S0, Y0, _ = est.proposed_algorithm_structured(*args)
    for precision, tol in (("f64", 1e-9), ("f32", 2e-5)):
        S1, Y1, _ = jb.proposed_algorithm(*args, precision=precision)
        assert _rel(S1, S0) < tol and _rel(Y1, Y0) < tol, (T, precision, _rel(S1, S0))
    # the factor entry must agree with the dense one on this (non-tensor-core) shape
    S2 = jb.proposed_algorithm_psi(t["subY"], t["Omega"], t["A"], t["Dt"], t["Psi_bar"], 100, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", precision="f64", nargout=1)
    assert _rel(S2, S0) < 1e-9
    X0 = est.mc_svt(t["subY"], t["Omega"], 30, t["tau_Y"], 0.1)                 # plot_errorVSsnr.m:152 (commented call shape)
    X1 = jb.mc_svt(t["subY"], t["Omega"], 30, t["tau_Y"], 0.1)
    assert _rel(X1, X0) < 1e-8


@pytest.mark.parametrize("T", [5, 35])
def test_framelength_sweep_vamp_and_omp(T):
    """vamp(y, Phi, 1, numOfnz) / OMP with Phi = kron(B.', A), y = vec(Y) (plot_errorVSframelength.m:78-79,99): 256 x 1024 and 1024 x 1024."""
    import jstsp19_b200 as jb
    shape = fx.Shape(Nt=8, Nr=32, L=4, Mr=4, T=T, combiner="fft")
    t = fx.make_trial(shape, 15.0, 400 + T)
    c = fx.conventional_problem(t)
    Phi, y = np.kron(c["B"].T, c["A"]), vec(c["Y"])
    assert Phi.shape == ({5: 256, 35: 1024}[T], 1024)
    for nit, tol in ((5, 1e-9), (20, 1e-3)):           # the recursion amplifies rounding on these ill-conditioned systems (see test_vamp_config0_system)
        x0 = ovamp.vamp_literal(y, Phi, 1.0, 50, nit=nit)
        x1 = jb.vamp(y, Phi, 1.0, 50, nit=nit)
        assert _rel(x1, x0) < tol, (nit, _rel(x1, x0))
    xh0, i0, _, _ = est.omp_kron_structured(c["A"], c["B"], c["Y"], 20)
    xh1, i1, _, _, amb = jb.OMP_kron(c["A"], c["B"], c["Y"], 20, return_ambiguous=True)
    xh2, i2, _, _ = jb.OMP(Phi, y, 20)
    assert amb == 0 and i1 == i0 and i2 == i0 and _rel(xh1, xh0) < 1e-8 and _rel(xh2, xh0) < 1e-8


@pytest.mark.parametrize("L", [2, 6, 10])
def test_delay_sweep(L):
    """plot_errorVSdelays.m:45-49: L in {2,4,6,8,10}, rho from sigma_1 (:128)."""
    import jstsp19_b200 as jb
    shape = fx.Shape(Nt=4, Nr=32, L=L, Mr=4, T=5 * (L // 2))
    t = fx.make_trial(shape, 5.0, 500 + L, rho_rule="sigma1")
    args = (t["subY"], t["Omega"], t["A"], t["B"], 100, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
    S0, Y0, _ = est.proposed_algorithm_structured(*args)
    S1, Y1, _ = jb.proposed_algorithm(*args, precision="f64")
    assert _rel(S1, S0) < 1e-9 and _rel(Y1, Y0) < 1e-9
    S1, Y1, _ = jb.proposed_algorithm_angles(t["subY"], t["Omega"], t["indx_S"], t["A"], t["B"], 100, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", 20, precision="f32")
    S0a, _, _ = est.proposed_algorithm_structured(*args, indx_S=t["indx_S"])
    assert _rel(S1, S0a) < 2e-5


@pytest.mark.parametrize("Nr,precision,tol", [(64, "f32", 3e-5), (64, "f64", 1e-9), (48, "f64", 1e-9), (48, "f32", 3e-5)])
def test_large_array_rows(Nr, precision, tol):
    """Config 4 geometry at reduced length: Nr = 64 receive rows (N = G = 64), L = 8 taps, Nt = 16.  At 64 fp64 rows the
    residual kernel reads A / A'A through L2 instead of staging them (its product tile alone is 128 KB of shared memory)."""
    import jstsp19_b200 as jb
    shape = fx.Shape(Nt=16, Nr=Nr, L=8, Mr=8, T=8)
    t = fx.make_trial(shape, 5.0, 640)
    args = (t["subY"], t["Omega"], t["A"], t["B"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
    S0, Y0, _ = est.proposed_algorithm_structured(*args)
    S1, Y1, _ = jb.proposed_algorithm(*args, precision=precision)
    assert _rel(S1, S0) < tol and _rel(Y1, Y0) < tol, (_rel(S1, S0), _rel(Y1, Y0))
    S2 = jb.proposed_algorithm_psi(t["subY"], t["Omega"], t["A"], t["Dt"], t["Psi_bar"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", precision=precision, nargout=1)
    assert _rel(S2, S0) < tol


def test_fp64_exact_ls_at_64_rows():
    """'std' branch (proposed_algorithm.m:29,53) at config 4's row count in the MEX gateways' precision."""
    import jstsp19_b200 as jb
    t = fx.make_trial(fx.Shape(Nt=16, Nr=64, L=8, Mr=8, T=16), 5.0, 641)
    args = (t["subY"], t["Omega"], t["A"], t["B"], 10, t["tau_Y"], t["tau_Z"], t["rho"], "std")
    S0, Y0, c0 = est.proposed_algorithm_structured(*args)
    S1, Y1, c1 = jb.proposed_algorithm(*args, precision="f64")
    assert _rel(S1, S0) < 1e-7 and _rel(Y1, Y0) < 1e-7, (_rel(S1, S0), _rel(Y1, Y0))
    np.testing.assert_allclose(np.asarray(c1, dtype=np.float64)[:, :2], c0[:, :2], rtol=1e-6)


def test_more_than_64_rows_is_refused_by_name():
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import JstspError
    t = fx.make_trial(fx.Shape(Nt=8, Nr=72, L=2, Mr=8, T=4), 5.0, 642)
    with pytest.raises(JstspError) as e:
        jb.proposed_algorithm(t["subY"], t["Omega"], t["A"], t["B"], 5, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", precision="f64")
    assert e.value.code == -3 and "64" in str(e.value)
