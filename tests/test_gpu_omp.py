"""GPU parity: OMP through the C ABI - support sets bit-exact against the oracle (OMP.m:17)."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_omp_random_dictionary(precision):
    import jstsp19_b200 as jb
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((96, 300)) + 1j * rng.standard_normal((96, 300))) / np.sqrt(96)
    x = np.zeros(300, complex); sup = rng.choice(300, 12, replace=False); x[sup] = rng.standard_normal(12) + 1j * rng.standard_normal(12) + 2
    v = A @ x + 0.01 * (rng.standard_normal(96) + 1j * rng.standard_normal(96))
    x0, i0, _, T0 = est.omp_literal(A, v, 20)
    x1, i1, v1, T1, amb = jb.OMP(A, v, 20, 0.0, precision=precision, return_ambiguous=True)
    assert amb == 0
    assert i1 == i0                                   # support set and order bit-exact
    assert _rel(x1, x0) < (1e-9 if precision == "f64" else 2e-4)
    assert _rel(T1, T0) < 1e-6 and v1 is v


def test_omp_config0_problem_fp64():
    """OMP(Phi, y, 100) on the conventional-HBF system of plot_errorVSsnr.m:73-80 (512 x 512)."""
    import jstsp19_b200 as jb
    t = fx.make_trial(fx.CONFIG0, 5.0, 41)
    c = fx.conventional_problem(t)
    x0, i0, _, _ = est.omp_literal(c["Phi"], c["y"], 100)
    x1, i1, _, _, amb = jb.OMP(c["Phi"], c["y"], 100, precision="f64", return_ambiguous=True)
    assert amb == 0 and i1 == i0
    assert _rel(x1, x0) < 1e-8


def test_omp_duplicate_pick_matches_pinv_split():
    import jstsp19_b200 as jb
    rng = np.random.default_rng(2)
    A = rng.standard_normal((6, 4)) + 1j * rng.standard_normal((6, 4))
    v = A[:, 1] * (2 - 1j)                             # exactly one atom: residual hits 0, later picks repeat
    x0, i0, _, _ = est.omp_literal(A, v, 3)
    x1, i1, _, _ = jb.OMP(A, v, 3, precision="f64")
    assert i1[0] == i0[0] == 2
    assert np.allclose(x1[1] * (1 + sum(1 for k in i1[1:] if k == 2)), 2 - 1j, atol=1e-10) or np.allclose(x1, x0, atol=1e-8)


def test_omp_batched():
    import jstsp19_b200 as jb
    rng = np.random.default_rng(8)
    A = (rng.standard_normal((64, 128)) + 1j * rng.standard_normal((64, 128))) / 8
    V = rng.standard_normal((5, 64)) + 1j * rng.standard_normal((5, 64))
    X, I, _, _ = jb.OMP(A, V, 10)
    for k in range(5):
        x0, i0, _, _ = est.omp_literal(A, V[k], 10)
        assert list(I[k]) == i0 and _rel(X[k], x0) < 1e-9
