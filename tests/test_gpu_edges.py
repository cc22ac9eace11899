"""Edge cases through the C ABI: ragged / odd shapes, single iterations, degenerate inputs, the reference's NaN guard, argument errors."""
import numpy as np
import pytest

from oracle import estimators as est

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _admm_problem(rng, N, M, G, P, frac=0.4):
    A = (rng.standard_normal((N, G)) + 1j * rng.standard_normal((N, G))) / np.sqrt(N)
    B = (rng.standard_normal((P, M)) + 1j * rng.standard_normal((P, M))) / np.sqrt(M)
    S = np.zeros((G, P), complex); S.flat[rng.choice(G * P, max(1, G * P // 8), replace=False)] = rng.standard_normal(max(1, G * P // 8)) + 1
    Om = (rng.random((N, M)) < frac).astype(float)
    Y = Om * (A @ S @ B + 0.01 * (rng.standard_normal((N, M)) + 1j * rng.standard_normal((N, M))))
    return Y, Om, A, B


@pytest.mark.parametrize("shape", [(7, 37, 5, 11), (3, 5, 3, 4), (17, 130, 9, 33), (1, 9, 1, 2)])
@pytest.mark.parametrize("precision,tol", [("f64", 1e-8), ("f32", 1e-4)])
def test_admm_ragged_shapes(shape, precision, tol):
    """Odd row counts, column counts that are no multiple of any tile, a single row."""
    import jstsp19_b200 as jb
    rng = np.random.default_rng(sum(shape))
    N, M, G, P = shape
    Y, Om, A, B = _admm_problem(rng, N, M, G, P)
    tY, tS, rho = 1.0 / max(np.linalg.norm(Y) ** 2, 1e-12), 0.05, 0.3
    S0, Y0, _ = est.proposed_algorithm_structured(Y, Om, A, B, 12, tY, tS, rho, "approximate", want_conv=False)
    S1, Y1 = jb.proposed_algorithm(Y, Om, A, B, 12, tY, tS, rho, "approximate", precision=precision, nargout=2)
    assert _rel(S1, S0) < tol and _rel(Y1, Y0) < tol, (_rel(S1, S0), _rel(Y1, Y0))


def test_admm_single_and_zero_iterations_and_zero_input():
    import jstsp19_b200 as jb
    rng = np.random.default_rng(1)
    Y, Om, A, B = _admm_problem(rng, 8, 24, 8, 6)
    S0, Y0, _ = est.proposed_algorithm_structured(Y, Om, A, B, 1, 0.1, 0.05, 0.3, "approximate", want_conv=False)
    S1, Y1 = jb.proposed_algorithm(Y, Om, A, B, 1, 0.1, 0.05, 0.3, "approximate", nargout=2)
    assert _rel(S1, S0) < 1e-9 and np.all(Y1 == 0)                  # first SVT input is all-zero: svt.m's NaN guard returns zeros
    # all-zero measurements: the line search of proposed_algorithm.m:48 is 0/0 in the first iteration and the reference then stops with an error
    # inside svd() on a NaN matrix at the second; the library takes a zero step instead (alpha = 0 when res' R res == 0) and returns the fixed
    # point S = 0 - a documented deviation on an input the reference cannot process at all
    for precision in ("f64", "f32"):
        Sz = jb.proposed_algorithm(0 * Y, Om, A, B, 5, 0.1, 0.05, 0.3, "approximate", nargout=1, precision=precision)
        assert np.all(Sz == 0)


def test_svt_zero_matrix_and_rank_deficient():
    """svt.m:7-13: any exactly-zero singular value (here: the zero matrix) makes the function return zeros."""
    import jstsp19_b200 as jb
    assert np.all(jb.svt(np.zeros((6, 20), complex), 0.5) == 0)
    rng = np.random.default_rng(2)
    Y = rng.standard_normal((5, 3)) @ rng.standard_normal((3, 40)) + 0j             # rank 3 < 5 rows: tiny but non-zero trailing singular values
    X0, X1 = est.svt_structured(Y, 0.7), jb.svt(Y, 0.7)
    assert _rel(X1, X0) < 1e-8


def test_omp_single_pick_and_tiny_dictionaries():
    import jstsp19_b200 as jb
    rng = np.random.default_rng(3)
    A = rng.standard_normal((5, 3)) + 1j * rng.standard_normal((5, 3))
    v = rng.standard_normal(5) + 1j * rng.standard_normal(5)
    x0, i0, _, _ = est.omp_literal(A, v, 1)
    x1, i1, _, _ = jb.OMP(A, v, 1)
    assert i1 == i0 and _rel(x1, x0) < 1e-10
    B = rng.standard_normal((2, 7)) + 1j * rng.standard_normal((2, 7))
    Y = rng.standard_normal((5, 7)) + 1j * rng.standard_normal((5, 7))
    x0, i0, _, _ = est.omp_kron_structured(A, B, Y, 4)
    x1, i1, _, _ = jb.OMP_kron(A, B, Y, 4)
    assert i1 == i0 and _rel(x1, x0) < 1e-9
    Z, sup, R = jb.somp(A, Y, 1)
    Z0, s0, R0 = est.somp_textbook(A, Y, 1)
    assert sup == s0 and _rel(Z, Z0) < 1e-10


def test_argument_errors_are_reported_not_crashes():
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import JstspError
    rng = np.random.default_rng(4)
    Y, Om, A, B = _admm_problem(rng, 4, 6, 4, 3)
    with pytest.raises(ValueError):
        jb.proposed_algorithm(Y, Om[:, :5], A, B, 3, 0.1, 0.1, 0.1, "approximate")
    with pytest.raises(JstspError) as e:
        jb.proposed_algorithm(Y, Om, A, B, -1, 0.1, 0.1, 0.1, "approximate", nargout=1)
    assert e.value.code == -1
    with pytest.raises(JstspError):
        jb.svt(np.zeros((65, 70), complex), 0.1)                     # Mr > 64 rows is outside the kernels' range and says so
    with pytest.raises(JstspError):
        jb.OMP_kron(A, B, Y, 0)
