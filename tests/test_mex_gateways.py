"""MEX gateways (jstsp19_b200/mex/*.c) driven through the mex.h shim (tests/mexharness.py).

CPU part: every gateway compiles, exports mexFunction, validates nargin / sizes with MATLAB-style error ids, and
fails loudly (jstsp:nogpu) when no CUDA device exists.  GPU part: each gateway reproduces the oracle on the
same seeded inputs, with the reference's RNG calls served through mexCallMATLAB in the reference's order."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle import fixtures as fx
from oracle import matlab_compat as mc
from oracle import system_model as sm
from oracle import vamp as ovamp

import mexharness as mh


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module", autouse=True)
def _built():
    mh.build()


@pytest.mark.parametrize("name", mh.GATEWAYS)
def test_gateway_builds_and_checks_nargin(name):
    g = mh.Gateway(name)
    assert hasattr(g.so, "mexFunction")
    with pytest.raises(mh.MexError) as e:
        g(1)                                   # no inputs at all
    assert e.value.ident == "jstsp:nargin"


def test_size_mismatch_is_a_matlab_error():
    g = mh.Gateway("proposed_algorithm")
    with pytest.raises(mh.MexError) as e:
        g(1, np.zeros((4, 6)), np.zeros((4, 5)), np.zeros((4, 4)), np.zeros((3, 6)), 5, 1.0, 1.0, 1.0, "approximate")
    assert e.value.ident == "jstsp:size"


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode")
def test_no_gpu_fails_loudly():
    g = mh.Gateway("svt")
    with pytest.raises(mh.MexError) as e:
        g(1, np.eye(4) + 0j, 0.1)
    assert e.value.ident == "jstsp:nogpu"


class Script:
    """Serves mexCallMATLAB from a RefRandom (the same seed drives the oracle) plus svd."""

    def __init__(self, seed):
        self.rng = mc.RefRandom(seed)

    def __call__(self, name, args):
        if name == "randn":
            shape = tuple(int(a.reshape(-1)[0]) for a in args)
            return [np.asarray(self.rng.randn(*shape)) if shape else np.asarray(self.rng.randn())]
        if name == "rand":
            return [np.asarray(self.rng.rand())]
        if name == "randperm":
            return [np.asarray(self.rng.randperm(int(args[0].reshape(-1)[0])), dtype=np.float64).reshape(1, -1)]
        if name == "randsrc":
            r, c = int(args[0].reshape(-1)[0]), int(args[1].reshape(-1)[0])
            return [np.asarray(self.rng.randsrc(r, c, np.asarray(args[2]).reshape(-1)))]
        if name == "randi":
            return [np.asarray(self.rng.randi(int(args[0].reshape(-1)[0]), int(args[1].reshape(-1)[0]), int(args[2].reshape(-1)[0])), dtype=np.float64)]
        if name == "svd":
            U, s, Vh = np.linalg.svd(args[0], full_matrices=True)
            S = np.zeros(args[0].shape)
            S[: s.size, : s.size] = np.diag(s)
            return [U, S, Vh.conj().T]
        raise KeyError(name)


@pytest.mark.gpu
def test_channel_gateway_reproduces_seeded_stream():
    g = mh.Gateway("wideband_mmwave_channel")
    g.set_matlab(Script(5))
    out = g(6, 3, 16, 8, 2, 3, 16, 8)
    ref = sm.wideband_mmwave_channel(3, 16, 8, 2, 3, 16, 8, mc.RefRandom(5))
    assert len(out) == 6
    for a, b in zip(out, ref):
        assert a.shape == b.shape and _rel(a, b) < 1e-12


@pytest.mark.gpu
def test_training_gateway_reproduces_seeded_stream():
    H = sm.wideband_mmwave_channel(2, 8, 4, 2, 3, 8, 4, mc.RefRandom(1))[0]
    g = mh.Gateway("wideband_hybBF_comm_system_training")
    g.set_matlab(Script(9))
    Yp, Yc, Wt, Pb, Om, Lr = g(6, H, 12, 0.1, 0.75)
    Yp0, Yc0, Wt0, Pb0, Om0, Lr0 = sm.wideband_hybBF_comm_system_training(H, 12, 0.1, 0.75, mc.RefRandom(9))
    assert int(Lr[0, 0]) == Lr0 and np.array_equal(Om, Om0)            # mask bit-exact
    assert np.all(Om.sum(axis=0) == Lr0)
    for a, b in ((Yp, Yp0), (Yc, Yc0), (Wt, Wt0), (Pb, Pb0)):
        assert a.shape == b.shape and _rel(a, b) < 1e-12


@pytest.mark.gpu
def test_hbf_gateways_match_oracle():
    s = fx.TINY
    t = fx.make_trial(s, 5.0, 3)
    Psi_i = sm.psi_i_literal(t["pilots"], s.M)
    g = mh.Gateway("proposed_hbf")
    g.set_matlab(Script(11))
    Y1, We1, Pb1, Om1, Yn1 = g(5, t["H"], t["N"], Psi_i, s.M, s.Mr_e, s.Mr, t["W"])
    Y0, We0, Pb0, Om0, Yn0 = sm.proposed_hbf(t["H"], t["N"], Psi_i, s.M, s.Mr_e, s.Mr, t["W"], mc.RefRandom(11))
    assert np.array_equal(Om1, Om0)
    for a, b in ((Y1, Y0), (We1, We0), (Pb1, Pb0), (Yn1, Yn0)):
        assert _rel(a, b) < 1e-12
    g2 = mh.Gateway("hbf")
    Yc1, Wc1, Pb2, Yn2 = g2(4, t["H"], t["N"], Psi_i, s.M, s.Nr, t["W"])
    Yc0, Wc0, Pb0, Yn0 = sm.hbf(t["H"], t["N"], Psi_i, s.M, s.Nr, t["W"])
    for a, b in ((Yc1, Yc0), (Wc1, Wc0), (Pb2, Pb0), (Yn2, Yn0)):
        assert _rel(a, b) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("type_", ["approximate", "std", "anything-else-is-std"])
def test_proposed_algorithm_gateway(type_):
    t = fx.make_trial(fx.TINY, 5.0, 11)
    g = mh.Gateway("proposed_algorithm")
    S1, Y1, c1 = g(3, t["subY"], t["Omega"], t["A"], t["B"], 30, t["tau_Y"], t["tau_Z"], t["rho"], type_)
    S0, Y0, c0 = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 30, t["tau_Y"], t["tau_Z"], t["rho"],
                                                   "approximate" if type_ == "approximate" else "std")
    assert _rel(S1, S0) < 1e-8 and _rel(Y1, Y0) < 1e-8
    assert c1.shape == (30, 3)
    np.testing.assert_allclose(c1[:, :2], c0[:, :2], rtol=1e-6)
    (S_only,) = g(1, t["subY"], t["Omega"], t["A"], t["B"], 30, t["tau_Y"], t["tau_Z"], t["rho"], type_)
    assert np.array_equal(S_only, S1)                                   # nargout does not change the result


@pytest.mark.gpu
def test_proposed_algorithm_angles_gateway():
    t = fx.make_trial(fx.CONFIG0, 5.0, 12)
    g = mh.Gateway("proposed_algorithm_angles")
    S1, Y1 = g(2, t["subY"], t["Omega"], t["indx_S"].astype(np.float64), t["A"], t["B"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", 20)
    S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate",
                                                  indx_S=t["indx_S"])
    assert _rel(S1, S0) < 1e-8 and _rel(Y1, Y0) < 1e-8


@pytest.mark.gpu
def test_proposed_algorithm_psi_gateway():
    """Factors (Dt, Psi_bar) instead of B, with and without the indx_S ranking: same numbers as the oracle on the dense B."""
    t = fx.make_trial(fx.CONFIG0, 5.0, 12)
    g = mh.Gateway("proposed_algorithm_psi")
    S1, Y1 = g(2, t["subY"], t["Omega"], t["A"], t["Dt"], t["Psi_bar"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
    S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
    assert _rel(S1, S0) < 1e-8 and _rel(Y1, Y0) < 1e-8
    (S2,) = g(1, t["subY"], t["Omega"], t["A"], t["Dt"], t["Psi_bar"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", t["indx_S"].astype(np.float64))
    S3, _, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate",
                                                 indx_S=t["indx_S"], want_conv=False)
    assert _rel(S2, S3) < 1e-8


@pytest.mark.gpu
def test_svt_family_gateways():
    t = fx.make_trial(fx.TINY, 5.0, 21)
    Yh, Om = t["subY"], t["Omega"]
    tau = 0.3 * np.linalg.norm(Yh, 2)
    (X1,) = mh.Gateway("svt")(1, Yh, tau)
    assert _rel(X1, est.svt_literal(Yh, tau)) < 1e-10
    (X2,) = mh.Gateway("mc_svt")(1, Yh, Om, 25, tau, 0.1)
    assert _rel(X2, est.mc_svt(Yh, Om, 25, tau, 0.1)) < 1e-9
    Htrue = t["Ynoiseless"][: Yh.shape[0], :]
    X3, c3 = mh.Gateway("mc_admm")(2, Htrue, Yh, Om, 25, tau, 0.1)
    X0, c0 = est.mc_admm_structured(Htrue, Yh, Om, 25, tau, 0.1)
    assert _rel(X3, X0) < 1e-9
    np.testing.assert_allclose(c3.reshape(-1), c0, rtol=1e-6)
    # real-valued (non-complex) inputs are accepted wherever complex is expected
    (Xr,) = mh.Gateway("svt")(1, np.asarray(Yh.real), tau)
    assert _rel(Xr, est.svt_literal(Yh.real + 0j, tau)) < 1e-10


@pytest.mark.gpu
def test_sparse_admm_gateway():
    rng = np.random.default_rng(3)
    Mr, Mt = 8, 4
    Dr, Dt = sm.dft_dictionary(Mr, Mr), sm.dft_dictionary(Mt, Mt)
    Htrue = rng.standard_normal((Mr, Mt)) + 1j * rng.standard_normal((Mr, Mt))
    OH = Htrue + 0.05 * (rng.standard_normal((Mr, Mt)) + 1j * rng.standard_normal((Mr, Mt)))
    S1, c1 = mh.Gateway("sparse_admm")(2, Htrue, OH, Dr, Dt, 20)
    S0, c0 = est.sparse_admm_structured(Htrue, OH, Dr, Dt, 20)
    assert _rel(S1, S0) < 1e-9
    np.testing.assert_allclose(c1.reshape(-1), c0, rtol=1e-6)


@pytest.mark.gpu
def test_omp_gateway_cell_output_and_support():
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((24, 40)) + 1j * rng.standard_normal((24, 40))) / np.sqrt(24)
    x = np.zeros(40, complex)
    x[[3, 17, 29]] = [1.0 + 0.5j, -0.8j, 0.6]
    v = A @ x
    g = mh.Gateway("OMP")
    xh, idx, vecho, tgt = g(4, A, v, 3, 10.0)
    x0, idx0, _, tgt0 = est.omp_literal(A, v, 3)
    assert isinstance(idx, list) and [int(c[0, 0]) for c in idx] == [int(i) for i in idx0]     # 1 x m cell of double scalars, bit-exact
    assert _rel(xh.reshape(-1), np.asarray(x0).reshape(-1)) < 1e-9 and _rel(tgt, tgt0) < 1e-12
    assert np.array_equal(vecho.reshape(-1), v)


@pytest.mark.gpu
def test_proposed_algorithm_pilots_gateway():
    """The pilot sequences s_k (rows of an Nt x M matrix) instead of Psi_bar: same numbers as the oracle on the dense B."""
    t = fx.make_trial(fx.CONFIG0, 5.0, 14)
    g = mh.Gateway("proposed_algorithm_pilots")
    S1, Y1 = g(2, t["subY"], t["Omega"], t["A"], t["Dt"], t["pilots"], float(fx.CONFIG0.L), 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
    S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 40, t["tau_Y"], t["tau_Z"], t["rho"], "approximate", want_conv=False)
    assert _rel(S1, S0) < 1e-8 and _rel(Y1, Y0) < 1e-8


@pytest.mark.gpu
def test_beamformer_and_qam_gateways():
    """createBeamformer / qam4mod by name, random codebooks and pilots drawn through MATLAB's own randsrc / randi."""
    g = mh.Gateway("createBeamformer")
    for kind in ("fft", "ps", "ZC", "quantized_4", "quantized"):
        for N in (5, 16, 70):
            (B,) = g(1, float(N), kind)
            assert _rel(B, sm.create_beamformer(N, kind)) < 1e-12, (kind, N)
    for kind in ("rand", "rand_ps"):
        g.set_matlab(Script(31))
        (B,) = g(1, 12.0, kind)
        assert _rel(B, sm.create_beamformer(12, kind, mc.RefRandom(31))) < 1e-12, kind
    q = mh.Gateway("qam4mod")
    q.set_matlab(Script(32))
    (s1,) = q(1, np.zeros((0, 0)), "mod", 40.0)
    assert np.array_equal(s1.reshape(-1), sm.qam4mod(40, mc.RefRandom(32)))
    soft = np.random.default_rng(3).standard_normal((6, 5)) + 1j * np.random.default_rng(4).standard_normal((6, 5))
    soft[0, 0] = 0.0; soft[1, 1] = 1.0; soft[2, 2] = -1j                  # points on the axes: the later rule of qam4mod.m:27-29 wins
    (s2,) = q(1, soft, "demod")
    assert np.array_equal(s2, sm.qam4demod(soft))
    with pytest.raises(mh.MexError):
        g(1, 8.0, "no_such_codebook")


@pytest.mark.gpu
def test_omp_kron_and_somp_gateways():
    rng = np.random.default_rng(6)
    N, M, G, P = 8, 12, 10, 9
    A = rng.standard_normal((N, G)) + 1j * rng.standard_normal((N, G))
    B = rng.standard_normal((P, M)) + 1j * rng.standard_normal((P, M))
    S = np.zeros((G, P), complex); S[2, 3] = 2 - 1j; S[7, 0] = 1.5j; S[4, 8] = -2
    Y = A @ S @ B
    xh, idx, xs, res = mh.Gateway("OMP_kron")(4, A, B, Y, 3)
    x0, i0, _, _ = est.omp_literal(np.kron(B.T, A), Y.reshape(-1, order="F"), 3)
    assert [int(c[0, 0]) for c in idx] == i0 and _rel(xh.reshape(-1), x0) < 1e-9
    assert _rel(xs.reshape(-1), np.array([x0[k - 1] for k in i0])) < 1e-9 and np.linalg.norm(res) < 1e-9 * np.linalg.norm(Y)
    Ys = Y @ np.linalg.pinv(B)                               # the drivers' Y*pinv(B) (plot_errorVSsnr.m:117)
    Z, sup, R = mh.Gateway("jstsp_somp")(3, A, Ys, 3)
    Z0, s0, R0 = est.somp_textbook(A, Ys, 3)
    assert [int(v) for v in np.asarray(sup).reshape(-1)] == s0 and _rel(Z, Z0) < 1e-9
    assert np.linalg.norm(R - R0) < 1e-9 * np.linalg.norm(Ys)        # three atoms recover Ys exactly: both residuals are rounding-level


@pytest.mark.gpu
def test_vamp_gateway_uses_matlab_svd():
    rng = np.random.default_rng(8)
    m, n = 24, 48
    A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(m)
    x = np.zeros(n, complex)
    x[rng.choice(n, 5, replace=False)] = rng.standard_normal(5) + 1j * rng.standard_normal(5)
    y = A @ x + 0.01 * (rng.standard_normal(m) + 1j * rng.standard_normal(m))
    g = mh.Gateway("vamp")
    g.set_matlab(Script(0))
    (x1,) = g(1, y, A, 1.0, 5)
    x0 = ovamp.vamp_literal(y, A, 1.0, 5)
    assert _rel(x1.reshape(-1), np.asarray(x0).reshape(-1)) < 1e-6
    # tall system: the gateway hands V of the same svd to the library (VampGlmEst.m:407-411)
    At = (rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))) / np.sqrt(n)
    xt = np.zeros(m, complex)
    xt[rng.choice(m, 3, replace=False)] = rng.standard_normal(3) + 1j * rng.standard_normal(3)
    yt = At @ xt + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    (x1,) = g(1, yt, At, 1.0, 3)
    x0 = ovamp.vamp_literal(yt, At, 1.0, 3)
    assert _rel(x1.reshape(-1), np.asarray(x0).reshape(-1)) < 1e-6


@pytest.mark.gpu
def test_ls_estimate_and_capacity_sweep_gateways():
    """The two gateways without a same-named reference function: ls_estimate (plot_errorVSsnr.m:83,117) and capacity_sweep (plot_capacity.m:45-64)."""
    from oracle import system_model as sm
    from oracle.matlab_compat import RefRandom, toeplitz_hermitian
    rng = np.random.default_rng(12)
    N, G, M, P = 8, 6, 40, 12
    A = rng.standard_normal((N, G)) + 1j * rng.standard_normal((N, G))
    B = rng.standard_normal((P, M)) + 1j * rng.standard_normal((P, M))
    Y = rng.standard_normal((N, M)) + 1j * rng.standard_normal((N, M))
    S, YpB = mh.Gateway("ls_estimate")(2, A, Y, B)
    assert _rel(S, est.ls_estimate(A, Y, B)) < 1e-9 and _rel(YpB, est.y_pinv_b(Y, B)) < 1e-9
    rr = RefRandom(3)
    Nt, Nr, L, T, Mr_e = 8, 16, 2, 6, 16
    H = sm.wideband_mmwave_channel(L, Nr, Nt, 2, 3, Nr, Nt, rr)[0]
    Psi_i = np.stack([toeplitz_hermitian(sm.qam4mod(T, rr)) for _ in range(Nt)], axis=2)
    Wz, Wq = sm.create_beamformer(Nr, "ZC"), sm.create_beamformer(Nr, "quantized")
    Yn = sm.hbf(H, np.zeros((Nr, T)), Psi_i, T, Nr, Wz)[3]
    ind = rr.randperm(Mr_e)
    scale = 1.0 / 10 ** (-15 / 10) / Nt
    mr_range = np.array([1.0, 4.0, 7.0, 10.0])
    (Cm,) = mh.Gateway("capacity_sweep")(1, Yn, Wz, Wq, mr_range, np.asarray(ind, dtype=np.float64), scale)
    assert Cm.shape == (4, 4)
    for i, Mr in enumerate(mr_range.astype(int)):
        ref = [est.capacity_literal(Yn, Wz, np.arange(1, Nr + 1), scale), est.capacity_literal(Yn, Wq, np.arange(1, Mr + 1), scale),
               est.capacity_literal(Yn, Wz, np.arange(1, Mr + 1), scale), est.capacity_literal(Yn, Wq, np.asarray(ind)[:Mr], scale)]
        assert np.allclose(Cm[i].real, ref, rtol=1e-8, atol=1e-8), (Mr, Cm[i], ref)
