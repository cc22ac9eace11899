"""GPU parity of the driver-side baselines and sweeps: the least-squares estimate pinv(A)*Y*pinv(B)
(plot_errorVSsnr.m:83,117, plot_errorVSsnr_approx.m:61,67), the capacity of a receiver design with its column selection
(plot_capacity.m:35-66) and the energy-efficiency power model (plot_ee.m:69-87), through the C ABI, against the oracle."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle import fixtures as fx
from oracle import system_model as sm

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


@pytest.mark.parametrize("precision,tol", [("f64", 1e-9), ("f32", 2e-3)])
def test_ls_estimate_default_shape(precision, tol):
    """plot_errorVSsnr.m defaults: A = W_c'*Dr 32 x 32, B 16 x 16 (T_hbf = 16), Y_hbf_nr 32 x 16."""
    import jstsp19_b200 as jb
    t = fx.make_trial(fx.CONFIG0, 5.0, 21)
    A = t["A"]
    B = t["B"][:, :16]
    rng = np.random.default_rng(3)
    Y = A @ t["Zbar"] @ B + 0.1 * (rng.standard_normal((A.shape[0], 16)) + 1j * rng.standard_normal((A.shape[0], 16)))
    S1, YpB1 = jb.ls_estimate(A, Y, B, precision=precision, want_YpinvB=True)
    S0, YpB0 = est.ls_estimate(A, Y, B), est.y_pinv_b(Y, B)
    cond = np.linalg.cond(B) * np.linalg.cond(A)
    assert _rel(S1, S0) < tol * max(1.0, cond / 1e2), (_rel(S1, S0), cond)
    assert _rel(YpB1, YpB0) < tol * max(1.0, np.linalg.cond(B) / 1e2)
    # NMSE of the baseline as the driver computes it (plot_errorVSsnr.m:84): same number within the tolerance
    n0, n1 = est.nmse(S0, t["Zbar"]), est.nmse(np.asarray(S1, complex), t["Zbar"])
    assert abs(n1 - n0) <= tol * max(1.0, cond / 1e2) * max(n0, 1e-3) * 10


@pytest.mark.parametrize("shape", [(16, 1024, 16, 256), (12, 40, 20, 9), (20, 9, 12, 30)])
def test_ls_estimate_batched_all_orientations(shape):
    """Metric shape (tall A, wide B, per-trial B) and the two transposed orientations (wide A / tall B): pinv through the short side."""
    import jstsp19_b200 as jb
    N, M, G, P = shape
    rng = np.random.default_rng(sum(shape))
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    A = cr(N, G) / np.sqrt(N)
    Bs = cr(3, P, M) / np.sqrt(M)
    Ys = cr(3, N, M)
    S1 = jb.ls_estimate(A, Ys, Bs, precision="f64")
    for k in range(3):
        assert _rel(S1[k], est.ls_estimate(A, Ys[k], Bs[k])) < 1e-8 * max(1.0, np.linalg.cond(A) * np.linalg.cond(Bs[k]) / 1e2)


def test_ls_after_std_estimator_matches_driver_lines():
    """plot_errorVSsnr_approx.m:60-62: S = pinv(A) * Y_proposed * pinv(B) on the estimator's second output."""
    import jstsp19_b200 as jb
    t = fx.make_trial(fx.Shape(Nt=4, Nr=16, L=2, Mr=12, T=18), 5.0, 33)       # P = 8 <= M = 72, G = 16 = N: the 'std' branch's rank conditions
    S, Yp = jb.proposed_algorithm(t["subY"], t["Omega"], t["A"], t["B"], 20, t["tau_Y"], t["tau_Z"], t["rho"], "std", precision="f64", nargout=2)
    S_ls = jb.ls_estimate(t["A"], Yp, t["B"], precision="f64")
    _, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"], 20, t["tau_Y"], t["tau_Z"], t["rho"], "std", want_conv=False)
    assert _rel(S_ls, est.ls_estimate(t["A"], Y0, t["B"])) < 1e-7


@pytest.mark.parametrize("precision,tol", [("f64", 1e-10), ("f32", 2e-5)])
def test_capacity_designs_and_column_selection(precision, tol):
    """plot_capacity.m:35-66 on a few trials: digital BF (all Nr columns), the conventional designs (first Mr columns, hbf.m:24) and the
    proposed design's random column subset W(:, ind(1:Mr)), every RF-chain count of Mr_range = 1:3:Mr_e."""
    import jstsp19_b200 as jb
    from oracle.matlab_compat import RefRandom, toeplitz_hermitian
    rng = RefRandom(5)
    Nt, Nr, L, T, Mr_e = 16, 32, 4, 5, 32
    scale = 1.0 / 10 ** (-15 / 10) / Nt                                         # 1/square_noise_variance*1/Nt, plot_capacity.m:19,47
    nb = 4
    Ys, Ws, inds = [], [], []
    for k in range(nb):
        H = sm.wideband_mmwave_channel(L, Nr, Nt, 2, 3, Nr, Nt, rng)[0]           # plot_capacity.m:36
        Psi_i = np.stack([toeplitz_hermitian(sm.qam4mod(T, rng)) for _ in range(Nt)], axis=2)      # :37-41
        W = sm.create_beamformer(Nr, "quantized" if k % 2 else "ZC")             # :45,50,55,61
        Y = sm.hbf(H, np.zeros((Nr, T)), Psi_i, T, Nr, W)[3]                     # noiseless received block (4th output, hbf.m:1)
        Ys.append(Y); Ws.append(W); inds.append(rng.randperm(Mr_e))               # ind = randperm(Mr_e), plot_capacity.m:63
    Ys, Ws, inds = np.stack(Ys), np.stack(Ws), np.stack(inds)
    for Mr in range(1, Mr_e + 1, 3):                                             # Mr_range, plot_capacity.m:11
        c_first = jb.capacity(Ys, Ws, Mr, scale, precision=precision)
        c_sel = jb.capacity(Ys, Ws, Mr, scale, cols=inds, precision=precision)
        for k in range(nb):
            r0 = est.capacity_literal(Ys[k], Ws[k], np.arange(1, Mr + 1), scale)
            r1 = est.capacity_literal(Ys[k], Ws[k], inds[k][:Mr], scale)
            assert abs(c_first[k] - r0) <= tol * max(1.0, abs(r0)), (Mr, k, c_first[k], r0)
            assert abs(c_sel[k] - r1) <= tol * max(1.0, abs(r1)), (Mr, k, c_sel[k], r1)
    c_dbf = jb.capacity(Ys, Ws, Nr, scale, precision=precision)                   # digital beamforming: all Nr columns (plot_capacity.m:46-47)
    for k in range(nb):
        assert abs(c_dbf[k] - est.capacity_literal(Ys[k], Ws[k], np.arange(1, Nr + 1), scale)) <= tol * 50


@pytest.mark.parametrize("precision,tol", [("f64", 1e-9), ("f32", 2e-4)])
def test_capacity_sweep_all_designs(precision, tol):
    """jstsp_capacity_sweep = the loop body of plot_capacity.m:35-66 for a whole Mr range in one call: the four receiver designs against the literal
    restatement, then the energy efficiency of plot_ee.m:84-87 from the mean rates and the power model."""
    import jstsp19_b200 as jb
    from oracle.matlab_compat import RefRandom, toeplitz_hermitian
    rng = RefRandom(9)
    Nt, Nr, L, T, Mr_e = 16, 32, 4, 5, 32
    scale = 1.0 / 10 ** (-15 / 10) / Nt
    nb, mr_range = 3, list(range(1, Mr_e + 1, 5))
    Wz, Wq = sm.create_beamformer(Nr, "ZC"), sm.create_beamformer(Nr, "quantized")
    Ys, inds = [], []
    for k in range(nb):
        H = sm.wideband_mmwave_channel(L, Nr, Nt, 2, 3, Nr, Nt, rng)[0]
        Psi_i = np.stack([toeplitz_hermitian(sm.qam4mod(T, rng)) for _ in range(Nt)], axis=2)
        Ys.append(sm.hbf(H, np.zeros((Nr, T)), Psi_i, T, Nr, Wz)[3]); inds.append(rng.randperm(Mr_e))
    Ys, inds = np.stack(Ys), np.stack(inds)
    out = jb.capacity_sweep(Ys, Wz, Wq, mr_range, inds, scale, precision=precision)
    assert out.shape == (len(mr_range), 4, nb)
    for i, Mr in enumerate(mr_range):
        for k in range(nb):
            ref = [est.capacity_literal(Ys[k], Wz, np.arange(1, Nr + 1), scale), est.capacity_literal(Ys[k], Wq, np.arange(1, Mr + 1), scale),
                   est.capacity_literal(Ys[k], Wz, np.arange(1, Mr + 1), scale), est.capacity_literal(Ys[k], Wq, inds[k][:Mr], scale)]
            for dsg in range(4):
                assert abs(out[i, dsg, k] - ref[dsg]) <= tol * max(1.0, abs(ref[dsg])) * (50 if dsg == 0 else 1), (Mr, dsg, k, out[i, dsg, k], ref[dsg])
        ee = jb.energy_efficiency(out[i].mean(axis=1), Nr, Mr, Mr_e)
        pw = est.ee_power_model(Nr, Mr, Mr_e)
        np.testing.assert_allclose(ee, out[i].mean(axis=1) / np.asarray(pw), rtol=1e-12)


def test_energy_efficiency_power_model():
    """plot_ee.m:69-87: the four power formulas and ee = mean capacity / power, over Mr_range = 1:3:Mr_e at Nr = 64."""
    import jstsp19_b200 as jb
    for Mr in range(1, 33, 3):
        p1, p0 = jb.power_model(64, Mr, 32), est.ee_power_model(64, Mr, 32)
        assert np.allclose(p1, p0, rtol=0, atol=1e-12)
        caps = (40.0 + Mr, 30.0 + Mr, 31.0 + Mr, 35.0 + Mr)
        assert np.allclose(jb.energy_efficiency(caps, 64, Mr, 32), [c / p for c, p in zip(caps, p0)], rtol=1e-14)
    assert abs(est.ee_power_model(64, 1, 32)[0] - (64 * 64 * 0.02 + 64 * 65 * 0.06)) < 1e-12      # digital beamforming, hand-computed
