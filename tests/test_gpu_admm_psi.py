"""GPU parity of jstsp_proposed_algorithm_psi: the estimator fed with the dictionary's factors (Dt, Psi_bar) as the
reference's drivers hold them before they form B (plot_errorVSsnr.m:132-136), against the fp64 oracle that follows
proposed_algorithm.m on the dense B built by those very lines.  Covers the Psi-domain tcgen05 kernel (path 2) and the
materialised-B route taken for anything without the Toeplitz / 4-QAM structure (path 1)."""
import numpy as np
import pytest

from oracle import estimators as est
from oracle import fixtures as fx
from oracle import system_model as sm

pytestmark = pytest.mark.gpu

TOL = {"f64": dict(S=1e-9, nmse=1e-9), "f32": dict(S=2e-5, nmse=1e-4)}


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _oracle(t, B=None, Imax=100, indx_S=None):
    S0, Y0, _ = est.proposed_algorithm_structured(t["subY"], t["Omega"], t["A"], t["B"] if B is None else B, Imax, t["tau_Y"], t["tau_Z"],
                                                   t["rho"], "approximate", want_conv=False, indx_S=indx_S)
    return S0, Y0


def _metric_trials():
    return [fx.make_trial(fx.METRIC, snr, 100 + k) for k, snr in enumerate([-15.0, 0.0, 15.0])]


def test_metric_shape_tensor_core_path():
    """Nt=64, Nr=16, K=16, L=4, per-trial 4-QAM Toeplitz pilots: must take the Psi-domain tcgen05 kernel and match the oracle."""
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import default_handle
    trials = _metric_trials()
    st = lambda k: np.stack([t[k] for t in trials])
    S1, Y1 = jb.proposed_algorithm_psi(st("subY"), st("Omega"), st("A"), trials[0]["Dt"], st("Psi_bar"), 100,
                                       [t["tau_Y"] for t in trials], [t["tau_Z"] for t in trials], [t["rho"] for t in trials],
                                       "approximate", precision="f32", nargout=2)
    assert default_handle().last_path == 2, "structured pilots did not reach the tensor-core kernel"
    for k, t in enumerate(trials):
        S0, Y0 = _oracle(t)
        n0, n1 = est.nmse(S0, t["Zbar"]), est.nmse(S1[k].astype(np.complex128), t["Zbar"])
        assert _rel(S1[k], S0) < TOL["f32"]["S"], (k, _rel(S1[k], S0))
        assert _rel(Y1[k], Y0) < TOL["f32"]["S"], (k, _rel(Y1[k], Y0))
        assert abs(n1 - n0) / n0 < TOL["f32"]["nmse"], (k, n0, n1)


def test_metric_shape_shared_pilots_and_angles():
    """One pilot matrix for all trials (ld_Psi = 0) + the growing support mask of proposed_algorithm_angles."""
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import default_handle
    t = fx.make_trial(fx.METRIC, 5.0, 31)
    subY = np.stack([t["subY"], 0.5 * t["subY"]])
    Om = np.stack([t["Omega"]] * 2)
    ix = np.stack([t["indx_S"]] * 2)
    S1 = jb.proposed_algorithm_psi(subY, Om, t["A"], t["Dt"], t["Psi_bar"], 40, [t["tau_Y"]] * 2, [t["tau_Z"]] * 2, [t["rho"]] * 2,
                                   "approximate", ix, precision="f32", nargout=1)
    assert default_handle().last_path == 2
    for k in range(2):
        tk = dict(t, subY=subY[k])
        S0, _ = _oracle(tk, Imax=40, indx_S=t["indx_S"])
        assert _rel(S1[k], S0) < TOL["f32"]["S"], (k, _rel(S1[k], S0))


def test_unstructured_pilots_take_dense_route():
    """Gaussian pilots (wideband_hybBF_comm_system_training.m:20-21 style, not exact in bf16) and a non-Toeplitz Psi_bar:
    same entry point, dense kernels on the device-built B, same answer."""
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import default_handle
    t = fx.make_trial(fx.METRIC, 5.0, 41)
    rng = np.random.default_rng(7)
    for kind in ("gaussian", "not_toeplitz"):
        if kind == "gaussian":
            pil = (rng.standard_normal(t["pilots"].shape) + 1j * rng.standard_normal(t["pilots"].shape)) / np.sqrt(2.0)
            Psi = sm.psi_bar_from_pilots(pil, fx.METRIC.M, fx.METRIC.L)
        else:
            Psi = t["Psi_bar"].copy()
            Psi[5, 700, 2] = -Psi[5, 700, 2]
        B = sm.dictionary_B(t["Dt"], Psi)
        S1 = jb.proposed_algorithm_psi(t["subY"], t["Omega"], t["A"], t["Dt"], Psi, 30, t["tau_Y"], t["tau_Z"], t["rho"], "approximate",
                                       precision="f32", nargout=1)
        assert default_handle().last_path == 1, kind
        S0, _ = _oracle(t, B=B, Imax=30)
        assert _rel(S1, S0) < TOL["f32"]["S"], (kind, _rel(S1, S0))


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_default_shape_factors_equal_dense(precision):
    """plot_errorVSsnr.m default sizes (Nt=4: outside the tensor-core kernel's shape) with diagnostics: factors == dense B."""
    import jstsp19_b200 as jb
    t = fx.make_trial(fx.CONFIG0, 5.0, 11)
    args = (100, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
    Sa, Ya, ca = jb.proposed_algorithm_psi(t["subY"], t["Omega"], t["A"], t["Dt"], t["Psi_bar"], *args, precision=precision)
    Sb, Yb, cb = jb.proposed_algorithm(t["subY"], t["Omega"], t["A"], t["B"], *args, precision=precision)
    tol = 1e-10 if precision == "f64" else 2e-5
    assert _rel(Sa, Sb) < tol and _rel(Ya, Yb) < tol
    S0, Y0 = _oracle(t)
    assert _rel(Sa, S0) < TOL[precision]["S"]


def test_device_engine_matches_host_call():
    """Device-resident tensors through engine.AdmmEngine.proposed_algorithm_psi == the HOST-buffer call."""
    import torch
    import jstsp19_b200 as jb
    from jstsp19_b200.engine import AdmmEngine
    trials = _metric_trials()[:2]
    st = lambda k: np.stack([t[k] for t in trials])
    tY, tZ, rh = ([t[k] for t in trials] for k in ("tau_Y", "tau_Z", "rho"))
    S_host = jb.proposed_algorithm_psi(st("subY"), st("Omega"), st("A"), trials[0]["Dt"], st("Psi_bar"), 25, tY, tZ, rh, "approximate",
                                       precision="f32", nargout=1)
    dev = torch.device("cuda", 0)
    cm = lambda x: torch.from_numpy(np.ascontiguousarray(np.swapaxes(x, -1, -2))).to(dev)
    eng = AdmmEngine(0, "f32")
    Psi = torch.from_numpy(np.ascontiguousarray(np.moveaxis(st("Psi_bar"), (-3, -2, -1), (-1, -2, -3)))).to(torch.complex64).to(dev)
    f64 = lambda v: torch.tensor(v, dtype=torch.float64, device=dev)
    S_dev = eng.proposed_algorithm_psi(cm(st("subY")).to(torch.complex64), cm(st("Omega")).to(torch.float32), cm(st("A")).to(torch.complex64),
                                       cm(trials[0]["Dt"]).to(torch.complex64)[None].contiguous(), Psi.contiguous(), 25, f64(tY), f64(tZ), f64(rh))
    torch.cuda.synchronize()
    assert eng.h.last_path == 2
    S_dev = np.swapaxes(S_dev.cpu().numpy(), -1, -2)
    assert _rel(S_dev, S_host) < 1e-6


@pytest.mark.parametrize("name,shape", [
    ("odd_taps", fx.Shape(Nt=64, Nr=16, L=3, Mr=4, T=4)),                 # L = 3: the last tap has no partner accumulator
    ("one_tap", fx.Shape(Nt=64, Nr=16, L=1, Mr=4, T=2)),                  # L = 1, a single 128-column chunk
    ("coarse_grid", fx.Shape(Nt=64, Nr=16, L=4, Mr=4, T=4, Gt=32)),       # Gt < Nt: Dt is not the unitary DFT -> generic rotations
    ("eight_taps", fx.Shape(Nt=64, Nr=16, L=8, Mr=6, T=2, Gt=16)),        # L = 8 (the kernel's maximum), tiny grid
])
def test_other_structured_shapes(name, shape):
    """Shapes around the metric one that still qualify for the tensor-core kernel (N = 16, Nt = 64, M % 128 == 0)."""
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import default_handle
    t = fx.make_trial(shape, 3.0, 77)
    S1, Y1 = jb.proposed_algorithm_psi(t["subY"], t["Omega"], t["A"], t["Dt"], t["Psi_bar"], 60, t["tau_Y"], t["tau_Z"], t["rho"], "approximate",
                                       precision="f32", nargout=2)
    assert default_handle().last_path == 2, name
    S0, Y0 = _oracle(t, Imax=60)
    assert _rel(S1, S0) < TOL["f32"]["S"] and _rel(Y1, Y0) < TOL["f32"]["S"], (name, _rel(S1, S0), _rel(Y1, Y0))


def test_host_passes_ping_pong():
    """HOST buffers split into several internal passes (staging sets alternate, the structure check runs per pass)."""
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import Handle
    shape = fx.Shape(Nt=64, Nr=16, L=2, Mr=4, T=2)
    trials = [fx.make_trial(shape, 5.0, 300 + k) for k in range(5)]
    st = lambda k: np.stack([t[k] for t in trials])
    h = Handle(0)
    h.set_chunk(2)                                                        # 3 passes: 2 + 2 + 1 trials
    S1 = jb.proposed_algorithm_psi(st("subY"), st("Omega"), st("A"), trials[0]["Dt"], st("Psi_bar"), 30, [t["tau_Y"] for t in trials],
                                   [t["tau_Z"] for t in trials], [t["rho"] for t in trials], "approximate", precision="f32", nargout=1, handle=h)
    assert h.last_path == 2
    for k, t in enumerate(trials):
        S0, _ = _oracle(t, Imax=30)
        assert _rel(S1[k], S0) < TOL["f32"]["S"], (k, _rel(S1[k], S0))
    h.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_pilots_entry_equals_psi_entry(precision):
    """jstsp_proposed_algorithm_pilots expands Psi_bar(k,:,l) = row l of toeplitz(s_k) (proposed_hbf.m:15-18) on the device:
    same result as handing over Psi_bar, at the metric shape (tensor-core path for f32) and at the default shape."""
    import jstsp19_b200 as jb
    from jstsp19_b200._lib import default_handle
    for shape, imax in ((fx.METRIC, 40), (fx.CONFIG0, 60)):
        t = fx.make_trial(shape, 5.0, 91)
        a = (t["subY"], t["Omega"], t["A"], t["Dt"])
        b = (imax, t["tau_Y"], t["tau_Z"], t["rho"], "approximate")
        S0, Y0 = jb.proposed_algorithm_psi(*a, t["Psi_bar"], *b, precision=precision, nargout=2)
        path0 = default_handle().last_path
        S1, Y1 = jb.proposed_algorithm_pilots(*a, t["pilots"], shape.L, *b, precision=precision, nargout=2)
        assert default_handle().last_path == path0
        assert np.array_equal(S1, S0) and np.array_equal(Y1, Y0)            # identical kernels on identical device inputs
    Sa = jb.proposed_algorithm_pilots(*a, t["pilots"], shape.L, *b, t["indx_S"], precision=precision, nargout=1)
    Sb = jb.proposed_algorithm_psi(*a, t["Psi_bar"], *b, t["indx_S"], precision=precision, nargout=1)
    assert np.array_equal(Sa, Sb)


def test_pilots_entry_batched_host_passes():
    import jstsp19_b200 as jb
    trials = [fx.make_trial(fx.METRIC, 5.0, 200 + k) for k in range(3)]
    st = lambda k: np.stack([t[k] for t in trials])
    args = (st("subY"), st("Omega"), trials[0]["A"], trials[0]["Dt"])
    par = (30, [t["tau_Y"] for t in trials], [t["tau_Z"] for t in trials], [t["rho"] for t in trials], "approximate")
    S0 = jb.proposed_algorithm_psi(*args, st("Psi_bar"), *par, precision="f32", nargout=1)
    S1 = jb.proposed_algorithm_pilots(*args, st("pilots"), 4, *par, precision="f32", nargout=1)
    assert np.array_equal(S1, S0)


def test_persistent_kernel_is_repeatable_bit_for_bit():
    """592 device-resident trials (four per SM, two interleaved per CTA) x 100 iterations, three runs: identical bits.  The persistent kernel
    hands shared memory between the generic and the async proxy in both directions (TMA refills of ring slots the workers just read, MMA reads
    of operands the workers just wrote); a missing proxy fence showed up exactly here, as rare run-to-run differences that no tolerance test saw."""
    import torch
    from jstsp19_b200 import synth
    from jstsp19_b200.engine import AdmmEngine
    nb = 592
    dev = torch.device("cuda", 0)
    snr = torch.tensor([-15.0 + 3 * (k % 11) for k in range(nb)], dtype=torch.float64)
    data = synth.make_batch(synth.METRIC, nb, snr, seed=31, device=dev)
    eng = AdmmEngine(0, "f32")
    runs = []
    for _ in range(3):
        S = eng.proposed_algorithm_psi(data["subY"], data["Omega"], data["A"], data["Dt"], data["Psi"], 100, data["tau_Y"], data["tau_Z"], data["rho"], "approximate")
        runs.append(S.clone())
    assert eng.h.last_path == 2 and eng.h.last_variant == 1
    assert bool(torch.isfinite(torch.view_as_real(runs[0])).all())
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
