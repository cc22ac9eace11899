"""GPU parity of the whole trial-loop body (plot_errorVSsnr.m:56-67,124-141) run through the library (engine.TrialPipeline):
every intermediate against the fp64 torch restatement on the same draws, the estimate against the oracle."""
import numpy as np
import pytest
import torch

from oracle import estimators as est

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name,precision,tol", [("metric", "f32", 3e-5), ("small", "f64", 1e-8)])
def test_trial_pipeline(name, precision, tol):
    from jstsp19_b200 import synth
    from jstsp19_b200.engine import TrialPipeline
    shape = synth.METRIC if name == "metric" else synth.Shape(Nt=4, Nr=32, L=4, Mr=4, T=10)
    imax = 30
    pipe = TrialPipeline(shape, 0, precision)
    draws = synth.draw(shape, 3, torch.tensor([5.0, -4.0, 12.0]), seed=77, first_trial=5, device="cuda")
    out = pipe.run_from_draws(*draws, imax=imax, keep=True)
    ref = synth.build_from_draws(shape, *draws, cdtype=torch.complex128)
    c = lambda t: t.cpu().numpy()
    assert np.array_equal(c(out["Omega"]).astype(np.float64), c(ref["Omega"]))                     # the sampling mask is bit-exact
    assert (c(ref["Omega"]).sum(axis=2) == shape.Mr).all()                                         # Mr ones per column (proposed_hbf.m:36-41)
    etol = 3e-6 if precision == "f32" else 1e-11
    assert _rel(c(out["Zbar"]), c(ref["Zbar"])) < etol and _rel(c(out["subY"]), c(ref["subY"])) < etol
    for k in ("tau_Y", "tau_Z", "rho"):
        np.testing.assert_allclose(c(out[k]), c(ref[k]), rtol=2e-5 if precision == "f32" else 1e-9)
    # the estimate of trial 0 against the oracle on the fp64 inputs, and the NMSE against the oracle's formula
    T = lambda t: np.swapaxes(c(t), -1, -2)
    subY, Om, A, B, Zb = T(ref["subY"])[0], T(ref["Omega"])[0], T(ref["A"])[0], T(ref["B"])[0], T(ref["Zbar"])[0]
    S0, _, _ = est.proposed_algorithm_structured(subY, Om, A, B, imax, float(ref["tau_Y"][0]), float(ref["tau_Z"][0]), float(ref["rho"][0]), "approximate", want_conv=False)
    S1 = T(out["S"])[0]
    assert _rel(S1, S0) < tol, _rel(S1, S0)
    assert abs(float(out["nmse"][0]) - est.nmse(S1.astype(np.complex128), Zb)) < 1e-5 * max(est.nmse(S1.astype(np.complex128), Zb), 1e-3)


def test_pipeline_is_gpu_count_invariant():
    """Trials are keyed by (seed, first trial): a shard gives the same NMSEs whatever else runs beside it."""
    from jstsp19_b200 import synth
    from jstsp19_b200.engine import TrialPipeline
    shape = synth.Shape(Nt=64, Nr=16, L=2, Mr=4, T=2)
    pipe = TrialPipeline(shape, 0, "f32")
    a = pipe.run(4, 5.0, seed=3, first_trial=8, imax=10)
    b = pipe.run(4, 5.0, seed=3, first_trial=8, imax=10)
    assert torch.equal(a, b) and bool(torch.isfinite(a).all())
