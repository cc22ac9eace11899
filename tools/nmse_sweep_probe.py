"""NMSE statistics of the whole trial loop (engine.TrialPipeline) at the metric shape along the SNR sweep, Imax = 100 and 400:
mean / median / quantiles of the spectral-norm NMSE (plot_errorVSsnr.m:138-141) and of the Frobenius error, 910 trials per point.
    python tools/nmse_sweep_probe.py        (needs a GPU; numbers quoted in DESIGN.md section 6)"""
import torch, sys
sys.path.insert(0, '.')
from jstsp19_b200 import synth
from jstsp19_b200.engine import TrialPipeline
shape, per = synth.METRIC, 910
pipe = TrialPipeline(shape, 0, "f32")
for imax in (100, 400):
    for i, snr in enumerate(range(-15, 16, 6)):
        draws = synth.draw(shape, per, float(snr), seed=2019, first_trial=i * per, device="cuda")
        out = pipe.run_from_draws(*draws, imax=imax, keep=True)
        nm = out["nmse"].double()
        S, Z = out["S"], out["Zbar"]
        fro = ((S - Z).abs() ** 2).sum((1, 2)) / (Z.abs() ** 2).sum((1, 2))
        print(imax, snr, "mean %.4f median %.4f clipped %.3f  p10 %.4f p90 %.4f | fro mean %.4f median %.4f" % (nm.mean(), nm.median(), (nm >= 1).double().mean(), nm.quantile(0.1), nm.quantile(0.9), fro.double().mean(), fro.double().median()))
