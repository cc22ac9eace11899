"""Device time of each stage of the trial-loop body (engine.TrialPipeline) at the metric shape: where the gap between the pipeline rate and
the solver rate goes.  Developer tool, needs a B200.   usage: python tools/pipeline_stages.py [trials]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jstsp19_b200 import synth  # noqa: E402
from jstsp19_b200.engine import TrialPipeline  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 888
pipe = TrialPipeline(synth.METRIC, 0, "f32")
snr = torch.tensor([-15.0 + 3 * (k % 11) for k in range(nb)], dtype=torch.float64)
h = pipe.eng.h
for rep in range(2):
    h.profile(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.run(nb, snr, seed=5 + rep, first_trial=0)
    e1.record(); torch.cuda.synchronize()
    pr = h.profile_read(); h.profile(0)
print(f"{nb} trials: {e0.elapsed_time(e1):.2f} ms per step on the device")
for k, v in pr.items():
    if v[1]:
        print(f"  {k:14s} {v[0]:8.3f} ms over {v[1]:4d} launches")
print(f"  kernels total  {sum(v[0] for v in pr.values()):8.3f} ms")
