"""jstsp_draw_trials (csrc/rng.cu): the draws of the Monte-Carlo trial loop on the device.  The reference draws from MATLAB's global stream
without a seed (wideband_mmwave_channel.m:19-22, plot_errorVSsnr.m:60,63-67, qam4mod.m:7-8, proposed_hbf.m:37), so what is checked is
  * every value against a host restatement of the same Philox4x32-10 streams (bit-exact integers, 1e-13 for the fp64 Box-Muller values),
  * independence of the partition: a trial's numbers do not depend on the batch it is drawn in,
  * the distributions: N(0,1), U(0,1), balanced 4-QAM, noise power sigma2, uniform random permutations with Mr-subset sampling."""
import math

import numpy as np
import pytest
import torch

from test_rng import philox_py

pytestmark = pytest.mark.gpu

SEED = (7 << 32) | 20190913


def _pipe(shape, precision="f32"):
    from jstsp19_b200.engine import TrialPipeline
    return TrialPipeline(shape, 0, precision)


def _blk(t, stream, idx):
    return philox_py([idx, stream, t & 0xFFFFFFFF, t >> 32], [SEED & 0xFFFFFFFF, SEED >> 32])


def test_draws_match_host_restatement():
    from jstsp19_b200 import synth
    s = synth.Shape(Nt=8, Nr=16, L=3, Mr=4, T=5)
    pipe = _pipe(s, "f64")
    first, b = 41, 3
    normals, uniforms, pilots, noise, perm, sigma2 = (x.cpu().numpy() for x in pipe.device_draws(b, torch.tensor([5.0, -3.0, 10.0]), SEED, first))
    Np, M = s.ncl * s.nray, s.M
    u01 = lambda x: (x + 0.5) / 4294967296.0
    for k in range(b):
        t = first + k
        for j in range(s.L * Np):
            r = _blk(t, 0, j)
            rad = math.sqrt(-2.0 * math.log(u01(r[0])))
            want = [rad * math.cos(2 * math.pi * u01(r[1])), rad * math.sin(2 * math.pi * u01(r[1]))]
            assert np.allclose(normals[k].reshape(-1, 2)[j], want, rtol=0, atol=1e-13)
            r = _blk(t, 1, j)
            assert np.array_equal(uniforms[k].reshape(-1, 2)[j], [u01(r[0]), u01(r[1])])
        flat = pilots[k].reshape(-1)
        for i in range((M * s.Nt + 63) // 64):
            r = _blk(t, 2, i)
            for e in range(64 * i, min(64 * i + 64, flat.size)):
                sym = (r[(e % 64) // 16] >> (2 * (e % 16))) & 3
                a = math.sqrt(0.5)
                assert flat[e] == complex(-a if sym & 1 else a, -a if sym & 2 else a)
        nflat = noise[k].reshape(-1)
        for i in range(0, (nflat.size + 1) // 2, 7):
            r = _blk(t, 3, i)
            for w in range(2):
                rad = math.sqrt(-2.0 * math.log(u01(r[2 * w])))
                want = math.sqrt(sigma2[k] / 2) * complex(rad * math.cos(2 * math.pi * u01(r[2 * w + 1])), rad * math.sin(2 * math.pi * u01(r[2 * w + 1])))
                assert abs(nflat[2 * i + w] - want) < 1e-12
        cpc = (s.Nr + 3) // 4
        for m in range(M):
            keys = []
            for j in range(cpc):
                keys += _blk(t, 4, m * cpc + j)
            order = sorted(range(s.Nr), key=lambda n: (keys[n], n))
            assert perm[k, m].tolist() == [n + 1 for n in order]


def test_draws_are_partition_invariant():
    from jstsp19_b200 import synth
    pipe = _pipe(synth.METRIC)
    snr = torch.tensor([-15.0 + 3 * (k % 11) for k in range(16)])
    whole = pipe.device_draws(16, snr, SEED, 100)
    part = pipe.device_draws(5, snr[8:13], SEED, 108)
    for a, b in zip(whole, part):
        assert torch.equal(a[8:13], b)
    other = pipe.device_draws(5, snr[8:13], SEED + 1, 108)
    assert not torch.equal(other[2], part[2])


def test_draw_distributions():
    from jstsp19_b200 import synth
    s = synth.METRIC
    pipe = _pipe(s)
    b = 64
    snr = torch.tensor([-15.0 + 3 * (k % 11) for k in range(b)])
    normals, uniforms, pilots, noise, perm, sigma2 = pipe.device_draws(b, snr, SEED, 0)
    n = normals.double().flatten()
    assert abs(float(n.mean())) < 4 / math.sqrt(n.numel()) and abs(float(n.var()) - 1) < 0.1
    u = uniforms.flatten()
    assert float(u.min()) > 0 and float(u.max()) < 1 and abs(float(u.mean()) - 0.5) < 4 / math.sqrt(12 * u.numel())
    # pilots: unit modulus, the four symbols equally likely
    assert torch.allclose(pilots.abs(), torch.ones_like(pilots.abs()), atol=1e-6)
    q = ((pilots.real < 0).long() + 2 * (pilots.imag < 0).long()).flatten()
    cnt = torch.bincount(q, minlength=4).double()
    assert float((cnt / q.numel() - 0.25).abs().max()) < 4 * math.sqrt(0.25 * 0.75 / q.numel())
    # noise: complex normal of variance sigma2 per trial (plot_errorVSsnr.m:60), real and imaginary parts uncorrelated
    nv = (noise.abs() ** 2).double().mean(dim=(1, 2))
    per = noise[0].numel()
    assert torch.allclose(nv, sigma2, rtol=6 / math.sqrt(per))
    z = noise.double() if False else noise
    assert abs(float((z.real * z.imag).double().mean() / nv.mean())) < 5 / math.sqrt(noise.numel())
    # sampling order: every column a permutation; each row is among the first Mr with probability Mr / Nr
    srt = perm.sort(dim=2).values
    assert torch.equal(srt, torch.arange(1, s.Nr + 1, dtype=torch.int32, device=perm.device).expand_as(srt))
    first = torch.bincount((perm[:, :, :s.Mr] - 1).flatten().long(), minlength=s.Nr).double()
    ncol = perm.shape[0] * perm.shape[1]
    pexp = s.Mr / s.Nr
    assert float((first / ncol - pexp).abs().max()) < 5 * math.sqrt(pexp * (1 - pexp) / ncol)
