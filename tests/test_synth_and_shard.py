"""CPU tests of the host logic: the torch input synthesiser against the oracle on identical
draws, trial sharding, and the world_size-2 gloo reduction."""
import os

import numpy as np
import pytest
import torch

from oracle import estimators as est
from oracle import fixtures as fx
from oracle import matlab_compat as mc
from oracle import system_model as sm


class Replay(mc.RefRandom):
    """Serves pre-recorded draws in the reference's consumption order."""

    def __init__(self, normals, uniforms, perms):
        self.n, self.u, self.p = list(normals), list(uniforms), list(perms)

    def randn(self, *shape):
        if len(shape) == 0 or shape == (1,):
            return float(self.n.pop(0))
        return np.asarray(self.n.pop(0)).reshape(shape, order="F")

    def rand(self, *shape):
        if len(shape) == 0 or shape == (1,):
            return float(self.u.pop(0))
        return np.asarray(self.u.pop(0)).reshape(shape, order="F")

    def randperm(self, n):
        return np.asarray(self.p.pop(0))


def test_synth_matches_oracle_on_identical_draws():
    from jstsp19_b200 import synth
    s = synth.Shape(Nt=3, Nr=8, L=2, Mr=3, T=5)
    so = fx.Shape(Nt=3, Nr=8, L=2, Mr=3, T=5)
    g = torch.Generator().manual_seed(5)
    f = dict(dtype=torch.float64, generator=g)
    b = 2
    coef = torch.complex(torch.randn(b, s.L, s.Np, **f), torch.randn(b, s.L, s.Np, **f)) / np.sqrt(2)
    u_r, u_t = torch.rand(b, s.L, s.Np, **f), torch.rand(b, s.L, s.Np, **f)
    noise = torch.complex(torch.randn(b, s.Nr, s.M, **f), torch.randn(b, s.Nr, s.M, **f)) / np.sqrt(2)
    sym = torch.randint(0, 4, (b, s.Nt, s.M), generator=g)
    rank = torch.rand(b, s.Nr, s.M, **f).argsort(dim=1).argsort(dim=1)
    sigma2 = torch.tensor([0.3, 0.05], dtype=torch.float64)
    d = synth.build_from_draws(s, coef, u_r, u_t, noise, sym, rank, sigma2, cdtype=torch.complex128)
    for k in range(b):
        normals, uniforms, perms = [], [], []
        for l in range(s.L):
            for r in range(s.Np):
                c = coef[k, l, r] * np.sqrt(2)
                normals += [c.real.item(), c.imag.item()]
                uniforms += [u_r[k, l, r].item(), u_t[k, l, r].item()]
        nz = noise[k].numpy() * np.sqrt(2)                      # oracle scales by sqrt(sigma2/2)
        normals += [nz.real.reshape(-1, order="F"), nz.imag.reshape(-1, order="F")]
        for kk in range(s.Nt):
            uniforms.append((sym[k, kk].numpy() + 0.5) / 4.0)  # randsrc: floor(4u) -> symbol index
        for t in range(s.M):
            perms.append(np.argsort(rank[k, :, t].numpy()) + 1)
        # fixtures.make_trial with the replayed stream
        rng = Replay(normals, uniforms, perms)
        H, Zbar, Ar, At, Dr, Dt = sm.wideband_mmwave_channel(so.L, so.Nr, so.Nt, 2, 3, so.Gr, so.Gt, rng)
        N = np.sqrt(sigma2[k].item() / 2.0) * (rng.randn(so.Nr, so.M) + 1j * rng.randn(so.Nr, so.M))
        pilots = np.stack([sm.qam4mod(so.M, rng) for _ in range(so.Nt)])
        Psi_bar = sm.psi_bar_from_pilots(pilots, so.M, so.L)
        W = sm.create_beamformer(so.Nr, "ZC")
        Yh, W_e, _, Omega, _ = sm.proposed_hbf(H, N, None, so.M, so.Nr, so.Mr, W, rng, Psi_bar=Psi_bar)
        tY, tZ, rho = est.admm_parameters(Yh, Zbar)
        A = W_e.conj().T @ Dr
        B = sm.dictionary_B(Dt, Psi_bar)
        cm = lambda x: x[k].numpy().T                               # (cols, rows) storage -> rows x cols
        assert np.allclose(cm(d["Zbar"]), Zbar, atol=1e-12)
        assert np.array_equal(cm(d["Omega"]), Omega)                # mask bit-exact
        assert np.allclose(cm(d["subY"]), Yh, atol=1e-11)
        assert np.allclose(d["A"][0].numpy().T, A, atol=1e-13)
        assert np.allclose(cm(d["B"]), B, atol=1e-12)
        assert d["tau_Y"][k].item() == pytest.approx(tY, rel=1e-11)
        assert d["tau_Z"][k].item() == pytest.approx(tZ, rel=1e-11)
        assert d["rho"][k].item() == pytest.approx(rho, rel=1e-9)


def test_shard_range_partitions_trials():
    from jstsp19_b200.engine import shard_range
    for n, w in [(10, 1), (10, 3), (7, 8), (10000, 8)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jstsp19_b200.engine import MonteCarlo, shard_range
    lo, hi = shard_range(11, rank, world)
    vals = torch.arange(lo, hi, dtype=torch.float64) / 10.0
    if rank == 1:
        vals[0] = float("nan")
    m = MonteCarlo("cpu")
    m.add(vals)
    q.put((rank, m.reduce()))
    dist.destroy_process_group()


def test_monte_carlo_reduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    lo1, _ = 6, 11
    expect = (sum(range(11)) - lo1) / 10.0 / 10
    for r in (0, 1):
        assert res[r]["trials"] == 10 and res[r]["flagged"] == 1
        assert res[r]["mean_nmse"] == pytest.approx(expect)


def test_draws_do_not_depend_on_the_partition():
    """trial t gets the same random numbers whatever shard / batch / GPU count it is computed under
    (blocks of synth.DRAW_BLOCK trials keyed by (seed, global block index))."""
    from jstsp19_b200 import synth
    s = synth.Shape(Nt=4, Nr=8, L=2, Mr=2, T=3)
    whole = synth.draw(s, 100, 5.0, 7, 0, "cpu")
    for lo, hi in [(0, 100 // 3), (100 // 3, 67), (67, 100), (31, 64), (32, 33)]:
        part = synth.draw(s, hi - lo, 5.0, 7, lo, "cpu")
        assert all(torch.equal(w[lo:hi], p) for w, p in zip(whole, part))
    other = synth.draw(s, 8, 5.0, 8, 0, "cpu")
    assert not torch.equal(other[0], whole[0][:8])
