"""Synthetic Monte-Carlo inputs on the device (torch = plumbing, not the product path).

Builds, for a batch of independent trials, exactly what the reference's
Monte-Carlo loop feeds the estimator (plot_errorVSsnr.m:56-67,124-136):
channel by the law of wideband_mmwave_channel.m (quirks included, SURVEY.md
section 8a-a1), 4-QAM Hermitian-Toeplitz pilots, ZC combiner, noise, random
spatial-sampling mask with exactly ``Mr`` ones per column, dictionaries
``A = W_e' Dr`` and ``B_l = Dt' Psi_l`` and the driver-side parameters
``tau_Y, tau_Z, rho``.  ``build_from_draws`` takes the raw random draws so the
CPU tests can check it against the oracle on identical draws; ``make_batch``
draws them from torch's counter-based (Philox) generator keyed by (seed, global block of 64 trials),
so a trial's inputs do not depend on how the trials are sharded.

Memory layout of every returned matrix: ``(batch, cols, rows)`` C-contiguous,
i.e. per-trial COLUMN-MAJOR like the C ABI expects.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class Shape:
    Nt: int
    Nr: int
    L: int
    Mr: int
    T: int
    ncl: int = 2
    nray: int = 3

    @property
    def M(self):
        return self.T * self.Nt

    @property
    def P(self):
        return self.L * self.Nt

    @property
    def Np(self):
        return self.ncl * self.nray


METRIC = Shape(Nt=64, Nr=16, L=4, Mr=4, T=16)      # BASELINE.json configs[1]
CONFIG0 = Shape(Nt=4, Nr=32, L=4, Mr=4, T=35)      # plot_errorVSsnr.m defaults


def _dft(Mant, G, device, cdtype):
    m = torch.arange(Mant, device=device, dtype=torch.float64)[:, None]
    g = torch.arange(G, device=device, dtype=torch.float64)[None, :]
    return (torch.exp(-1j * m * 2.0 * math.pi * g / G) / math.sqrt(Mant)).to(cdtype)


def _zc(N, device, cdtype):
    n = torch.arange(N, device=device, dtype=torch.float64)[:, None]
    k = torch.arange(1, N + 1, device=device, dtype=torch.float64)[None, :]
    return (torch.exp(-1j * 11.0 * n * math.pi * k / N) / math.sqrt(N)).to(cdtype)   # createBeamformer.m:15-17


def _laplacian_angle(u):
    beta = 1.0 / (1.0 - math.exp(-math.sqrt(2.0) * math.pi / 50.0))
    return beta * (math.exp(-math.sqrt(2.0) / 50.0 * math.pi) - torch.cosh(u))       # wideband_mmwave_channel.m:56-62


def _steer(phi, Mant):
    m = torch.arange(Mant, device=phi.device, dtype=phi.dtype)
    return torch.exp(-1j * math.pi * torch.sin(-phi)[..., None] * m)                # :42-52


def build_from_draws(s: Shape, coef, u_r, u_t, noise_unit, sym_idx, mask_rank, sigma2, cdtype=torch.complex128):
    """coef (b,L,Np) complex CN(0,1); u_r,u_t (b,L,Np) uniforms; noise_unit (b,Nr,M) CN(0,1);
    sym_idx (b,Nt,M) ints in 0..3; mask_rank (b,Nr,M): per column, the rows whose rank is < Mr
    are sampled (a uniformly random Mr-subset); sigma2 (b,) noise variance."""
    dev = coef.device
    rd = torch.float64 if cdtype == torch.complex128 else torch.float32
    b = coef.shape[0]
    Nr, Nt, L, M = s.Nr, s.Nt, s.L, s.M
    # --- channel (wideband_mmwave_channel.m:12-38) ---
    ar = _steer(_laplacian_angle(u_r[:, 0, :].double()), Nr)           # page-1 quirk: tap 1 vectors for all taps
    at = _steer(_laplacian_angle(u_t[:, 0, :].double()), Nt)
    w = torch.ones(s.Np, dtype=torch.float64, device=dev)
    for c in range(s.ncl):                                             # Hl keeps accumulating inside the cluster loop (:29)
        w[c * s.nray:(c + 1) * s.nray] = s.ncl - c
    cw = coef.to(torch.complex128) * w / math.sqrt(s.Np)               # (b,L,Np)
    H = torch.einsum("blp,bpr,bpt->blrt", cw, ar, at.conj())           # (b,L,Nr,Nt)
    Dr = _dft(Nr, Nr, dev, torch.complex128)
    Dt = _dft(Nt, Nt, dev, torch.complex128)
    Z = torch.einsum("gr,blrt,th->blgh", Dr.conj().T, H, Dt)           # Dr' H_l Dt
    Zbar = Z.permute(0, 2, 1, 3).reshape(b, Nr, L * Nt)                # reshape(Z, Gr, L*Gt): column l*Gt+j
    # --- pilots: Psi_bar(k,:,l) = row l of toeplitz(s_k) (proposed_hbf.m:15-18) ---
    qam = torch.tensor([1 + 1j, -1 + 1j, 1 - 1j, -1 - 1j], dtype=torch.complex128, device=dev) / math.sqrt(2.0)
    sk = qam[sym_idx]                                                  # (b,Nt,M)
    j = torch.arange(M, device=dev)
    Psi = torch.empty(b, L, Nt, M, dtype=torch.complex128, device=dev)
    for l in range(L):
        d = j - l
        row = sk[:, :, d.abs()]
        Psi[:, l] = torch.where(d >= 0, row, row.conj())
    # --- received signal, combiner, mask (proposed_hbf.m:14-42) ---
    R = torch.einsum("blrt,bltm->brm", H, Psi) + noise_unit.to(torch.complex128) * torch.sqrt(sigma2.double())[:, None, None]
    W = _zc(Nr, dev, torch.complex128)
    WR = torch.einsum("qr,brm->bqm", W.conj().T, R)
    Omega = (mask_rank < s.Mr).to(torch.float64)
    subY = Omega * WR
    A = W.conj().T @ Dr                                                # plot_errorVSsnr.m:132
    B = torch.einsum("gt,bltm->blgm", Dt.conj().T, Psi).reshape(b, L * Nt, M)   # :133-136
    # --- driver-side parameters (:127-130) ---
    fy = (subY.abs() ** 2).sum(dim=(1, 2))
    tau_Y = 1.0 / fy
    tau_Z = 0.5 / (Zbar.abs() ** 2).sum(dim=(1, 2))
    ev = torch.linalg.eigvalsh(subY @ subY.conj().transpose(1, 2))     # ascending; eigs() -> 6 largest
    lam6 = ev[:, -6] if Nr >= 6 else torch.zeros_like(ev[:, 0])
    rho = torch.sqrt(torch.clamp(lam6, min=0.0) / fy)
    cm = lambda x: x.transpose(1, 2).contiguous().to(cdtype)           # per-trial column-major
    return dict(subY=cm(subY), Omega=Omega.transpose(1, 2).contiguous().to(rd), A=A.T.contiguous().to(cdtype)[None],
                B=cm(B), Zbar=cm(Zbar), tau_Y=tau_Y.double().contiguous(), tau_Z=tau_Z.double().contiguous(),
                rho=rho.double().contiguous(), H=H,
                # factors of B as the drivers hold them: Dt (1,Gt,Nt) and Psi_bar (b,L,M,Nt), per-trial column-major
                Dt=Dt.T.contiguous().to(cdtype)[None], Psi=Psi.transpose(2, 3).contiguous().to(cdtype),
                # the pilot sequences themselves (b,M,Nt): row k of the Nt x M matrix is s_k (plot_errorVSsnr.m:63-67)
                pilots=sk.transpose(1, 2).contiguous().to(cdtype))


DRAW_BLOCK = 64      # trials per generator key


def _draw_block(s: Shape, seed, block, device):
    """The raw numbers of the DRAW_BLOCK trials of global block `block`: one Philox stream per (seed, block)."""
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 1000003 + int(block)) & 0x7FFFFFFFFFFFFFFF)
    n = DRAW_BLOCK
    f = dict(device=device, dtype=torch.float64, generator=g)
    coef = torch.randn(n, 2, s.L, s.Np, **f)
    u_r = torch.rand(n, s.L, s.Np, **f)
    u_t = torch.rand(n, s.L, s.Np, **f)
    noise = torch.randn(n, 2, s.Nr, s.M, **f)
    sym = torch.randint(0, 4, (n, s.Nt, s.M), device=device, generator=g)
    order = torch.rand(n, s.Nr, s.M, **f)
    return coef, u_r, u_t, noise, sym, order


def draw(s: Shape, batch, snr_db, seed, first_trial=0, device="cuda"):
    """Raw draws of the global trials [first_trial, first_trial+batch): coefficient, angle uniforms, unit noise, 4-QAM symbol indices, mask
    ranks and the noise variance.  Trials are generated in blocks of DRAW_BLOCK keyed by (seed, global block index), so trial t gets the same
    numbers whichever shard, batch size or GPU count it is computed under (results independent of the partition)."""
    lo, hi = int(first_trial), int(first_trial) + int(batch)
    parts = []
    for blk in range(lo // DRAW_BLOCK, (hi + DRAW_BLOCK - 1) // DRAW_BLOCK):
        a, b = max(lo, blk * DRAW_BLOCK) - blk * DRAW_BLOCK, min(hi, (blk + 1) * DRAW_BLOCK) - blk * DRAW_BLOCK
        parts.append(tuple(t[a:b] for t in _draw_block(s, seed, blk, device)))
    coef, u_r, u_t, noise, sym, order = (torch.cat(ts, dim=0) if len(ts) > 1 else ts[0].contiguous() for ts in zip(*parts))
    coef = torch.complex(coef[:, 0], coef[:, 1]) / math.sqrt(2.0)
    noise = torch.complex(noise[:, 0], noise[:, 1]) / math.sqrt(2.0)
    rank = order.argsort(dim=1).argsort(dim=1)            # rank of each row within its column: the Mr smallest are sampled (randperm, proposed_hbf.m:37-41)
    snr = torch.as_tensor(snr_db, device=device, dtype=torch.float64).expand(batch) if not torch.is_tensor(snr_db) else snr_db.to(device).double()
    sigma2 = 10.0 ** (-snr / 10.0)
    return coef, u_r, u_t, noise, sym, rank, sigma2


def make_batch(s: Shape, batch, snr_db, seed, first_trial=0, device="cuda", cdtype=torch.complex64):
    """:func:`draw` + :func:`build_from_draws` (torch fp64 arithmetic: the bench's input generator and the pipeline's cross-check)."""
    return build_from_draws(s, *draw(s, batch, snr_db, seed, first_trial, device), cdtype=cdtype)


def nmse_spectral(S, Zbar):
    """norm(S-Zbar)^2/norm(Zbar)^2 with matrix 2-norms, clipped at 1 (plot_errorVSsnr.m:138-141).
    Both in (batch, cols, rows) layout - the spectral norm is transpose invariant."""
    e = torch.linalg.matrix_norm((S - Zbar).to(torch.complex128), ord=2) ** 2 / torch.linalg.matrix_norm(Zbar.to(torch.complex128), ord=2) ** 2
    return torch.clamp(e, max=1.0)
