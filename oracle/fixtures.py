"""Seeded per-trial inputs for parity tests and the CPU baseline (test
infrastructure - see ``oracle/__init__.py``).

One trial follows the body of the reference's Monte-Carlo loop
(plot_errorVSsnr.m:56-67,124-136,143): channel, noise, 4-QAM Toeplitz pilots,
ZC (or fft) combiner, random spatial-sampling mask, dictionaries A and B and the
driver-side parameters tau_Y, tau_Z, rho.  All randomness comes from a
``RefRandom(seed)`` in the reference's consumption order, so oracle and engine
are fed identical arrays.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import estimators as est
from . import system_model as sm
from .matlab_compat import RefRandom, sort_descend_idx, vec


@dataclass
class Shape:
    Nt: int
    Nr: int
    L: int
    Mr: int          # RF chains active per training instant ("NRF")
    T: int           # training frames ("K"); measurement columns M = T*Nt
    Gr: int | None = None
    Gt: int | None = None
    Mr_e: int | None = None
    ncl: int = 2
    nray: int = 3
    combiner: str = "ZC"

    def __post_init__(self):
        self.Gr = self.Gr or self.Nr
        self.Gt = self.Gt or self.Nt
        self.Mr_e = self.Mr_e or self.Nr

    @property
    def M(self):
        return self.T * self.Nt

    @property
    def P(self):
        return self.L * self.Gt


# BASELINE.json configs
CONFIG0 = Shape(Nt=4, Nr=32, L=4, Mr=4, T=35)                  # plot_errorVSsnr.m defaults
METRIC = Shape(Nt=64, Nr=16, L=4, Mr=4, T=16)                  # headline metric shape
TINY = Shape(Nt=2, Nr=8, L=2, Mr=2, T=6)                       # unit-test scale


def make_trial(shape: Shape, snr_db: float, seed: int, rho_rule: str = "sigma6"):
    """Draw one Monte-Carlo trial (plot_errorVSsnr.m:56-67,124-136,143)."""
    rng = RefRandom(seed)
    s = shape
    sigma2 = 10.0 ** (-snr_db / 10.0)                             # :49
    H, Zbar, Ar, At, Dr, Dt = sm.wideband_mmwave_channel(s.L, s.Nr, s.Nt, s.ncl, s.nray, s.Gr, s.Gt, rng)  # :57
    M = s.M
    re = rng.randn(s.Nr, M)
    im = rng.randn(s.Nr, M)
    N = np.sqrt(sigma2 / 2.0) * (re + 1j * im)                    # :60
    pilots = np.zeros((s.Nt, M), dtype=np.complex128)
    for k in range(s.Nt):
        pilots[k, :] = sm.qam4mod(M, rng)                         # :63-67
    Psi_bar = sm.psi_bar_from_pilots(pilots, M, s.L)
    W = sm.create_beamformer(s.Nr, s.combiner)                    # :124
    Y_hbf, W_e, Psi_bar, Omega, Ynl = sm.proposed_hbf(H, N, None, M, s.Mr_e, s.Mr, W, rng, Psi_bar=Psi_bar)  # :125
    tau_Y, tau_Z, rho = est.admm_parameters(Y_hbf, Zbar, rho_rule)  # :127-130
    A = W_e.conj().T @ Dr                                         # :132
    B = sm.dictionary_B(Dt, Psi_bar)                              # :133-136
    indx_S = sort_descend_idx(np.abs(vec(Zbar)))                  # :143
    return dict(H=H, Zbar=Zbar, Dr=Dr, Dt=Dt, N=N, pilots=pilots, Psi_bar=Psi_bar, W=W, W_e=W_e,
                Omega=Omega, subY=Y_hbf, Ynoiseless=Ynl, A=A, B=B, tau_Y=tau_Y, tau_Z=tau_Z, rho=rho,
                indx_S=indx_S, sigma2=sigma2, shape=s)


def conventional_problem(trial, numOfnz=None):
    """The conventional-HBF linear system of plot_errorVSsnr.m:73-80 used by
    VAMP / OMP: Phi = kron((B*B').', A), y = vec(Y*B')."""
    s = trial["shape"]
    from .matlab_compat import mround
    T_hbf = mround(s.T / (s.Nr / s.Mr)) * s.Nt                    # :22
    Psi_bar = trial["Psi_bar"][:, :T_hbf, :]
    # hbf(H, N(:,1:T_hbf), Psi_i(1:T_hbf,1:T_hbf,:), ...) :73 - leading principal
    # block of a Toeplitz matrix is the Toeplitz of the truncated sequence.
    Y, W_c, _, _ = sm.hbf(trial["H"], trial["N"][:, :T_hbf], None, T_hbf, s.Nr, trial["W"], Psi_bar=Psi_bar)
    A = W_c.conj().T @ trial["Dr"]                                # :74
    B = sm.dictionary_B(trial["Dt"], Psi_bar)                     # :75-78
    Phi = np.kron((B @ B.conj().T).T, A)                          # :79
    y = vec(Y @ B.conj().T)                                       # :80
    return dict(Y=Y, A=A, B=B, Phi=Phi, y=y, T_hbf=T_hbf)
