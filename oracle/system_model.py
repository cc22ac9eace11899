"""Oracle restatement of the reference's system / measurement model (test
infrastructure - see ``oracle/__init__.py``).  fp64, column-major semantics.

Files followed (relative to the reference root):
  basic_system_functions/wideband_mmwave_channel.m
  basic_system_functions/wideband_hybBF_comm_system_training.m
  basic_system_functions/proposed_hbf.m, hbf.m, createBeamformer.m, qam4mod.m
"""
from __future__ import annotations

import math

import numpy as np

from .matlab_compat import RefRandom, mround, toeplitz_hermitian, toeplitz_hermitian_rows


# ----------------------------------------------------------------------------
# wideband_mmwave_channel.m
# ----------------------------------------------------------------------------
def dft_dictionary(Mant, G):
    """``1/sqrt(M)*exp(-1j*(0:M-1)'*2*pi*(0:G-1)/G)`` (wideband_mmwave_channel.m:9-10)."""
    m = np.arange(Mant)[:, None]
    g = np.arange(G)[None, :]
    return np.exp(-1j * m * 2.0 * math.pi * g / G) / math.sqrt(Mant)


def steering(phi, Mant):
    """Local function ``angle`` (wideband_mmwave_channel.m:42-52): un-normalised
    ULA response exp(-1j*pi*sin(0-phi)*(0:M-1)')."""
    ghz = 90.0
    wavelength = 30.0 / ghz
    spacing = 0.5 * wavelength
    wavenumber = 2.0 * math.pi / wavelength
    phase_shift = wavenumber * spacing * math.sin(0.0 - phi) * np.arange(Mant)
    return np.exp(-1j * phase_shift)


def laplacian_angle(u):
    """Local function ``genLaplacianSamples`` (wideband_mmwave_channel.m:56-62)
    evaluated at the uniform draw ``u``."""
    sigma_phi = 50.0
    beta = 1.0 / (1.0 - math.exp(-math.sqrt(2.0) * math.pi / sigma_phi))
    return beta * (math.exp(-math.sqrt(2.0) / sigma_phi * math.pi) - math.cosh(u))


def wideband_mmwave_channel(L, Mr, Mt, ncl, nray, Gr, Gt, rng: RefRandom):
    """[H,Zbar,Ar,At,Dr,Dt] = wideband_mmwave_channel(...) (wideband_mmwave_channel.m:1-39).

    Reproduced quirks: ``Ar(:,index)`` with two subscripts reads page 1 (:24-25);
    ``H(:,:,l) += Hl`` sits inside the cluster loop while ``Hl`` keeps
    accumulating (:29); RNG order per ray randn, randn, rand, rand (:19-22).
    """
    Np = ncl * nray
    H = np.zeros((Mr, Mt, L), dtype=np.complex128)
    Z = np.zeros((Gr, Gt, L), dtype=np.complex128)
    Ar = np.zeros((Mr, Np, L), dtype=np.complex128)
    At = np.zeros((Mt, Np, L), dtype=np.complex128)
    Dr = dft_dictionary(Mr, Gr)
    Dt = dft_dictionary(Mt, Gt)
    for l in range(L):
        Hl = np.zeros((Mr, Mt), dtype=np.complex128)
        index = 0
        for _tap in range(ncl):
            for _ray in range(nray):
                re = rng.randn()
                im = rng.randn()
                coeff = (re + 1j * im) / math.sqrt(2.0)           # :19
                phi_r = laplacian_angle(rng.rand())               # :20
                Ar[:, index, l] = steering(phi_r, Mr)             # :21
                At[:, index, l] = steering(laplacian_angle(rng.rand()), Mt)  # :22
                # :24 - two-subscript indexing of a 3-D array => page 1
                Hl = Hl + coeff * np.outer(Ar[:, index, 0], np.conj(At[:, index, 0]))
                index += 1
            H[:, :, l] = H[:, :, l] + Hl                          # :29 (inside the cluster loop)
        H[:, :, l] = H[:, :, l] / math.sqrt(nray * ncl)           # :33
        Z[:, :, l] = Dr.conj().T @ H[:, :, l] @ Dt                # :35
    Zbar = Z.reshape((Gr, L * Gt), order="F")                     # :38
    return H, Zbar, Ar, At, Dr, Dt


# ----------------------------------------------------------------------------
# createBeamformer.m / qam4mod.m
# ----------------------------------------------------------------------------
def create_beamformer(N, kind, rng: RefRandom | None = None):
    """createBeamformer.m:1-35."""
    n = np.arange(N)[:, None]
    if kind == "fft":
        return np.fft.fft(np.eye(N), axis=0) / math.sqrt(N)       # :6
    if kind == "rand":
        return rng.randsrc(N, N, np.array([1, -1, 1j, -1j])) / math.sqrt(N)  # :8
    if kind == "rand_ps":
        Gr = 32
        return np.exp(-1j * n * 2 * math.pi * rng.randi(Gr, 1, N) / Gr) / math.sqrt(N)  # :10-11
    if kind == "ps":
        Gr = N
        return np.exp(-1j * n * 2 * math.pi * np.arange(Gr)[None, :] / Gr) / math.sqrt(N)  # :13-14
    if kind == "ZC":
        R = 11
        return np.exp(-1j * R * n * math.pi * np.arange(1, N + 1)[None, :] / N) / math.sqrt(N)  # :16-17
    if kind in ("quantized_4", "quantized"):
        nq = 4 if kind == "quantized_4" else 6                   # :19,26
        A = np.arange(2 ** nq)
        K = int(math.ceil(N / A.size))
        A = np.tile(A, K)                                         # vec(kron(ones(K,1),A)).' :22
        omega = 2 * math.pi / 2 ** nq * A[:N]
        return np.exp(-1j * n * omega[None, :]) / math.sqrt(N)
    raise ValueError(kind)


QAM4 = np.array([(1 + 1j), (-1 + 1j), (1 - 1j), (-1 - 1j)]) / math.sqrt(2.0)  # qam4mod.m:7


def qam4mod(N, rng: RefRandom):
    """qam4mod([], 'mod', N) (qam4mod.m:6-8)."""
    return rng.randsrc(N, 1, QAM4).reshape(-1)


def qam4demod(soft):
    """qam4mod(input, 'demod') (qam4mod.m:12-31).  MATLAB's relational operators on complex numbers compare REAL parts only, so
    ``softDecision >= 0 & conjSymbols <= 0`` with ``conjSymbols = -1j*softDecision`` reads re >= 0 & im <= 0; rules apply in order."""
    x = np.asarray(soft, dtype=np.complex128)
    a = 1.0 / math.sqrt(2.0)
    out = np.full(x.shape, a + 1j * a)
    re, im = x.real, x.imag                                        # real(-1j*x) = imag(x)
    out[(re >= 0) & (im <= 0)] = a - 1j * a                        # :21,27
    out[(re <= 0) & (im >= 0)] = -a + 1j * a                       # :22,28
    out[(re <= 0) & (im <= 0)] = -a - 1j * a                       # :23,29
    return out


# ----------------------------------------------------------------------------
# proposed_hbf.m / hbf.m / wideband_hybBF_comm_system_training.m
# ----------------------------------------------------------------------------
def psi_bar_from_pilots(pilots, T, L):
    """``Psi_bar(k,:,l) = Psi_i(l,:,k)`` with ``Psi_i(:,:,k) = toeplitz(s_k)``
    (proposed_hbf.m:15-18, plot_errorVSsnr.m:63-67).  ``pilots`` is (Nt, T):
    row k is the sequence s_k.  Returns Psi_bar (Nt, T, L)."""
    Nt = pilots.shape[0]
    Psi_bar = np.zeros((Nt, T, L), dtype=np.complex128)
    for k in range(Nt):
        rows = toeplitz_hermitian_rows(pilots[k, :T], L)          # rows 1..L of toeplitz(s_k)
        for l in range(L):
            Psi_bar[k, :, l] = rows[l, :]
    return Psi_bar


def psi_i_literal(pilots, T):
    """Dense ``Psi_i (T,T,Nt)`` exactly as the drivers build it (plot_errorVSsnr.m:61-67)."""
    Nt = pilots.shape[0]
    Psi_i = np.zeros((T, T, Nt), dtype=np.complex128)
    for k in range(Nt):
        Psi_i[:, :, k] = toeplitz_hermitian(pilots[k, :T])
    return Psi_i


def received_signal(H, Psi_bar):
    """``Y = sum_l H(:,:,l)*Psi_bar(:,:,l)`` (proposed_hbf.m:14-20, hbf.m:12-18)."""
    L = H.shape[2]
    Y = np.zeros((H.shape[0], Psi_bar.shape[1]), dtype=np.complex128)
    for l in range(L):
        Y = Y + H[:, :, l] @ Psi_bar[:, :, l]
    return Y


def sampling_mask(rows, T, Lr, rng: RefRandom):
    """``Omega`` with exactly ``Lr`` ones per column from T x randperm
    (proposed_hbf.m:36-41, wideband_hybBF_comm_system_training.m:47-52)."""
    Omega = np.zeros((rows, T))
    for t in range(T):
        indices = rng.randperm(rows)
        Omega[indices[:Lr] - 1, t] = 1.0
    return Omega


def proposed_hbf(H, N, Psi_i, T, Lr_e, Lr, W, rng: RefRandom, Psi_bar=None):
    """[Y_proposed_hbf, W_e, Psi_bar, Omega, Y] = proposed_hbf(...) (proposed_hbf.m:1-44).
    ``Psi_i`` is the literal (T,T,Nt) array, or None when ``Psi_bar`` is supplied
    (identical values, avoids the T^2*Nt allocation at large shapes)."""
    _, Nt, L = H.shape
    if Psi_bar is None:
        Psi_bar = np.zeros((Nt, T, L), dtype=np.complex128)
        for l in range(L):
            for k in range(Nt):
                Psi_bar[k, :, l] = Psi_i[l, :, k]                 # :17
    W_e = W[:, :Lr_e]                                             # :11
    Y = received_signal(H, Psi_bar)                               # :14-20
    R = Y + N                                                     # :22
    Omega = sampling_mask(Lr_e, T, Lr, rng)                       # :36-41
    Y_proposed = Omega * (W_e.conj().T @ R)                       # :42
    return Y_proposed, W_e, Psi_bar, Omega, Y


def hbf(H, N, Psi_i, T, Lr, W, Psi_bar=None):
    """[Y_conventional_hbf, W_c, Psi_bar, Y] = hbf(...) (hbf.m:1-26)."""
    _, Nt, L = H.shape
    if Psi_bar is None:
        Psi_bar = np.zeros((Nt, T, L), dtype=np.complex128)
        for l in range(L):
            for k in range(Nt):
                Psi_bar[k, :, l] = Psi_i[l, :, k]                 # :15
    Y = received_signal(H, Psi_bar)
    R = Y + N                                                     # :20
    W_c = W[:, :Lr]                                               # :23
    return W_c.conj().T @ R, W_c, Psi_bar, Y                      # :24


def wideband_hybBF_comm_system_training(H, T, snr, ratio, rng: RefRandom):
    """[Yp, Yc, W_tilde, Psi_bar, Omega, Lr] = wideband_hybBF_comm_system_training(...)
    (wideband_hybBF_comm_system_training.m:1-58).  RNG order: randn(Nr,T) twice
    (:16), per k two randn(1,T) (:20), then T x randperm(Nr) (:49)."""
    Nr, Nt, L = H.shape
    Lr = mround(ratio * Nr)                                       # :5
    W_tilde = np.fft.fft(np.eye(Nr), axis=0) / math.sqrt(Nr)      # :10
    re = rng.randn(Nr, T)
    im = rng.randn(Nr, T)
    N = math.sqrt(snr / 2.0) * (re + 1j * im)                     # :16
    pilots = np.zeros((Nt, T), dtype=np.complex128)
    for k in range(Nt):
        sre = rng.randn(1, T)
        sim = rng.randn(1, T)
        pilots[k, :] = ((sre + 1j * sim) / math.sqrt(2.0)).reshape(-1)  # :20
    Psi_bar = psi_bar_from_pilots(pilots, T, L)                   # :21,27-29
    R = received_signal(H, Psi_bar) + N                           # :25-33
    Omega = sampling_mask(Nr, T, Lr, rng)                         # :47-52
    WR = W_tilde.conj().T @ R
    return Omega * WR, WR, W_tilde, Psi_bar, Omega, Lr            # :53,56


def dictionary_B(Dt, Psi_bar):
    """``B((l-1)*Gt+1:l*Gt,:) = Dt'*Psi_bar(:,:,l)`` (plot_errorVSsnr.m:133-136)."""
    L = Psi_bar.shape[2]
    return np.concatenate([Dt.conj().T @ Psi_bar[:, :, l] for l in range(L)], axis=0)
