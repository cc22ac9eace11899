"""Host-side mirror of the reference's MATLAB interface for the hot path.

Each function keeps the name, argument order and return values of the MATLAB
function it replaces (cited file:line, relative to the reference root) and
calls the C ABI with HOST buffers - this is the same call a MEX gateway makes
(``jstsp19_b200/mex/``).  Inputs are NumPy arrays in natural
``(rows, cols)`` indexing; a leading batch axis solves independent trials in one
call (2-D ``A`` / ``B`` / ``Omega`` are then shared by all trials).

``precision``: ``"f64"`` (default; complex128 arithmetic on the GPU, matches
the fp64 reference to ~1e-12) or ``"f32"`` (throughput path, ~1e-6 relative).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import AdmmDesc, default_handle

_CD = {"f32": np.complex64, "f64": np.complex128}
_RD = {"f32": np.float32, "f64": np.float64}
_DT = {"f32": _lib.F32, "f64": _lib.F64}


def _cm(x, dtype, nd_single=2):
    """(..., R, C) -> C-contiguous (..., C, R): per-trial column-major storage."""
    x = np.asarray(x)
    return np.ascontiguousarray(np.swapaxes(x, -1, -2), dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _batch_of(x, nd):
    x = np.asarray(x)
    if x.ndim == nd:
        return None
    if x.ndim == nd + 1:
        return x.shape[0]
    raise ValueError(f"expected {nd}-D or {nd + 1}-D array, got shape {x.shape}")


def _per_trial(v, batch):
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size == 1:
        a = np.full(batch, float(a[0]))
    if a.size != batch:
        raise ValueError("per-trial parameter has wrong length")
    return np.ascontiguousarray(a)


def _type_code(type_):
    # proposed_algorithm.m:23-30: 'approximate' or anything else (exact LS)
    return _lib.APPROXIMATE if type_ == "approximate" else _lib.STD


def _admm(subY, Omega, indx_S, A, B, Imax, tau_Y, tau_S, rho, type_, precision, handle, nargout, psi=None, pilots_L=None):
    h = handle or default_handle()
    cd, rd = _CD[precision], _RD[precision]
    bs = _batch_of(subY, 2)
    single = bs is None
    batch = 1 if single else bs
    sY = _cm(subY, cd)
    N, M = sY.shape[-1], sY.shape[-2]
    om = _cm(Omega, rd)
    Am = _cm(A, cd)
    if psi is None:
        Bm = _cm(B, cd)
        G, P = Am.shape[-2], Bm.shape[-1]
        bad = Bm.shape[-2] != M
    else:
        # Dt (Nt, Gt) [or (b, Nt, Gt)], Psi_bar (Nt, M, L) [or (b, Nt, M, L)] in MATLAB indexing -> column-major storage
        Dt, Psi_bar = np.asarray(psi[0]), np.asarray(psi[1])
        Dm = _cm(Dt, cd)
        if pilots_L is None:
            Pm = np.ascontiguousarray(np.moveaxis(Psi_bar, (-3, -2, -1), (-1, -2, -3)), dtype=cd)  # (..., L, M, Nt)
            Nt, Gt, L = Dm.shape[-1], Dm.shape[-2], Pm.shape[-3]
        else:
            Pm = _cm(Psi_bar, cd)                                                                  # pilot sequences (..., M, Nt)
            Nt, Gt, L = Dm.shape[-1], Dm.shape[-2], int(pilots_L)
        G, P = Am.shape[-2], L * Gt
        Bm = None
        bad = Pm.shape[-1] != Nt or Pm.shape[-2] != M
    if Am.shape[-1] != N or bad or om.shape[-2:] != (M, N):
        raise ValueError("inconsistent shapes: subY N x M, Omega N x M, A N x G, B P x M (or Dt Nt x Gt, Psi_bar Nt x M x L)")
    d = AdmmDesc()
    d.N, d.M, d.G, d.P, d.imax, d.type, d.batch = N, M, G, P, int(Imax), _type_code(type_), batch
    d.ld_subY = N * M
    d.ld_omega = N * M if om.ndim == 3 else 0
    d.ld_A = N * G if Am.ndim == 3 else 0
    d.ld_B = P * M if (Bm is not None and Bm.ndim == 3) else 0
    d.ld_S, d.ld_Y, d.ld_conv = G * P, N * M, 3 * int(Imax)
    tY, tS, rh = _per_trial(tau_Y, batch), _per_trial(tau_S, batch), _per_trial(rho, batch)
    S = np.empty((batch, P, G), dtype=cd)
    Y = np.empty((batch, M, N), dtype=cd) if nargout >= 2 else None
    conv = np.empty((batch, 3, int(Imax)), dtype=rd) if nargout >= 3 else None
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    ix = None
    if indx_S is not None:
        ixa = np.asarray(indx_S)
        # per-trial rankings only when the leading dimension is the batch and it is not MATLAB's (n, 1) column; anything else is one shared ranking
        per_trial = ixa.ndim == 2 and batch > 1 and ixa.shape[0] == batch and ixa.shape[1] != 1
        ix = np.ascontiguousarray(ixa.reshape(batch, -1) if per_trial else ixa.reshape(1, -1), dtype=np.int32)
        d.n_indx = ix.shape[1]
        d.ld_indx = ix.shape[1] if ix.shape[0] == batch and batch > 1 else 0
    if psi is not None and pilots_L is not None:
        rc = _lib.lib.jstsp_proposed_algorithm_pilots(h.ptr, C.byref(d), _DT[precision], _lib.HOST, _ptr(sY), _ptr(om), _ptr(ix), _ptr(Am),
                                                      _ptr(Dm), Nt * Gt if Dm.ndim == 3 else 0, _ptr(Pm), Nt * M if Pm.ndim == 3 else 0, Nt, L,
                                                      dp(tY), dp(tS), dp(rh), _ptr(S), _ptr(Y), _ptr(conv))
    elif psi is not None:
        rc = _lib.lib.jstsp_proposed_algorithm_psi(h.ptr, C.byref(d), _DT[precision], _lib.HOST, _ptr(sY), _ptr(om), _ptr(ix), _ptr(Am),
                                                   _ptr(Dm), Nt * Gt if Dm.ndim == 3 else 0, _ptr(Pm), Nt * M * L if Pm.ndim == 4 else 0, Nt, L,
                                                   dp(tY), dp(tS), dp(rh), _ptr(S), _ptr(Y), _ptr(conv))
    elif indx_S is None:
        rc = _lib.lib.jstsp_proposed_algorithm(h.ptr, C.byref(d), _DT[precision], _lib.HOST, _ptr(sY), _ptr(om), _ptr(Am), _ptr(Bm),
                                               dp(tY), dp(tS), dp(rh), _ptr(S), _ptr(Y), _ptr(conv))
    else:
        rc = _lib.lib.jstsp_proposed_algorithm_angles(h.ptr, C.byref(d), _DT[precision], _lib.HOST, _ptr(sY), _ptr(om), _ptr(ix),
                                                      _ptr(Am), _ptr(Bm), dp(tY), dp(tS), dp(rh), _ptr(S), _ptr(Y), _ptr(conv))
    h.check(rc)
    outs = [np.swapaxes(S, -1, -2)]
    if Y is not None:
        outs.append(np.swapaxes(Y, -1, -2))
    if conv is not None:
        outs.append(np.swapaxes(conv, -1, -2))
    if single:
        outs = [o[0] for o in outs]
    return tuple(outs) if len(outs) > 1 else outs[0]


def proposed_algorithm(subY, Omega, A, B, Imax, tau_Y, tau_S, rho, type, *, precision="f64", handle=None, nargout=3):
    """[S, Y, convergence_error] = proposed_algorithm(subY, Omega, A, B, Imax, tau_Y, tau_S, rho, type)
    (basic_system_functions/proposed_algorithm.m:1)."""
    return _admm(subY, Omega, None, A, B, Imax, tau_Y, tau_S, rho, type, precision, handle, nargout)


def proposed_algorithm_angles(subY, Omega, indx_S, A, B, Imax, tau_Y, tau_S, rho, type, greedy_nnz=None, *,
                              precision="f64", handle=None, nargout=3):
    """[S, Y, convergence_error] = proposed_algorithm_angles(subY, Omega, indx_S, A, B, Imax, tau_Y, tau_S,
    rho, type, greedy_nnz)  (basic_system_functions/proposed_algorithm_angles.m:1).  ``indx_S`` is 1-based
    like in MATLAB; ``greedy_nnz`` is accepted and ignored exactly like the reference does."""
    return _admm(subY, Omega, indx_S, A, B, Imax, tau_Y, tau_S, rho, type, precision, handle, nargout)


def proposed_algorithm_psi(subY, Omega, A, Dt, Psi_bar, Imax, tau_Y, tau_S, rho, type, indx_S=None, *,
                           precision="f64", handle=None, nargout=3):
    """proposed_algorithm / proposed_algorithm_angles with the dictionary given by the factors the reference's drivers
    hold before they form ``B((l-1)*Gt+1:l*Gt,:) = Dt'*Psi_bar(:,:,l)`` (plot_errorVSsnr.m:132-136): ``Dt`` (Nt x Gt) from
    wideband_mmwave_channel.m:1 and ``Psi_bar`` (Nt x M x L) from proposed_hbf.m:1.  Same outputs as
    ``proposed_algorithm(subY, Omega, A, B, ...)`` with that ``B``; Toeplitz 4-QAM pilots take the tensor-core path."""
    return _admm(subY, Omega, indx_S, A, None, Imax, tau_Y, tau_S, rho, type, precision, handle, nargout, psi=(Dt, Psi_bar))


def proposed_algorithm_pilots(subY, Omega, A, Dt, pilots, L, Imax, tau_Y, tau_S, rho, type, indx_S=None, *,
                              precision="f64", handle=None, nargout=3):
    """:func:`proposed_algorithm_psi` fed the pilot sequences themselves: ``pilots`` is Nt x M with row k = ``s_k``, the vector the
    drivers hand to ``toeplitz`` (plot_errorVSsnr.m:63-67); ``Psi_bar(k,:,l)`` = row l of ``toeplitz(s_k)`` (proposed_hbf.m:15-18) is
    formed on the device, so the call moves L times fewer dictionary bytes.  Same outputs."""
    return _admm(subY, Omega, indx_S, A, None, Imax, tau_Y, tau_S, rho, type, precision, handle, nargout, psi=(Dt, pilots), pilots_L=L)


def svt(Y, tau, *, precision="f64", handle=None):
    """X = svt(Y, tau)  (benchmark_algorithms/svt.m:1)."""
    h = handle or default_handle()
    cd = _CD[precision]
    bs = _batch_of(Y, 2)
    batch = 1 if bs is None else bs
    Ym = _cm(Y, cd)
    Mr, Mt = Ym.shape[-1], Ym.shape[-2]
    X = np.empty((batch, Mt, Mr), dtype=cd)
    t = _per_trial(tau, batch)
    h.check(_lib.lib.jstsp_svt(h.ptr, _DT[precision], _lib.HOST, Mr, Mt, batch, _ptr(Ym), Mr * Mt, _ptr(t), _ptr(X), Mr * Mt))
    X = np.swapaxes(X, -1, -2)
    return X[0] if bs is None else X


def mc_svt(OH, Omega, Imax, tau, rho, *, precision="f64", handle=None):
    """X = mc_svt(OH, Omega, Imax, tau, rho)  (benchmark_algorithms/mc_svt.m:1)."""
    h = handle or default_handle()
    cd, rd = _CD[precision], _RD[precision]
    bs = _batch_of(OH, 2)
    batch = 1 if bs is None else bs
    Om = _cm(OH, cd)
    Mr, Mt = Om.shape[-1], Om.shape[-2]
    om = _cm(Omega, rd)
    X = np.empty((batch, Mt, Mr), dtype=cd)
    t, r = _per_trial(tau, batch), _per_trial(rho, batch)
    h.check(_lib.lib.jstsp_mc_svt(h.ptr, _DT[precision], _lib.HOST, Mr, Mt, batch, int(Imax), _ptr(Om), Mr * Mt,
                                  _ptr(om), Mr * Mt if om.ndim == 3 else 0, _ptr(t), _ptr(r), _ptr(X), Mr * Mt))
    X = np.swapaxes(X, -1, -2)
    return X[0] if bs is None else X


def mc_admm(Htrue, OH, Omega, Imax, tau, rho, *, precision="f64", handle=None, nargout=2):
    """[X, convergence_error] = mc_admm(Htrue, OH, Omega, Imax, tau, rho)
    (benchmark_algorithms/mc_admm.m:1)."""
    h = handle or default_handle()
    cd, rd = _CD[precision], _RD[precision]
    bs = _batch_of(OH, 2)
    batch = 1 if bs is None else bs
    Om = _cm(OH, cd)
    Mr, Mt = Om.shape[-1], Om.shape[-2]
    om = _cm(Omega, rd)
    Ht = _cm(Htrue, cd) if (Htrue is not None and nargout >= 2) else None
    X = np.empty((batch, Mt, Mr), dtype=cd)
    conv = np.empty((batch, int(Imax)), dtype=rd) if Ht is not None else None
    t, r = _per_trial(tau, batch), _per_trial(rho, batch)
    h.check(_lib.lib.jstsp_mc_admm(h.ptr, _DT[precision], _lib.HOST, Mr, Mt, batch, int(Imax),
                                   _ptr(Ht), (Mr * Mt if (Ht is not None and Ht.ndim == 3) else 0), _ptr(Om), Mr * Mt,
                                   _ptr(om), Mr * Mt if om.ndim == 3 else 0, _ptr(t), _ptr(r), _ptr(X), Mr * Mt,
                                   _ptr(conv), int(Imax)))
    X = np.swapaxes(X, -1, -2)
    if bs is None:
        X = X[0]
        conv = conv[0] if conv is not None else None
    return (X, conv) if nargout >= 2 else X


def OMP(A, v, m, snr=None, *, precision="f64", handle=None, return_ambiguous=False):
    """[x_hat, indexSet, v, targetMatrix] = OMP(A, v, m, snr)  (benchmark_algorithms/OMP.m:1).
    ``indexSet`` is a list of 1-based indices (the reference returns a 1 x m cell of scalars);
    ``snr`` is accepted and ignored exactly like the reference does.  A leading batch axis on
    ``v`` (and optionally ``A``) solves independent problems in one call."""
    h = handle or default_handle()
    cd = _CD[precision]
    v_in = v
    va = np.asarray(v)
    single = va.ndim == 1 or (va.ndim == 2 and 1 in va.shape and np.asarray(A).ndim == 2)
    vb = np.ascontiguousarray(va.reshape(1, -1) if single else va.reshape(va.shape[0], -1), dtype=cd)
    batch, measures = vb.shape
    Am = _cm(A, cd)
    size_d = Am.shape[-2]
    if Am.shape[-1] != measures:
        raise ValueError("A must be measures x size_d")
    m = int(m)
    x = np.empty((batch, size_d), dtype=cd)
    idx = np.empty((batch, m), dtype=np.int32)
    tgt = np.empty((batch, m, measures), dtype=cd)
    amb = np.zeros(batch, dtype=np.int32)
    tol = 1e-10 if precision == "f64" else 1e-4
    h.check(_lib.lib.jstsp_omp(h.ptr, _DT[precision], _lib.HOST, measures, size_d, m, batch, _ptr(Am),
                               measures * size_d if Am.ndim == 3 else 0, _ptr(vb), measures, _ptr(x), size_d,
                               _ptr(idx), _ptr(tgt), _ptr(amb), tol))
    T = np.swapaxes(tgt, -1, -2)
    if single:
        out = (x[0], [int(k) for k in idx[0]], v_in, T[0])
        return out + (int(amb[0]),) if return_ambiguous else out
    out = (x, idx, v_in, T)
    return out + (amb,) if return_ambiguous else out


def OMP_kron(A, B, Y, m, *, precision="f64", handle=None, want_x_hat=True, return_ambiguous=False):
    """[x_hat, indexSet, x_sel, residual] = OMP(kron(B.', A), vec(Y), m) without forming the Kronecker
    dictionary (benchmark_algorithms/OMP.m:1-32 on the operands of plot_errorVSdelays.m:77-78).
    ``A`` N x G, ``B`` P x M, ``Y`` N x M; a leading batch axis on ``Y`` (and optionally on ``A`` / ``B``)
    solves independent trials.  ``indexSet`` holds 1-based linear indices into the G x P unknown;
    ``x_hat`` (G*P, ``None`` when ``want_x_hat`` is false), ``x_sel`` the coefficient of each pick."""
    h = handle or default_handle()
    cd = _CD[precision]
    bs = _batch_of(Y, 2)
    batch = 1 if bs is None else bs
    Ym = _cm(Y, cd).reshape(batch, -1)
    Am, Bm = _cm(A, cd), _cm(B, cd)
    G, N = Am.shape[-2], Am.shape[-1]
    M, P = Bm.shape[-2], Bm.shape[-1]
    if Ym.shape[1] != N * M:
        raise ValueError("Y must be size(A,1) x size(B,2)")
    m = int(m)
    x = np.empty((batch, G * P), dtype=cd) if want_x_hat else None
    idx = np.empty((batch, m), dtype=np.int32)
    xs = np.empty((batch, m), dtype=cd)
    res = np.empty((batch, M, N), dtype=cd)
    amb = np.zeros(batch, dtype=np.int32)
    tol = 1e-10 if precision == "f64" else 1e-4
    h.check(_lib.lib.jstsp_omp_kron(h.ptr, _DT[precision], _lib.HOST, N, M, G, P, m, batch,
                                    _ptr(Am), N * G if Am.ndim == 3 else 0, _ptr(Bm), P * M if Bm.ndim == 3 else 0,
                                    _ptr(Ym), N * M, _ptr(x), G * P, _ptr(idx), _ptr(xs), _ptr(res), N * M, _ptr(amb), tol))
    R = np.swapaxes(res, -1, -2)
    if bs is None:
        out = (x[0] if want_x_hat else None, [int(k) for k in idx[0]], xs[0], R[0])
        return out + (int(amb[0]),) if return_ambiguous else out
    out = (x, idx, xs, R)
    return out + (amb,) if return_ambiguous else out


def somp(A, Y, K, *, precision="f64", handle=None, res_tol=0.0):
    """Z, support, residual of the joint (MMV) OMP that stands in for
    ``spx.pursuit.joint.OrthogonalMatchingPursuit(A, K).solve(Y)`` (plot_errorVSsnr.m:116-118).
    ``A`` N x D, ``Y`` N x S (leading batch axis allowed); ``support`` is 1-based, in selection order."""
    h = handle or default_handle()
    cd = _CD[precision]
    bs = _batch_of(Y, 2)
    batch = 1 if bs is None else bs
    Ym = _cm(Y, cd)
    Am = _cm(A, cd)
    D, N = Am.shape[-2], Am.shape[-1]
    S = Ym.shape[-2]
    if Ym.shape[-1] != N:
        raise ValueError("Y must have size(A,1) rows")
    K = int(K)
    Z = np.empty((batch, S, D), dtype=cd)
    sup = np.zeros((batch, K), dtype=np.int32)
    nit = np.zeros(batch, dtype=np.int32)
    res = np.empty((batch, S, N), dtype=cd)
    h.check(_lib.lib.jstsp_somp(h.ptr, _DT[precision], _lib.HOST, N, D, S, K, batch, _ptr(Am), N * D if Am.ndim == 3 else 0,
                                _ptr(Ym.reshape(batch, -1)), N * S, _ptr(Z), D * S, _ptr(sup), _ptr(nit), _ptr(res), N * S, float(res_tol)))
    Zo, Ro = np.swapaxes(Z, -1, -2), np.swapaxes(res, -1, -2)
    if bs is None:
        return Zo[0], [int(k) for k in sup[0][: nit[0]]], Ro[0]
    return Zo, [[int(k) for k in sup[b][: nit[b]]] for b in range(batch)], Ro


class _SompResult:
    """Result object of :class:`spx_joint_OrthogonalMatchingPursuit` (fields the drivers read: ``Z``)."""

    def __init__(self, Z, support, R):
        self.Z, self.support, self.R, self.iterations = Z, support, R, len(support)


class spx_joint_OrthogonalMatchingPursuit:
    """Shim with the call shape of ``spx.pursuit.joint.OrthogonalMatchingPursuit(A, K)`` /
    ``.solve(Y)`` / ``.Z`` so the drivers' lines (plot_errorVSsnr.m:116-118) translate one to one."""

    def __init__(self, A, K, *, precision="f64", handle=None):
        self.A, self.K, self._kw = A, int(K), dict(precision=precision, handle=handle)

    def solve(self, Y):
        return _SompResult(*somp(self.A, Y, self.K, **self._kw))


class _OmpResult:
    def __init__(self, z, support):
        self.z, self.support, self.iterations = z, support, len(support)


class spx_single_OrthogonalMatchingPursuit:
    """Shim with the call shape of ``spx.pursuit.single.OrthogonalMatchingPursuit(Phi, K)`` / ``.solve(y)`` /
    ``.z`` (plot_time_comparisions.m:83-85), served by OMP.m's semantics (:func:`OMP`)."""

    def __init__(self, Phi, K, *, precision="f64", handle=None):
        self.Phi, self.K, self._kw = Phi, int(K), dict(precision=precision, handle=handle)

    def solve(self, y):
        x, idx, _, _ = OMP(self.Phi, np.asarray(y).reshape(-1), self.K, **self._kw)
        return _OmpResult(x, idx)


def sparse_admm(Htrue, OH, Dr, Dt, Imax, *, precision="f64", handle=None, nargout=2):
    """[S, convergence_error] = sparse_admm(Htrue, OH, Dr, Dt, Imax)  (benchmark_algorithms/sparse_admm.m:1)."""
    h = handle or default_handle()
    cd, rd = _CD[precision], _RD[precision]
    bs = _batch_of(OH, 2)
    batch = 1 if bs is None else bs
    Om = _cm(OH, cd)
    Mr, Mt = Om.shape[-1], Om.shape[-2]
    Drm, Dtm = _cm(Dr, cd), _cm(Dt, cd)
    if Drm.shape[-2:] != (Mr, Mr) or Dtm.shape[-2:] != (Mt, Mt):
        raise ValueError("sparse_admm needs square dictionaries (Gr == Mr, Gt == Mt), like the reference")
    Ht = _cm(Htrue, cd) if (Htrue is not None and nargout >= 2) else None
    S = np.empty((batch, Mt, Mr), dtype=cd)
    conv = np.empty((batch, int(Imax)), dtype=rd) if Ht is not None else None
    h.check(_lib.lib.jstsp_sparse_admm(h.ptr, _DT[precision], _lib.HOST, Mr, Mt, batch, int(Imax),
                                       _ptr(Ht), (Mr * Mt if Ht.ndim == 3 else 0) if Ht is not None else 0, _ptr(Om), Mr * Mt, _ptr(Drm), Mr * Mr if Drm.ndim == 3 else 0,
                                       _ptr(Dtm), Mt * Mt if Dtm.ndim == 3 else 0, _ptr(S), Mr * Mt, _ptr(conv), int(Imax)))
    S = np.swapaxes(S, -1, -2)
    if bs is None:
        S = S[0]
        conv = conv[0] if conv is not None else None
    return (S, conv) if nargout >= 2 else S


def vamp(y, A, sigma, L, *, precision="f64", handle=None, nit=100, damp=0.85):
    """x = vamp(y, A, sigma, L)  (benchmark_algorithms/vamp.m:1).  The spectral decomposition that
    vamp.m:32 obtains from MATLAB's ``svd`` is taken from the host's LAPACK here (NumPy); the 100
    VampGlmEst iterations run on the GPU."""
    h = handle or default_handle()
    cd, rd = _CD[precision], _RD[precision]
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    U, s, Vh = np.linalg.svd(A, full_matrices=True)               # complex svd of A; embedding doubles each singular value
    if m <= n:
        d = np.concatenate([s ** 2, np.zeros(m - s.size)])
    else:                                                         # VampGlmEst.m:72-86,407-411 work in the eigenbasis of A'A when M > N
        U, d = Vh.conj().T, s ** 2
    yv = np.ascontiguousarray(np.asarray(y).reshape(1, -1), dtype=cd)
    Am, Um = _cm(A, cd), _cm(U, cd)
    dv = np.ascontiguousarray(d, dtype=rd)
    x = np.empty((1, n), dtype=cd)
    sg, Ln = _per_trial(sigma, 1), _per_trial(L, 1)
    h.check(_lib.lib.jstsp_vamp(h.ptr, _DT[precision], _lib.HOST, m, n, 1, int(nit), float(damp), _ptr(yv), m, _ptr(Am), 0,
                                _ptr(sg), _ptr(Ln), _ptr(Um), 0, _ptr(dv), 0, _ptr(x), n))
    return x[0]


_BF_TYPES = {"fft": 0, "rand": 1, "rand_ps": 2, "ps": 3, "ZC": 4, "quantized_4": 5, "quantized": 6}
_QAM4 = np.array([1 + 1j, -1 + 1j, 1 - 1j, -1 - 1j]) / np.sqrt(2.0)


def createBeamformer(N, beamformer_type, *, draws=None, precision="f64", handle=None):
    """B = createBeamformer(N, beamformer_type)  (basic_system_functions/createBeamformer.m:1).  The random codebooks take their draws
    as an input: ``'rand'`` the N x N matrix ``randsrc(N,N,[1 -1 1j -1j])`` itself, ``'rand_ps'`` the row ``randi(32,1,N)``."""
    h = handle or default_handle()
    cd = _CD[precision]
    if beamformer_type not in _BF_TYPES:
        raise ValueError(f"unknown beamformer_type {beamformer_type!r}")
    t = _BF_TYPES[beamformer_type]
    N = int(N)
    d = None
    if t == 1:
        a = np.asarray(draws).reshape(N, N)
        alphabet = np.array([1, -1, 1j, -1j])
        d = np.ascontiguousarray(np.argmin(np.abs(a.T[..., None] - alphabet), axis=-1), dtype=np.int32)      # column-major indices
    elif t == 2:
        d = np.ascontiguousarray(np.asarray(draws).reshape(N), dtype=np.int32)
    B = np.empty((N, N), dtype=cd)
    h.check(_lib.lib.jstsp_create_beamformer(h.ptr, _DT[precision], _lib.HOST, N, t, _ptr(d), _ptr(B)))
    return B.T


def qam4mod(input, mode, N=None, *, draws=None, precision="f64", handle=None):
    """symbols = qam4mod(input, mode, N)  (basic_system_functions/qam4mod.m:1).  ``'mod'``: ``draws`` is the N x 1 ``randsrc`` result
    (alphabet symbols) or their indices 0..3; ``'demod'``: hard decision of ``input``."""
    h = handle or default_handle()
    cd = _CD[precision]
    if mode == "mod":
        a = np.asarray(draws).reshape(-1)
        idx = a.astype(np.int32) if np.issubdtype(a.dtype, np.integer) else np.argmin(np.abs(a[:, None] - _QAM4), axis=-1).astype(np.int32)
        out = np.empty(idx.size, dtype=cd)
        h.check(_lib.lib.jstsp_qam4mod(h.ptr, _DT[precision], _lib.HOST, 0, idx.size, _ptr(np.ascontiguousarray(idx)), None, _ptr(out)))
        return out
    if mode == "demod":
        x = np.asarray(input)
        xm = np.ascontiguousarray(x.T if x.ndim == 2 else x, dtype=cd)
        out = np.empty(xm.shape, dtype=cd)
        h.check(_lib.lib.jstsp_qam4mod(h.ptr, _DT[precision], _lib.HOST, 1, xm.size, None, _ptr(xm), _ptr(out)))
        return out.T if x.ndim == 2 else out
    raise ValueError("mode must be 'mod' or 'demod'")


def wideband_mmwave_channel(L, Mr, Mt, total_num_of_clusters, total_num_of_rays, Gr, Gt, *, normals, uniforms,
                            precision="f64", handle=None):
    """[H, Zbar, Ar, At, Dr, Dt] = wideband_mmwave_channel(L, Mr, Mt, Ncl, Nray, Gr, Gt)
    (basic_system_functions/wideband_mmwave_channel.m:1).  ``normals`` / ``uniforms``: the randn / rand
    draws in the reference's order, shape (L*Ncl*Nray, 2) each (a MEX gateway takes them from MATLAB's RNG)."""
    h = handle or default_handle()
    cd = _CD[precision]
    Np = total_num_of_clusters * total_num_of_rays
    nr = np.ascontiguousarray(np.asarray(normals, dtype=np.float64).reshape(L * Np * 2))
    un = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(L * Np * 2))
    H = np.empty((L, Mt, Mr), dtype=cd); Zbar = np.empty((L * Gt, Gr), dtype=cd)
    Ar = np.empty((L, Np, Mr), dtype=cd); At = np.empty((L, Np, Mt), dtype=cd)
    Dr = np.empty((Gr, Mr), dtype=cd); Dt = np.empty((Gt, Mt), dtype=cd)
    h.check(_lib.lib.jstsp_wideband_mmwave_channel(h.ptr, _DT[precision], _lib.HOST, L, Mr, Mt, total_num_of_clusters, total_num_of_rays,
                                                   Gr, Gt, 1, _ptr(nr), _ptr(un), _ptr(H), _ptr(Zbar), _ptr(Ar), _ptr(At), _ptr(Dr), _ptr(Dt)))
    t3 = lambda a: np.transpose(a, (2, 1, 0))                     # (L, cols, rows) storage -> rows x cols x L
    return t3(H), Zbar.T, t3(Ar), t3(At), Dr.T, Dt.T


def _measure(H, N, Psi_i, T, Wc, Lr, W, perm, precision, handle, pilots=None):
    h = handle or default_handle()
    cd, rd = _CD[precision], _RD[precision]
    H = np.asarray(H)
    Nr, Nt, L = H.shape
    d = _lib.MeasDesc()
    d.Nr, d.Nt, d.L, d.T, d.Wc, d.Lr, d.batch = Nr, Nt, L, int(T), int(Wc), int(Lr), 1
    Hm = np.ascontiguousarray(np.transpose(H, (2, 1, 0)), dtype=cd)
    Nm = _cm(N, cd) if N is not None else None
    if pilots is not None:
        d.psi_mode, d.Tp = 1, 0
        Pm = _cm(pilots, cd)
    else:
        Psi_i = np.asarray(Psi_i)
        d.psi_mode, d.Tp = 0, Psi_i.shape[0]
        Pm = np.ascontiguousarray(np.transpose(Psi_i, (2, 1, 0)), dtype=cd)
    Wm = _cm(W, cd)
    pm = np.ascontiguousarray(perm, dtype=np.int32) if perm is not None else None
    Y = np.empty((T, Wc), dtype=cd); We = np.empty((Wc, Nr), dtype=cd); Pb = np.empty((L, T, Nt), dtype=cd)
    Om = np.empty((T, Wc), dtype=rd) if perm is not None else None
    Yn = np.empty((T, Nr), dtype=cd)
    h.check(_lib.lib.jstsp_measure(h.ptr, C.byref(d), _DT[precision], _lib.HOST, _ptr(Hm), _ptr(Nm), _ptr(Pm), _ptr(Wm), _ptr(pm),
                                   _ptr(Y), _ptr(We), _ptr(Pb), _ptr(Om), _ptr(Yn)))
    return Y.T, We.T, np.transpose(Pb, (2, 1, 0)), (Om.T if Om is not None else None), Yn.T


def proposed_hbf(H, N, Psi_i, T, Lr_e, Lr, W, *, perm, precision="f64", handle=None, pilots=None):
    """[Y_proposed_hbf, W_e, Psi_bar, Omega, Y] = proposed_hbf(H, N, Psi_i, T, Lr_e, Lr, W)
    (basic_system_functions/proposed_hbf.m:1).  ``perm``: the T randperm(Lr_e) draws of :38, shape (T, Lr_e),
    1-based.  ``pilots`` (Nt x T) may replace the dense ``Psi_i`` (T x T x Nt) - same values, no T^2 array."""
    return _measure(H, N, Psi_i, T, Lr_e, Lr, W, perm, precision, handle, pilots)


def hbf(H, N, Psi_i, T, Lr, W, *, precision="f64", handle=None, pilots=None):
    """[Y_conventional_hbf, W_c, Psi_bar, Y] = hbf(H, N, Psi_i, T, Lr, W)  (basic_system_functions/hbf.m:1)."""
    Y, Wc, Pb, _, Yn = _measure(H, N, Psi_i, T, Lr, 0, W, None, precision, handle, pilots)
    return Y, Wc, Pb, Yn


def wideband_hybBF_comm_system_training(H, T, snr, subSamplingRatio, *, noise_normals, pilot_normals, perm, precision="f64", handle=None):
    """[Y_proposed_hbf, Y_conventional_hbf, W_tilde, Psi_bar, Omega, Lr] = wideband_hybBF_comm_system_training(H, T, snr, ratio)
    (basic_system_functions/wideband_hybBF_comm_system_training.m:1).  Draws in the reference's order:
    ``noise_normals`` (2, Nr, T) = the two randn(Nr,T) of :16; ``pilot_normals`` (Nt, 2, T) = the randn(1,T) pairs of :20;
    ``perm`` (T, Nr) = the randperm(Nr) of :49."""
    import math
    H = np.asarray(H)
    Nr, Nt, L = H.shape
    Lr = int(math.floor(abs(subSamplingRatio * Nr) + 0.5))        # MATLAB round (:5)
    nn, pn = np.asarray(noise_normals, dtype=np.float64), np.asarray(pilot_normals, dtype=np.float64)
    N = math.sqrt(snr / 2.0) * (nn[0] + 1j * nn[1])                # :16
    pilots = (pn[:, 0, :] + 1j * pn[:, 1, :]) / math.sqrt(2.0)     # :20
    W = np.fft.fft(np.eye(Nr), axis=0) / math.sqrt(Nr)             # :10
    Yp, _, Pb, Om, _ = _measure(H, N, None, T, Nr, Lr, W, perm, precision, handle, pilots)
    Yc, _, _, _, _ = _measure(H, N, None, T, Nr, 0, W, None, precision, handle, pilots)
    return Yp, Yc, W, Pb, Om, Lr


def nmse(S, Zbar, *, precision="f64", handle=None):
    """min(1, norm(S-Zbar)^2/norm(Zbar)^2), matrix 2-norms (plot_errorVSsnr.m:138-141)."""
    h = handle or default_handle()
    cd = _CD[precision]
    bs = _batch_of(S, 2)
    batch = 1 if bs is None else bs
    Sm, Zm = _cm(S, cd), _cm(Zbar, cd)
    G, P = Sm.shape[-1], Sm.shape[-2]
    out = np.empty(batch, dtype=np.float64)
    h.check(_lib.lib.jstsp_nmse(h.ptr, _DT[precision], _lib.HOST, G, P, batch, _ptr(Sm), G * P, _ptr(Zm), G * P, _ptr(out)))
    return float(out[0]) if bs is None else out


def admm_parameters(Y, Zbar, rho_rule="sigma6", *, precision="f64", handle=None):
    """(tau_Y, tau_Z, rho) of plot_errorVSsnr.m:127-130 (rho_rule 'sigma6') or plot_errorVSdelays.m:127-128 ('sigma1')."""
    h = handle or default_handle()
    cd = _CD[precision]
    Ym, Zm = _cm(Y, cd), _cm(Zbar, cd)
    N, M, G, P = Ym.shape[-1], Ym.shape[-2], Zm.shape[-1], Zm.shape[-2]
    tY, tZ, rho = (np.empty(1, dtype=np.float64) for _ in range(3))
    h.check(_lib.lib.jstsp_admm_parameters(h.ptr, _DT[precision], _lib.HOST, N, M, G, P, 1, 6 if rho_rule == "sigma6" else 1,
                                           _ptr(Ym), N * M, _ptr(Zm), G * P, _ptr(tY), _ptr(tZ), _ptr(rho)))
    return float(tY[0]), float(tZ[0]), float(rho[0])


def log2det_rate(X, scale, *, precision="f64", handle=None):
    """real(log2(det(eye(n) + scale*X*X'))): the rate metric of plot_rateVSframelength.m:113,130 (X = Zbar,
    scale = 1/(Nr*(sigma2 + nmse))) and the capacity of plot_capacity.m:47-66 (X = W_c'*Y, scale = 1/(sigma2*Nt))."""
    h = handle or default_handle()
    cd = _CD[precision]
    bs = _batch_of(X, 2)
    batch = 1 if bs is None else bs
    Xm = _cm(X, cd)
    n, m = Xm.shape[-1], Xm.shape[-2]
    sc = _per_trial(scale, batch)
    out = np.empty(batch, dtype=np.float64)
    h.check(_lib.lib.jstsp_log2det_rate(h.ptr, _DT[precision], _lib.HOST, n, m, batch, _ptr(Xm), n * m, _ptr(sc), _ptr(out)))
    return float(out[0]) if bs is None else out


def ls_estimate(A, Y, B, *, precision="f64", handle=None, want_YpinvB=False):
    """S_ls = pinv(A)*Y*pinv(B), the least-squares baseline of plot_errorVSsnr.m:83 and plot_errorVSsnr_approx.m:61,67; with
    ``want_YpinvB`` also Y*pinv(B), the right-hand sides the drivers hand to the joint OMP (plot_errorVSsnr.m:117).
    A (N x G) and B (P x M) may be shared or per trial; Y is (N x M) or (batch x N x M)."""
    h = handle or default_handle()
    cd = _CD[precision]
    bs = _batch_of(Y, 2)
    batch = 1 if bs is None else bs
    Am, Bm, Ym = _cm(A, cd), _cm(B, cd), _cm(Y, cd)
    N, G = Am.shape[-1], Am.shape[-2]
    P, M = Bm.shape[-1], Bm.shape[-2]
    if Ym.shape[-1] != N or Ym.shape[-2] != M:
        raise ValueError("Y must be N x M with A N x G and B P x M")
    S = np.empty((batch, P, G), dtype=cd)
    YpB = np.empty((batch, P, N), dtype=cd) if want_YpinvB else None
    rc = _lib.lib.jstsp_ls_estimate(h.ptr, _DT[precision], _lib.HOST, N, M, G, P, batch, _ptr(Am), N * G if Am.ndim == 3 else 0,
                                    _ptr(Bm), P * M if Bm.ndim == 3 else 0, _ptr(Ym), N * M, _ptr(S), G * P, _ptr(YpB), N * P)
    h.check(rc)
    S = np.swapaxes(S, -1, -2)
    out = S[0] if bs is None else S
    if want_YpinvB:
        YpB = np.swapaxes(YpB, -1, -2)
        return out, (YpB[0] if bs is None else YpB)
    return out


def capacity(Y, W, Mr, scale, cols=None, *, precision="f64", handle=None):
    """real(log2(det(eye(Mr) + scale*Wsel'*(Y*Y')*Wsel))) with Wsel = W(:, cols(1:Mr)) - one receiver design of plot_capacity.m:47-64 /
    plot_ee.m:47-64 (scale = 1/square_noise_variance*1/Nt).  ``cols`` are 1-based (ind = randperm(Mr_e), plot_capacity.m:63);
    None keeps the first Mr columns (W_c = W(:, 1:Lr), hbf.m:24).  Y is the noiseless block (Nr x T) or a batch of them."""
    h = handle or default_handle()
    cd = _CD[precision]
    bs = _batch_of(Y, 2)
    batch = 1 if bs is None else bs
    Ym, Wm = _cm(Y, cd), _cm(W, cd)
    Nr, T = Ym.shape[-1], Ym.shape[-2]
    Wc = Wm.shape[-2]
    sc = _per_trial(scale, batch)
    out = np.empty(batch, dtype=np.float64)
    ci, ldc = None, 0
    if cols is not None:
        ci = np.ascontiguousarray(np.asarray(cols)[..., :Mr], dtype=np.int32)
        ldc = Mr if ci.ndim == 2 else 0
    h.check(_lib.lib.jstsp_capacity(h.ptr, _DT[precision], _lib.HOST, Nr, T, Wc, int(Mr), batch, _ptr(Ym), Nr * T, _ptr(Wm), Nr * Wc if Wm.ndim == 3 else 0,
                                    _ptr(ci), ldc, _ptr(sc), _ptr(out)))
    return float(out[0]) if bs is None else out


def capacity_sweep(Y, W_zc, W_q, Mr_range, ind, scale, *, precision="f64", handle=None):
    """Spectral efficiency of the four receiver designs of plot_capacity.m:45-64 for every Mr of ``Mr_range`` on a batch of noiseless blocks ``Y`` (b, Nr, T):
    returns (len(Mr_range), 4, b) = [digital BF, conventional HBF with phase shifters, conventional HBF with ZC, proposed].  ``ind`` holds the 1-based
    permutations ind = randperm(Mr_e) of the proposed design, one row per trial (or one shared row); ``scale`` = 1/(square_noise_variance*Nt)."""
    h = handle or default_handle()
    cd = _CD[precision]
    Ym = _cm(Y, cd)
    if Ym.ndim == 2:
        Ym = Ym[None]
    batch, T, Nr = Ym.shape
    Wz, Wq = _cm(W_zc, cd), _cm(W_q, cd)
    mr = np.ascontiguousarray(np.asarray(Mr_range), dtype=np.int32)
    ci = np.ascontiguousarray(np.asarray(ind), dtype=np.int32)
    ldi = ci.shape[-1] if ci.ndim == 2 and ci.shape[0] == batch and batch > 1 else 0
    sc = _per_trial(scale, batch)
    out = np.empty((mr.size, 4, batch), dtype=np.float64)
    h.check(_lib.lib.jstsp_capacity_sweep(h.ptr, _DT[precision], _lib.HOST, Nr, T, batch, int(mr.size), _ptr(mr), _ptr(Ym), Nr * T, _ptr(Wz), _ptr(Wq),
                                          _ptr(ci), ldi, _ptr(sc), _ptr(out)))
    return out


def power_model(Nr, Mr, Mr_e):
    """(power_dbf, power_hbf, power_hbf_zc, power_proposed) of plot_ee.m:69-77."""
    out = np.empty(4, dtype=np.float64)
    if _lib.lib.jstsp_power_model(int(Nr), int(Mr), int(Mr_e), _ptr(out)) != 0:
        raise ValueError("power_model: bad argument")
    return tuple(float(v) for v in out)


def energy_efficiency(mean_capacity4, Nr, Mr, Mr_e):
    """ee_dbf, ee_hbf_ps, ee_hbf_zc, ee_proposed = mean capacity / power of the design (plot_ee.m:84-87)."""
    pw = power_model(Nr, Mr, Mr_e)
    return tuple(float(c) / p for c, p in zip(mean_capacity4, pw))
