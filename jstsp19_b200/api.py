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


def _admm(subY, Omega, indx_S, A, B, Imax, tau_Y, tau_S, rho, type_, precision, handle, nargout):
    h = handle or default_handle()
    cd, rd = _CD[precision], _RD[precision]
    bs = _batch_of(subY, 2)
    single = bs is None
    batch = 1 if single else bs
    sY = _cm(subY, cd)
    N, M = sY.shape[-1], sY.shape[-2]
    om = _cm(Omega, rd)
    Am = _cm(A, cd)
    Bm = _cm(B, cd)
    G, P = Am.shape[-2], Bm.shape[-1]
    if Am.shape[-1] != N or Bm.shape[-2] != M or om.shape[-2:] != (M, N):
        raise ValueError("inconsistent shapes: subY N x M, Omega N x M, A N x G, B P x M")
    d = AdmmDesc()
    d.N, d.M, d.G, d.P, d.imax, d.type, d.batch = N, M, G, P, int(Imax), _type_code(type_), batch
    d.ld_subY = N * M
    d.ld_omega = N * M if om.ndim == 3 else 0
    d.ld_A = N * G if Am.ndim == 3 else 0
    d.ld_B = P * M if Bm.ndim == 3 else 0
    d.ld_S, d.ld_Y, d.ld_conv = G * P, N * M, 3 * int(Imax)
    tY, tS, rh = _per_trial(tau_Y, batch), _per_trial(tau_S, batch), _per_trial(rho, batch)
    S = np.empty((batch, P, G), dtype=cd)
    Y = np.empty((batch, M, N), dtype=cd) if nargout >= 2 else None
    conv = np.empty((batch, 3, int(Imax)), dtype=rd) if nargout >= 3 else None
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    if indx_S is None:
        rc = _lib.lib.jstsp_proposed_algorithm(h.ptr, C.byref(d), _DT[precision], _lib.HOST, _ptr(sY), _ptr(om), _ptr(Am), _ptr(Bm),
                                               dp(tY), dp(tS), dp(rh), _ptr(S), _ptr(Y), _ptr(conv))
    else:
        ix = np.ascontiguousarray(np.asarray(indx_S).reshape(batch, -1) if np.asarray(indx_S).ndim > 1
                                  else np.asarray(indx_S).reshape(1, -1), dtype=np.int32)
        d.n_indx = ix.shape[1]
        d.ld_indx = ix.shape[1] if ix.shape[0] == batch and batch > 1 else 0
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
