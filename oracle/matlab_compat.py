"""MATLAB built-in semantics the reference path depends on (test infrastructure).

Every helper names the MATLAB built-in it restates and the reference call
site(s) that rely on the behaviour (paths relative to the reference root).
"""
from __future__ import annotations

import math

import numpy as np

EPS = float(np.finfo(np.float64).eps)  # MATLAB ``eps`` (vamp.m:45, VampGlmEst.m:401)


def vec(X):
    """``X(:)`` - column-major vectorisation (basic_system_functions/vec.m:2)."""
    return np.asarray(X).reshape(-1, order="F")


def unvec(x, rows, cols):
    """``reshape(x, rows, cols)`` - column-major (proposed_algorithm.m:40,57)."""
    return np.asarray(x).reshape((rows, cols), order="F")


def mround(x):
    """MATLAB ``round``: half away from zero (plot_errorVSsnr.m:22,
    wideband_hybBF_comm_system_training.m:5).  NumPy rounds half to even."""
    return int(math.floor(abs(x) + 0.5) * (1 if x >= 0 else -1))


def toeplitz_hermitian(s):
    """MATLAB one-argument ``toeplitz(s)`` for a (complex) row vector
    (wideband_hybBF_comm_system_training.m:21, plot_errorVSsnr.m:66).

    First ROW is ``s``; below the diagonal the entries are conjugated; the main
    diagonal is ``s(1)`` itself:  T(i,j) = s(j-i) for j>=i, conj(s(i-j)) for i>j.
    """
    s = np.asarray(s).reshape(-1)
    n = s.size
    i = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    d = j - i
    return np.where(d >= 0, s[np.abs(d)], np.conj(s[np.abs(d)]))


def toeplitz_hermitian_rows(s, nrows):
    """First ``nrows`` rows of :func:`toeplitz_hermitian` without building the
    T x T matrix (the path only ever reads rows 1..L: proposed_hbf.m:17)."""
    s = np.asarray(s).reshape(-1)
    n = s.size
    i = np.arange(nrows)[:, None]
    j = np.arange(n)[None, :]
    d = j - i
    return np.where(d >= 0, s[np.abs(d)], np.conj(s[np.abs(d)]))


def norm2(X):
    """MATLAB ``norm(X)``: largest singular value for a matrix, Euclidean norm
    for a vector (plot_errorVSsnr.m:84,138; proposed_algorithm.m:51,67,69)."""
    X = np.asarray(X)
    if X.ndim == 1 or 1 in X.shape:
        return float(np.linalg.norm(X.reshape(-1)))
    if X.size == 0:
        return 0.0
    return float(np.linalg.norm(X, 2))


def fro(X):
    """``norm(X,'fro')`` (plot_errorVSsnr.m:127-128)."""
    return float(np.linalg.norm(np.asarray(X).reshape(-1)))


def eigs6(Mh):
    """MATLAB ``eigs(M)`` with defaults: the 6 largest-magnitude eigenvalues
    (plot_errorVSsnr.m:129).  ``M`` is Hermitian PSD at every call site."""
    w = np.linalg.eigvalsh(np.asarray(Mh))
    w = w[np.argsort(-np.abs(w), kind="stable")]
    return w[: min(6, w.size)]


def sort_descend_idx(x):
    """``[~, idx] = sort(x, 'descend')`` - stable, 1-based (plot_errorVSsnr.m:143)."""
    x = np.asarray(x).reshape(-1)
    return np.argsort(-x, kind="stable") + 1


def argmax_first(x):
    """``[~, idx] = max(x)``: first maximal element, 0-based here (OMP.m:17)."""
    return int(np.argmax(np.asarray(x).reshape(-1)))


def msign(x):
    """MATLAB ``sign`` for reals: sign(0) = 0 (proposed_algorithm.m:56)."""
    return np.sign(x)


def soft_complex(v, thr):
    """Separate real/imaginary soft threshold (proposed_algorithm.m:56,
    sparse_admm.m:22)."""
    re = np.maximum(np.abs(v.real) - thr, 0.0) * np.sign(v.real)
    im = np.maximum(np.abs(v.imag) - thr, 0.0) * np.sign(v.imag)
    return re + 1j * im


class RefRandom:
    """The reference's RNG calls (``randn``, ``rand``, ``randperm``, ``randsrc``)
    served from a seeded NumPy generator, filling arrays in MATLAB's
    column-major order.  The reference never seeds; the
    harness injects this object so oracle and engine see identical draws."""

    def __init__(self, seed=0):
        self.g = np.random.default_rng(seed)

    def randn(self, *shape):
        if len(shape) == 0 or shape == (1,):
            return float(self.g.standard_normal())
        n = int(np.prod(shape))
        return self.g.standard_normal(n).reshape(shape, order="F")

    def rand(self, *shape):
        if len(shape) == 0 or shape == (1,):
            return float(self.g.random())
        n = int(np.prod(shape))
        return self.g.random(n).reshape(shape, order="F")

    def randperm(self, n):
        """1-based permutation, like MATLAB."""
        return self.g.permutation(n) + 1

    def randsrc(self, rows, cols, alphabet):
        """``randsrc(rows, cols, alphabet)``: i.i.d. uniform draws from the
        alphabet (qam4mod.m:8).  Communications-Toolbox internals are not
        reproducible outside MATLAB; the harness treats pilots as inputs."""
        alphabet = np.asarray(alphabet)
        idx = np.minimum((self.rand(rows, cols) * alphabet.size).astype(int), alphabet.size - 1)
        return alphabet[idx]

    def randi(self, imax, rows, cols):
        return np.minimum((self.rand(rows, cols) * imax).astype(int), imax - 1) + 1
