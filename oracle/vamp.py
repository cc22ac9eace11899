"""Oracle restatement of ``vamp.m`` and the GAMPmatlab code it drives (test
infrastructure - see ``oracle/__init__.py``).  fp64.

Files followed (relative to the reference root, under
benchmark_algorithms/):
  vamp.m:1-55
  MPbased_solvers/VAMP/VampGlmEst.m:324-521   (main loop)
  MPbased_solvers/VAMP/VampGlmOpt.m:5-9,25-27 (defaults)
  MPbased_solvers/main/SparseScaEstim.m:76-174
  MPbased_solvers/main/CAwgnEstimIn.m:94-102,181-184
  MPbased_solvers/main/CAwgnEstimOut.m:97-108
"""
from __future__ import annotations

import math

import numpy as np

from .matlab_compat import EPS

GAM_MIN = 1e-8      # VampGlmOpt.m:7
GAM_MAX = 1e14      # VampGlmOpt.m:8
GAM1X_INIT = 1e-8   # VampGlmOpt.m:25
GAM1Z_INIT = 1e-8   # VampGlmOpt.m:27
NIT_MAX = 100       # vamp.m:9,38
DAMP = 0.85         # vamp.m:11,40


def _clip(g):
    return min(max(g, GAM_MIN), GAM_MAX)


def sparse_sca_estim(rhat, rvar, p1, var0):
    """SparseScaEstim.estim with estim1 = CAwgnEstimIn(0, var0), x0 = 0
    (SparseScaEstim.m:76-165).  ``rhat`` is complex-typed in vamp.m because
    r1init = eps*1i (vamp.m:45), so the complex branch (:101-103) is taken."""
    rhat = np.asarray(rhat, dtype=np.complex128)
    loglike1 = -(math.log(math.pi) + np.log(var0 + rvar) + np.abs(rhat) ** 2 / (var0 + rvar))  # CAwgnEstimIn.m:181-184
    rvar = np.maximum(rvar, EPS)                                   # SparseScaEstim.m:96
    loglike0 = -(math.log(math.pi) + np.log(rvar) + np.abs(rhat) ** 2 / rvar)  # :101-103
    exparg = loglike0 - loglike1 + math.log(1.0 - p1) - math.log(p1)  # :108
    exparg = np.maximum(np.minimum(exparg, 500.0), -500.0)         # :109-110
    py1 = 1.0 / (1.0 + np.exp(exparg))                             # :111
    py0 = 1.0 - py1                                                # :112
    gain = var0 / (var0 + rvar)                                    # CAwgnEstimIn.m:100
    xhat1 = gain * rhat                                            # :101 (mean0 = 0)
    xvar1 = gain * rvar                                            # :102
    xhat = py1 * xhat1                                             # SparseScaEstim.m:161
    xvar = py1 * (np.abs(xhat1) ** 2 - np.abs(xhat) ** 2) + py1 * xvar1 + py0 * (0.0 - np.abs(xhat) ** 2)  # :164-165
    return xhat, xvar


def cawgn_estim_out(y, wvar, phat, pvar):
    """CAwgnEstimOut.estim with scale = 1 (CAwgnEstimOut.m:97-108)."""
    gain = pvar / (pvar + wvar)
    zhat = gain * (y - phat) + phat
    zvar = wvar * gain
    return zhat, zvar


def vamp_literal(y, A, sigma, L, nit=NIT_MAX, return_state=False):
    """x = vamp(y, A, sigma, L) (vamp.m:1-55): real embedding (:3-4), full svd
    (:32-34), VampGlmEst for exactly ``nit`` iterations (VampGlmEst.m:509-511)."""
    A = np.asarray(A, dtype=np.complex128)
    y = np.asarray(y, dtype=np.complex128).reshape(-1)
    Bm = np.block([[A.real, -A.imag], [A.imag, A.real]])           # vamp.m:3
    b = np.concatenate([y.real, y.imag])                           # vamp.m:4
    M, N = Bm.shape                                                # MM, nx
    wvar = float(sigma)                                            # :20
    beta = float(L) / N                                            # :23
    var0 = 1.0 / beta                                              # :24
    if M <= N:
        U, s, _ = np.linalg.svd(Bm, full_matrices=True)            # :32
        d = np.concatenate([s ** 2, np.zeros(M - s.size)])         # :34
    else:
        # vamp.m only passes U,d; VampGlmEst recomputes eig(A'A) when M>N (:72-86)
        d, V = np.linalg.eigh(Bm.T @ Bm)
    dele = M / N                                                   # VampGlmEst.m:249
    r1 = np.complex128(EPS * 1j)                                   # vamp.m:45 (scalar, broadcast)
    p1 = np.zeros(M)                                               # VampGlmEst.m:331
    gam1x = GAM1X_INIT
    gam1z = GAM1Z_INIT
    x1 = z2 = None
    gam2z = None
    for i in range(1, nit + 1):
        if i > 1:                                                  # :357-362
            x1old, z2old, gam2zold, gam1xold = x1, z2, gam2z, gam1x
        x1, xvar1 = sparse_sca_estim(r1, np.full(N, 1.0 / gam1x), beta, var0)  # :364
        eta1x = 1.0 / float(np.mean(xvar1))                        # :365
        if i > 1:
            x1 = DAMP * x1 + (1 - DAMP) * x1old                    # :367
        gam2x = eta1x - gam1x                                      # :369
        r2 = (x1 * eta1x - r1 * gam1x) / gam2x                     # :370 (unclipped gam2x)
        gam2x = _clip(gam2x)                                       # :379
        z1, zvar1 = cawgn_estim_out(b, wvar, p1, np.full(M, 1.0 / gam1z))  # :381
        eta1z = 1.0 / float(np.mean(zvar1))                        # :382
        gam2z = eta1z - gam1z                                      # :383
        p2 = (z1 * eta1z - p1 * gam1z) / gam2z                     # :384
        gam2z = _clip(gam2z)                                       # :393
        if i > 1:
            gam2z = DAMP * gam2z + (1 - DAMP) * gam2zold           # :395
        inv = 1.0 / (d + gam2x / gam2z)                            # :400
        alf = float(d @ inv) / N - EPS                             # :401
        if M <= N:                                                 # :402-406
            Ar2 = Bm @ r2
            t = (U.T @ (p2 - Ar2)) * inv
            x2 = r2 + Bm.T @ (U @ t)
            z2 = Ar2 + U @ (d * t)
        else:                                                      # :407-411
            t = V.T @ (r2 * (gam2x / gam2z) + Bm.T @ p2)
            x2 = V @ (t * inv)
            z2 = Bm @ x2
        if i > 1:
            z2 = DAMP * z2 + (1 - DAMP) * z2old                    # :415
        r1 = (x2 - r2 * (1 - alf)) / alf                           # :467
        p1 = (dele * z2 - p2 * alf) / (dele - alf)                 # :468
        gam1x = _clip(gam2x * alf / (1 - alf))                     # :472,481
        gam1z = _clip(gam2z * (dele - alf) / alf)                  # :482,491
        if i > 1:
            gam1x = DAMP * gam1x + (1 - DAMP) * gam1xold           # :494
    n = N // 2
    x = x1[:n] + 1j * x1[n:]                                       # vamp.m:54
    if return_state:
        return x, dict(x1=x1, r1=r1, p1=p1, gam1x=gam1x, gam1z=gam1z)
    return x
