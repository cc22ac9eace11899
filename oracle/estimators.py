"""Oracle restatement of the reference's estimators (test infrastructure - see
``oracle/__init__.py``).  fp64, column-major semantics.

Each solver exists in two modes:
  * ``*_literal``    - follows the MATLAB file line by line, including the
    dense Kronecker operators it materialises (small shapes only);
  * ``*_structured`` - the algebraically identical Kronecker-free form
     used as ground truth at large shapes.
``tests/test_oracle.py`` proves literal == structured.

Files followed (relative to the reference root):
  basic_system_functions/proposed_algorithm.m, proposed_algorithm_angles.m
  benchmark_algorithms/svt.m, mc_svt.m, mc_admm.m, sparse_admm.m, OMP.m
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .matlab_compat import norm2, soft_complex, unvec, vec


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def mpinv(A):
    """MATLAB ``pinv(A)``: SVD with tolerance max(size(A))*eps(norm(A)) (OMP.m:19,
    plot_errorVSsnr.m:83)."""
    A = np.asarray(A)
    if A.size == 0:
        return np.zeros(A.shape[::-1], dtype=A.dtype)
    U, s, Vh = np.linalg.svd(A, full_matrices=False)
    tol = max(A.shape) * np.spacing(s[0]) if s.size else 0.0
    keep = s > tol
    sinv = np.zeros_like(s)
    sinv[keep] = 1.0 / s[keep]
    return (Vh.conj().T * sinv) @ U.conj().T


def _ratio(num, den):
    """MATLAB scalar division: x/0 = Inf, 0/0 = NaN, no exception."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(num) / np.float64(den)


# ----------------------------------------------------------------------------
# svt.m
# ----------------------------------------------------------------------------
def svt_literal(Y, tau):
    """X = svt(Y, tau) (benchmark_algorithms/svt.m:1-15): full svd, soft threshold
    ``max(0,lambda-tau).*lambda./abs(lambda)``; if any lambda is exactly 0 the
    0/0 = NaN trips ``if(~isnan(softThres))`` and the result is all zeros (:7-13)."""
    Y = np.asarray(Y, dtype=np.complex128)
    Mr, Mt = Y.shape
    Uy, lam, Vh = np.linalg.svd(Y, full_matrices=True)           # :5
    with np.errstate(divide="ignore", invalid="ignore"):
        soft = np.maximum(0.0, lam - tau) * lam / np.abs(lam)      # :7
    if not np.any(np.isnan(soft)):                                # :8
        SS = np.zeros((Mr, Mt))
        SS[: soft.size, : soft.size] = np.diag(soft)              # :9
        return Uy @ SS @ Vh                                       # :10  (Vh == Vy')
    return np.zeros((Mr, Mt), dtype=np.complex128)                # :12


def svt_structured(Y, tau):
    """Economy-SVD form of :func:`svt_literal` (output is basis independent)."""
    Y = np.asarray(Y, dtype=np.complex128)
    U, lam, Vh = np.linalg.svd(Y, full_matrices=False)
    if np.any(lam == 0.0):
        return np.zeros_like(Y)
    return (U * np.maximum(0.0, lam - tau)) @ Vh


# ----------------------------------------------------------------------------
# mc_svt.m
# ----------------------------------------------------------------------------
def mc_svt(OH, Omega, Imax, tau, rho, svt=svt_structured):
    """X = mc_svt(OH, Omega, Imax, tau, rho) (benchmark_algorithms/mc_svt.m:1-12)."""
    OH = np.asarray(OH, dtype=np.complex128)
    Y = np.zeros_like(OH)
    X = np.zeros_like(OH)
    for _ in range(int(Imax)):
        X = svt(Y, tau / rho)                                     # :8
        Y = Y + rho * (OH - Omega * X)                            # :9
    return X


# ----------------------------------------------------------------------------
# mc_admm.m
# ----------------------------------------------------------------------------
def _diag_vec_omega_literal(Omega):
    """``sum_i kron(diag(Omega(i,:))', Eii)`` (mc_admm.m:11-16,
    proposed_algorithm.m:14-19): the dense (NM x NM) matrix diag(vec(Omega))."""
    N, M = Omega.shape
    K1 = np.zeros((N * M, N * M))
    for i in range(N):
        Eii = np.zeros((N, N))
        Eii[i, i] = 1.0
        K1 = K1 + np.kron(np.diag(Omega[i, :]).T, Eii)
    return K1


def mc_admm_literal(Htrue, OH, Omega, Imax, tau, rho):
    """[X, conv] = mc_admm(...) (benchmark_algorithms/mc_admm.m:1-34), dense
    ``A\\`` solve each iteration."""
    OH = np.asarray(OH, dtype=np.complex128)
    Mr, Mt = OH.shape
    conv = np.zeros(int(Imax))
    X = np.zeros_like(OH)
    Y = np.zeros_like(OH)
    Z = np.zeros_like(OH)
    A = _diag_vec_omega_literal(Omega) + rho * np.eye(Mr * Mt)    # :11-17
    for i in range(int(Imax)):
        X = svt_literal(Y - Z / rho, tau / rho)                   # :22
        y = np.linalg.solve(A, vec(OH) + vec(Z) + rho * vec(X))   # :24
        Y = unvec(y, Mr, Mt)                                      # :25
        Z = Z + rho * (X - Y)                                     # :26
        conv[i] = _ratio(norm2(X - Htrue) ** 2, norm2(Htrue) ** 2)  # :28
    return X, conv


def mc_admm_structured(Htrue, OH, Omega, Imax, tau, rho):
    OH = np.asarray(OH, dtype=np.complex128)
    conv = np.zeros(int(Imax))
    X = np.zeros_like(OH)
    Y = np.zeros_like(OH)
    Z = np.zeros_like(OH)
    D = 1.0 / (Omega + rho)
    for i in range(int(Imax)):
        X = svt_structured(Y - Z / rho, tau / rho)
        Y = (OH + Z + rho * X) * D
        Z = Z + rho * (X - Y)
        conv[i] = _ratio(norm2(X - Htrue) ** 2, norm2(Htrue) ** 2)
    return X, conv


# ----------------------------------------------------------------------------
# sparse_admm.m
# ----------------------------------------------------------------------------
SPARSE_ADMM_RHO = 0.01        # sparse_admm.m:12
SPARSE_ADMM_TAU = 0.0001      # sparse_admm.m:13


def sparse_admm_literal(Htrue, OH, Dr, Dt, Imax):
    """[S, conv] = sparse_admm(Htrue, OH, Dr, Dt, Imax)
    (benchmark_algorithms/sparse_admm.m:1-36).  Note the ``- rho*I`` sign (:16)
    and that ``reshape(s,Mr,Mt)`` forces Gr*Gt == Mr*Mt (:9,16,23)."""
    OH = np.asarray(OH, dtype=np.complex128)
    Mr, Mt = OH.shape
    Gr, Gt = Dr.shape[1], Dt.shape[1]
    conv = np.zeros(int(Imax))
    Z = np.zeros((Mr, Mt), dtype=np.complex128)
    R = np.zeros((Gr, Gt), dtype=np.complex128)
    rho, tau_s = SPARSE_ADMM_RHO, SPARSE_ADMM_TAU
    A = np.kron(np.conj(Dt), Dr)                                  # :15
    Bm = A.conj().T @ A - rho * np.eye(Mr * Mt)                   # :16
    S = np.zeros((Mr, Mt), dtype=np.complex128)
    for i in range(int(Imax)):
        v = vec(R + Z / rho)                                      # :21
        s = soft_complex(v, tau_s / rho)                          # :22
        S = unvec(s, Mr, Mt)                                      # :23
        r = np.linalg.solve(Bm, vec(Z) - rho * s + A.conj().T @ vec(OH))  # :26
        R = unvec(r, Mr, Mt)                                      # :27
        Z = Z + rho * (R - S)                                     # :30
        conv[i] = _ratio(norm2(Dr @ S @ Dt.conj().T - Htrue) ** 2, norm2(Htrue) ** 2)  # :32
    return S, conv


def sparse_admm_structured(Htrue, OH, Dr, Dt, Imax):
    """Kronecker-free form: A vec(S) = vec(Dr S Dt^H); the solve
    runs in the eigenbases of Dr^H Dr and Dt^H Dt."""
    OH = np.asarray(OH, dtype=np.complex128)
    Mr, Mt = OH.shape
    Gr, Gt = Dr.shape[1], Dt.shape[1]
    assert Gr * Gt == Mr * Mt
    conv = np.zeros(int(Imax))
    rho, tau_s = SPARSE_ADMM_RHO, SPARSE_ADMM_TAU
    lr, Qr = np.linalg.eigh(Dr.conj().T @ Dr)
    lt, Qt = np.linalg.eigh(Dt.conj().T @ Dt)
    # (A^H A) vec(S) = vec(Dr^H Dr S (Dt^H Dt)^T) ; eigenvalues lr_i * lt_j
    den = lr[:, None] * lt[None, :] - rho
    AhOH = Dr.conj().T @ OH @ Dt                                  # A' * vec(OH)
    Z = np.zeros((Mr, Mt), dtype=np.complex128)
    R = np.zeros((Gr, Gt), dtype=np.complex128)
    S = np.zeros((Mr, Mt), dtype=np.complex128)
    for i in range(int(Imax)):
        S = soft_complex(R + Z / rho, tau_s / rho).reshape(Mr, Mt)
        rhs = (Z - rho * S).reshape(Gr, Gt) + AhOH
        R = Qr @ ((Qr.conj().T @ rhs @ np.conj(Qt)) / den) @ Qt.T
        Z = Z + rho * (R.reshape(Mr, Mt) - S)
        conv[i] = _ratio(norm2(Dr @ S.reshape(Gr, Gt) @ Dt.conj().T - Htrue) ** 2, norm2(Htrue) ** 2)
    return S, conv


# ----------------------------------------------------------------------------
# OMP.m
# ----------------------------------------------------------------------------
def omp_literal(A, v, m, snr=None):
    """[x_hat, indexSet, v, targetMatrix] = OMP(A, v, m, snr)
    (benchmark_algorithms/OMP.m:1-32).  ``indexSet`` is returned 1-based; no
    stopping rule; later duplicate picks overwrite ``x_hat`` (:29-31)."""
    A = np.asarray(A, dtype=np.complex128)
    v = np.asarray(v, dtype=np.complex128).reshape(-1)
    measures, size_d = A.shape
    r = v.copy()                                                  # :10
    target = np.zeros((measures, 0), dtype=np.complex128)         # :12
    index_set = []
    x = np.zeros(0, dtype=np.complex128)
    for _t in range(int(m)):                                      # :16
        corr = np.abs(A.conj().T @ r)                             # :17
        idx = int(np.argmax(corr))                                #     first maximum
        index_set.append(idx + 1)
        target = np.concatenate([target, A[:, idx : idx + 1]], axis=1)  # :18
        x = mpinv(target) @ v                                     # :19
        a = target @ x                                            # :20
        r = v - a                                                 # :21
    x_hat = np.zeros(size_d, dtype=np.complex128)                 # :27
    for t, idx1 in enumerate(index_set):                          # :29-31
        x_hat[idx1 - 1] = x[t]
    return x_hat, index_set, v, target


def somp_textbook(A, Y, K, res_tol=0.0):
    """Row-l2 simultaneous OMP standing in for spx.pursuit.joint.OrthogonalMatchingPursuit(A,K).solve(Y)
    (plot_errorVSsnr.m:116-118).  sparse-plex is external and unpinned (README.md:9): PARITY UNPINNED -
    this states the algorithm the CUDA replacement implements, not sparse-plex's source.
    t = 1..K: d = first argmax_d ||A(:,d)' R||_2; Z(support,:) = lstsq(A(:,support), Y); R = Y - A(:,support) Z;
    stops at min(N, D) atoms, on a repeated pick, or when ||R||_F <= res_tol ||Y||_F."""
    A = np.asarray(A, dtype=np.complex128); Y = np.asarray(Y, dtype=np.complex128)
    N, D = A.shape
    R = Y.copy(); support = []
    X = np.zeros((0, Y.shape[1]), dtype=np.complex128)
    ynorm = np.linalg.norm(Y)
    for _t in range(int(K)):
        if len(support) >= min(N, D):
            break
        c = np.sum(np.abs(A.conj().T @ R) ** 2, axis=1)
        d = int(np.argmax(c))
        if d in support:
            break
        support.append(d)
        X = np.linalg.lstsq(A[:, support], Y, rcond=None)[0]
        R = Y - A[:, support] @ X
        if np.linalg.norm(R) <= res_tol * ynorm:
            break
    Z = np.zeros((D, Y.shape[1]), dtype=np.complex128)
    Z[support, :] = X
    return Z, [d + 1 for d in support], R


def omp_kron_structured(A, B, Y, m):
    """OMP(kron(B.', A), vec(Y), m) (OMP.m:1-32 on Phi of plot_errorVSdelays.m:77-78) without forming
    Phi: Phi' r = vec(A' R B') (:17), column j = g + G p is vec(A(:,g) B(p,:)) (:18); pinv re-solve on the
    materialised selected atoms (:19).  Returns (x_hat, indexSet 1-based, x per pick, residual N x M).
    tests/test_oracle.py proves it equal to omp_literal on the materialised Phi."""
    A = np.asarray(A, dtype=np.complex128); B = np.asarray(B, dtype=np.complex128)
    Y = np.asarray(Y, dtype=np.complex128)
    N, G = A.shape; P, M = B.shape
    v = Y.reshape(-1, order="F")
    r = v.copy()
    target = np.zeros((N * M, 0), dtype=np.complex128)
    index_set = []
    x = np.zeros(0, dtype=np.complex128)
    AH, BH = A.conj().T, B.conj().T
    for _t in range(int(m)):
        corr = np.abs((AH @ r.reshape(N, M, order="F") @ BH).reshape(-1, order="F"))   # :17
        idx = int(np.argmax(corr))
        index_set.append(idx + 1)
        g, p = idx % G, idx // G
        atom = np.outer(A[:, g], B[p, :]).reshape(-1, order="F")                          # :18
        target = np.concatenate([target, atom[:, None]], axis=1)
        x = mpinv(target) @ v                                                             # :19
        r = v - target @ x                                                                # :20-21
    x_hat = np.zeros(G * P, dtype=np.complex128)                                          # :27
    for t, idx1 in enumerate(index_set):                                                  # :29-31
        x_hat[idx1 - 1] = x[t]
    return x_hat, index_set, x, r.reshape(N, M, order="F")


# ----------------------------------------------------------------------------
# proposed_algorithm.m / proposed_algorithm_angles.m
# ----------------------------------------------------------------------------
def _k3_literal(Omega_S):
    """proposed_algorithm_angles.m:37-43 - dense diag(vec(Omega_S))."""
    G = Omega_S.shape[0]
    K3 = np.zeros((Omega_S.size, Omega_S.size))
    for ii in range(G):
        Eii = np.zeros((G, G))
        Eii[ii, ii] = 1.0
        K3 = K3 + np.kron(np.diag(Omega_S[ii, :]).T, Eii)
    return K3


def proposed_algorithm_literal(subY, Omega, A, B, Imax, tau_Y, tau_S, rho, type_,
                               indx_S=None):
    """[S, Y, conv] = proposed_algorithm(subY, Omega, A, B, Imax, tau_Y, tau_S, rho, type)
    (basic_system_functions/proposed_algorithm.m:1-73); with ``indx_S`` (1-based)
    it is proposed_algorithm_angles.m:1-85.  Materialises K1, K2, R (or lu(K2))
    exactly like the reference - small shapes only."""
    subY = np.asarray(subY, dtype=np.complex128)
    N, M = subY.shape
    Gr, Gt = A.shape[1], B.shape[0]
    Imax = int(Imax)
    conv = np.zeros((Imax, 3))
    X = np.zeros((N, M), dtype=np.complex128)
    V1 = np.zeros_like(X)
    V2 = np.zeros_like(X)
    C = np.zeros_like(X)
    s = np.zeros(Gr * Gt, dtype=np.complex128)
    K1 = _diag_vec_omega_literal(Omega)                           # :14-19
    iK1 = 1.0 / np.diag(K1 + 2 * rho * np.eye(N * M))             # :20 (sparse diag)
    K2 = np.kron(B.T, A)                                          # :22
    approx = type_ == "approximate"
    if approx:
        R = K2.conj().T @ K2                                      # :25
        v = np.zeros(R.shape[1], dtype=np.complex128)             # :27
    else:
        Pm, Ll, Uu = sla.lu(K2)                                   # :29  [L,U]=lu(K2) => L = P*L
        Lp = Pm @ Ll
        v = np.zeros(Gr * Gt, dtype=np.complex128)
    Omega_S = np.zeros((Gr, Gt))
    Y = np.zeros_like(X)
    S = np.zeros((Gr, Gt), dtype=np.complex128)
    for i in range(1, Imax + 1):
        if indx_S is not None:                                    # angles :36-43 (cumulative)
            cnt = min(10 + 5 * i, Gt * Gr)
            flat = vec(Omega_S).copy()
            flat[np.asarray(indx_S[:cnt], dtype=int) - 1] = 1.0
            Omega_S = unvec(flat, Gr, Gt)
            K3 = _k3_literal(Omega_S)
        Y = svt_literal(X - V1 / rho, tau_Y / rho)                # :35
        b = vec(V1) + rho * vec(Y) + vec(subY) + vec(V2) + rho * vec(C) + rho * (K2 @ s)  # :38
        x = iK1 * b                                               # :39
        X = unvec(x, N, M)                                        # :40
        k = vec(X) - vec(V2) / rho - vec(C)                       # :43
        if approx:
            res = K2.conj().T @ k - R @ v                         # :47
            with np.errstate(divide="ignore", invalid="ignore"):
                alpha = np.vdot(res, res) / np.vdot(res, R @ res) # :48
            prev_v = v                                            # :49
            v = v + alpha * res                                   # :50
            conv[i - 1, 2] = _ratio(np.linalg.norm(prev_v - v) ** 2, np.linalg.norm(prev_v) ** 2)  # :51
        else:
            w = np.linalg.lstsq(Lp, k, rcond=None)[0]             # :53  L\k (rectangular => LS)
            v = sla.solve_triangular(Uu, w)                       #      U\(...)
        s = soft_complex(v, tau_S / rho)                          # :56
        if indx_S is not None:
            s = K3 @ s                                            # angles :68
        S = unvec(s, Gr, Gt)                                      # :57
        Xs = A @ S @ B                                            # :58
        C = rho / (rho + 1) * (X - Xs - V2 / rho)                 # :61
        V1 = V1 + rho * (Y - X)                                   # :64
        V2 = V2 + rho * (C - X + Xs)                              # :65
        conv[i - 1, 0] = _ratio(norm2(V1) ** 2, norm2(X) ** 2)    # :67
        conv[i - 1, 1] = _ratio(norm2(V2) ** 2, norm2(X) ** 2)    # :69
    return S, Y, conv


def proposed_algorithm_structured(subY, Omega, A, B, Imax, tau_Y, tau_S, rho, type_,
                                  indx_S=None, want_conv=True, trace=None):
    """Kronecker-free form of proposed_algorithm(.m) / _angles:
    K1 = diag(vec Omega); K2 s = vec(A S B); K2'k = vec(A^H K B^H);
    R v = vec(A^H A V B B^H); 'std' => V = pinv(A) K pinv(B)."""
    subY = np.asarray(subY, dtype=np.complex128)
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.complex128)
    N, M = subY.shape
    G, P = A.shape[1], B.shape[0]
    Imax = int(Imax)
    conv = np.zeros((Imax, 3))
    X = np.zeros((N, M), dtype=np.complex128)
    V1 = np.zeros_like(X)
    V2 = np.zeros_like(X)
    C = np.zeros_like(X)
    Xs = np.zeros_like(X)
    V = np.zeros((G, P), dtype=np.complex128)
    S = np.zeros((G, P), dtype=np.complex128)
    Y = np.zeros_like(X)
    D = 1.0 / (Omega + 2.0 * rho)
    AH = A.conj().T
    BH = B.conj().T
    approx = type_ == "approximate"
    if approx:
        AHA = AH @ A
        BBH = B @ BH
    else:
        pA = np.linalg.pinv(A)
        pB = np.linalg.pinv(B)
    mask_flat = np.zeros(G * P)
    for i in range(1, Imax + 1):
        if indx_S is not None:
            cnt = min(10 + 5 * i, G * P)
            mask_flat[np.asarray(indx_S[:cnt], dtype=int) - 1] = 1.0
        Y = svt_structured(X - V1 / rho, tau_Y / rho)
        X = (V1 + rho * Y + subY + V2 + rho * C + rho * Xs) * D
        Kt = X - V2 / rho - C
        if approx:
            Res = AH @ Kt @ BH - AHA @ V @ BBH
            Q = AHA @ Res @ BBH
            with np.errstate(divide="ignore", invalid="ignore"):
                alpha = np.float64(np.vdot(Res, Res).real) / np.float64(np.vdot(Res, Q).real)
            Vprev = V
            V = V + alpha * Res
            if want_conv:
                conv[i - 1, 2] = _ratio(np.linalg.norm(Vprev - V) ** 2, np.linalg.norm(Vprev) ** 2)
        else:
            V = pA @ Kt @ pB
        S = soft_complex(V, tau_S / rho)
        if indx_S is not None:
            S = S * unvec(mask_flat, G, P)
        Xs = A @ S @ B
        C = rho / (rho + 1) * (X - Xs - V2 / rho)
        V1 = V1 + rho * (Y - X)
        V2 = V2 + rho * (C - X + Xs)
        if want_conv:
            nx = norm2(X) ** 2
            conv[i - 1, 0] = _ratio(norm2(V1) ** 2, nx)
            conv[i - 1, 1] = _ratio(norm2(V2) ** 2, nx)
        if trace is not None:
            trace.append(dict(Y=Y.copy(), X=X.copy(), V=V.copy(), S=S.copy()))
    return S, Y, conv


# ----------------------------------------------------------------------------
# driver-side parameters and metrics (plot_errorVSsnr.m:127-130,138-141)
# ----------------------------------------------------------------------------
def admm_parameters(Y_hbf, Zbar, rho_rule="sigma6"):
    """tau_Y = 1/||Y||_F^2 ; tau_Z = 1/||Zbar||_F^2/2 ;
    rho = sqrt(min(eigs(Y'Y)) / ||Y||_F^2)  (plot_errorVSsnr.m:127-130) or with
    max(eigs) (plot_errorVSdelays.m:127-128, rho_rule='sigma1')."""
    from .matlab_compat import eigs6, fro
    fy = fro(Y_hbf) ** 2
    tau_Y = 1.0 / fy
    tau_Z = 1.0 / fro(Zbar) ** 2 / 2.0
    # eigs(Y'Y): the non-zero eigenvalues equal those of the small Gram Y Y'
    Ysm = Y_hbf if Y_hbf.shape[0] <= Y_hbf.shape[1] else Y_hbf.conj().T
    ev = eigs6(Ysm @ Ysm.conj().T)
    if ev.size < 6 and min(Y_hbf.shape) < 6 <= max(Y_hbf.shape):
        ev = np.concatenate([ev, np.zeros(6 - ev.size)])
    lam = float(np.min(ev)) if rho_rule == "sigma6" else float(np.max(ev))
    rho = float(np.sqrt(max(lam, 0.0) / fy))
    return tau_Y, tau_Z, rho


def nmse(S, Zbar):
    """``norm(S-Zbar)^2/norm(Zbar)^2`` with matrix 2-norms, clipped at 1
    (plot_errorVSsnr.m:138-141)."""
    e = _ratio(norm2(S - Zbar) ** 2, norm2(Zbar) ** 2)
    return float(min(e, 1.0)) if not np.isnan(e) else float("nan")


def log2det_rate(X, scale):
    """real(log2(det(eye(n) + scale*X*X'))) exactly as the sweep drivers write it
    (plot_rateVSframelength.m:113,130,135 with X = Zbar, scale = 1/(Nr*(sigma2+nmse));
    plot_capacity.m:47-66 with X = W_c'*Y, scale = 1/(sigma2*Nt))."""
    X = np.asarray(X, dtype=np.complex128)
    n = X.shape[0]
    return float(np.real(np.log2(np.linalg.det(np.eye(n) + scale * (X @ X.conj().T)))))


def ls_estimate(A, Y, B):
    """S_ls = pinv(A)*Y*pinv(B)  (plot_errorVSsnr.m:83; plot_errorVSsnr_approx.m:61,67 with Y = the estimator's second output)."""
    return np.linalg.pinv(np.asarray(A, complex)) @ np.asarray(Y, complex) @ np.linalg.pinv(np.asarray(B, complex))


def y_pinv_b(Y, B):
    """Y_hbf_nr*pinv(B): the right-hand sides of the joint OMP call (plot_errorVSsnr.m:117)."""
    return np.asarray(Y, complex) @ np.linalg.pinv(np.asarray(B, complex))


def capacity_literal(Y, W, cols, scale):
    """real(log2(det(eye(Mr) + scale*W(:,cols)'*(Y*Y')*W(:,cols))))  (plot_capacity.m:47,52,57,64; plot_ee.m:47,52,57,64).
    ``cols`` 1-based, like ind(1:Mr) of ind = randperm(Mr_e) (plot_capacity.m:63) or 1:Lr (hbf.m:24)."""
    Y = np.asarray(Y, complex); W = np.asarray(W, complex)
    Ws = W[:, np.asarray(cols, int) - 1]
    Mr = Ws.shape[1]
    return float(np.real(np.log2(np.linalg.det(np.eye(Mr) + scale * (Ws.conj().T @ (Y @ Y.conj().T) @ Ws)))))


def ee_power_model(Nr, Mr, Mr_e):
    """power_dbf, power_hbf, power_hbf_zc, power_proposed  (plot_ee.m:69-77)."""
    Pcirc, Psw, Pps, Plna, Pps_zc = 0, 0.005, 0.015, 0.02, 0.06                  # plot_ee.m:69-73
    return (Pcirc + Nr * Nr * Plna + Nr * (Nr + 1) * Pps_zc,                       # :74
            Pcirc + Mr * Nr * Plna + Nr * (Mr + 1) * Pps,                          # :75
            Pcirc + Mr * Nr * Plna + Nr * (Mr + 1) * Pps_zc,                       # :76
            Pcirc + Mr_e * Nr * Plna + Mr_e * Psw + Nr * (Mr_e + 1) * Pps)         # :77
