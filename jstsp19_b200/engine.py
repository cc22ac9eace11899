"""Device-resident batched engine: torch tensors in, torch tensors out, all compute in
``libjstsp_b200.so`` on torch's current CUDA stream (torch is used for device memory,
streams and ``torch.distributed`` only).

Tensor layout everywhere: ``(batch, cols, rows)`` C-contiguous = per-trial column-major,
complex64 (``precision="f32"``) or complex128 (``"f64"``).

Multi-GPU: Monte-Carlo trials are independent (the reference uses ``parfor`` over them,
plot_errorVSdelays.m:51), so ranks take contiguous trial ranges with NO data-path
collective; ``MonteCarlo.reduce`` is the single NCCL all-reduce of the NMSE partial sums.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import AdmmDesc, Handle, MeasDesc

_CD = {"f32": torch.complex64, "f64": torch.complex128}
_RD = {"f32": torch.float32, "f64": torch.float64}
_DT = {"f32": _lib.F32, "f64": _lib.F64}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class AdmmEngine:
    def __init__(self, device=0, precision="f32", max_trials_per_pass=0):
        if not torch.cuda.is_available():
            raise RuntimeError("jstsp19_b200.engine needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", device)
        self.precision = precision
        self.h = Handle(device)
        if max_trials_per_pass:
            self.h.set_chunk(max_trials_per_pass)

    def _bind_stream(self):
        self.h.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def proposed_algorithm(self, subY, Omega, A, B, imax, tau_Y, tau_S, rho, type="approximate", indx_S=None,
                           S_out=None, Y_out=None):
        """Batched proposed_algorithm / proposed_algorithm_angles on device tensors.
        subY (b,M,N), Omega (b,M,N) real, A (b|1,G,N), B (b|1,M,P), tau_Y/tau_S/rho (b,) float64.
        Returns S (b,P,G) [and fills Y_out (b,M,N) when given]; asynchronous on the current stream."""
        cd, rdt = _CD[self.precision], _RD[self.precision]
        b, M, N = subY.shape
        G, P = A.shape[-2], B.shape[-1]
        for t, dt in ((subY, cd), (A, cd), (B, cd), (Omega, rdt), (tau_Y, torch.float64), (tau_S, torch.float64), (rho, torch.float64)):
            if t.dtype != dt or not t.is_contiguous() or t.device != self.device:
                raise ValueError("engine tensors must be contiguous, on the engine's device, and of the engine's precision")
        if S_out is None:
            S_out = torch.empty(b, P, G, dtype=cd, device=self.device)
        d = AdmmDesc()
        d.N, d.M, d.G, d.P, d.imax, d.batch = N, M, G, P, int(imax), b
        d.type = _lib.APPROXIMATE if type == "approximate" else _lib.STD
        d.ld_subY = N * M
        for t in (Omega, A, B):
            if t.shape[0] not in (1, b):
                raise ValueError("leading dimension must be 1 (shared by all trials) or batch")
        d.ld_omega = N * M if Omega.shape[0] == b else 0
        d.ld_A = N * G if A.shape[0] == b else 0
        d.ld_B = P * M if B.shape[0] == b else 0
        d.ld_S, d.ld_Y, d.ld_conv = G * P, N * M, 0
        self._bind_stream()
        if indx_S is None:
            rc = _lib.lib.jstsp_proposed_algorithm(self.h.ptr, C.byref(d), _DT[self.precision], _lib.DEVICE, _p(subY), _p(Omega), _p(A), _p(B),
                                                   _p(tau_Y), _p(tau_S), _p(rho), _p(S_out), _p(Y_out), None)
        else:
            d.n_indx = indx_S.shape[-1]
            d.ld_indx = indx_S.shape[-1] if indx_S.dim() == 2 and indx_S.shape[0] == b else 0
            rc = _lib.lib.jstsp_proposed_algorithm_angles(self.h.ptr, C.byref(d), _DT[self.precision], _lib.DEVICE, _p(subY), _p(Omega),
                                                          _p(indx_S), _p(A), _p(B), _p(tau_Y), _p(tau_S), _p(rho), _p(S_out), _p(Y_out), None)
        self.h.check(rc)
        return S_out

    def proposed_algorithm_psi(self, subY, Omega, A, Dt, Psi, imax, tau_Y, tau_S, rho, type="approximate", indx_S=None,
                               S_out=None, Y_out=None):
        """Same estimator with the dictionary given by its factors (jstsp_proposed_algorithm_psi):
        Dt (b|1,Gt,Nt) and Psi_bar (b|1,L,M,Nt) in per-trial column-major storage; B = (I (x) Dt') Psi_bar is never an input."""
        cd, rdt = _CD[self.precision], _RD[self.precision]
        b, M, N = subY.shape
        G = A.shape[-2]
        Gt, Nt = Dt.shape[-2], Dt.shape[-1]
        L = Psi.shape[-3]
        P = L * Gt
        for t, dt in ((subY, cd), (A, cd), (Dt, cd), (Psi, cd), (Omega, rdt), (tau_Y, torch.float64), (tau_S, torch.float64), (rho, torch.float64)):
            if t.dtype != dt or not t.is_contiguous() or t.device != self.device:
                raise ValueError("engine tensors must be contiguous, on the engine's device, and of the engine's precision")
        if Psi.shape[-1] != Nt or Psi.shape[-2] != M:
            raise ValueError("Psi_bar must be (b|1, L, M, Nt)")
        if S_out is None:
            S_out = torch.empty(b, P, G, dtype=cd, device=self.device)
        d = AdmmDesc()
        d.N, d.M, d.G, d.P, d.imax, d.batch = N, M, G, P, int(imax), b
        d.type = _lib.APPROXIMATE if type == "approximate" else _lib.STD
        d.ld_subY = N * M
        d.ld_omega = N * M if Omega.shape[0] == b else 0
        d.ld_A = N * G if A.shape[0] == b else 0
        d.ld_S, d.ld_Y, d.ld_conv = G * P, N * M, 0
        if indx_S is not None:
            d.n_indx = indx_S.shape[-1]
            d.ld_indx = indx_S.shape[-1] if indx_S.dim() == 2 and indx_S.shape[0] == b else 0
        self._bind_stream()
        rc = _lib.lib.jstsp_proposed_algorithm_psi(self.h.ptr, C.byref(d), _DT[self.precision], _lib.DEVICE, _p(subY), _p(Omega), _p(indx_S), _p(A),
                                                   _p(Dt), Nt * Gt if (Dt.dim() == 3 and Dt.shape[0] == b and b > 1) else 0,
                                                   _p(Psi), Nt * M * L if (Psi.dim() == 4 and Psi.shape[0] == b and b > 1) else 0, Nt, L,
                                                   _p(tau_Y), _p(tau_S), _p(rho), _p(S_out), _p(Y_out), None)
        self.h.check(rc)
        return S_out

    def proposed_algorithm_pilots(self, subY, Omega, A, Dt, pilots, L, imax, tau_Y, tau_S, rho, type="approximate", indx_S=None, S_out=None, Y_out=None):
        """:meth:`proposed_algorithm_psi` fed the pilot sequences (jstsp_proposed_algorithm_pilots): ``pilots`` (b|1, M, Nt) = per-trial
        column-major Nt x M with row k = s_k; Psi_bar is expanded inside the library."""
        cd, rdt = _CD[self.precision], _RD[self.precision]
        b, M, N = subY.shape
        G = A.shape[-2]
        Gt, Nt = Dt.shape[-2], Dt.shape[-1]
        P = int(L) * Gt
        for t, dt in ((subY, cd), (A, cd), (Dt, cd), (pilots, cd), (Omega, rdt), (tau_Y, torch.float64), (tau_S, torch.float64), (rho, torch.float64)):
            if t.dtype != dt or not t.is_contiguous() or t.device != self.device:
                raise ValueError("engine tensors must be contiguous, on the engine's device, and of the engine's precision")
        if pilots.shape[-1] != Nt or pilots.shape[-2] != M:
            raise ValueError("pilots must be (b|1, M, Nt)")
        if S_out is None:
            S_out = torch.empty(b, P, G, dtype=cd, device=self.device)
        d = AdmmDesc()
        d.N, d.M, d.G, d.P, d.imax, d.batch = N, M, G, P, int(imax), b
        d.type = _lib.APPROXIMATE if type == "approximate" else _lib.STD
        d.ld_subY = N * M
        d.ld_omega = N * M if Omega.shape[0] == b else 0
        d.ld_A = N * G if A.shape[0] == b else 0
        d.ld_S, d.ld_Y, d.ld_conv = G * P, N * M, 0
        if indx_S is not None:
            d.n_indx = indx_S.shape[-1]
            d.ld_indx = indx_S.shape[-1] if indx_S.dim() == 2 and indx_S.shape[0] == b else 0
        self._bind_stream()
        rc = _lib.lib.jstsp_proposed_algorithm_pilots(self.h.ptr, C.byref(d), _DT[self.precision], _lib.DEVICE, _p(subY), _p(Omega), _p(indx_S), _p(A),
                                                      _p(Dt), Nt * Gt if (Dt.dim() == 3 and Dt.shape[0] == b and b > 1) else 0,
                                                      _p(pilots), Nt * M if (pilots.dim() == 3 and pilots.shape[0] == b and b > 1) else 0, Nt, int(L),
                                                      _p(tau_Y), _p(tau_S), _p(rho), _p(S_out), _p(Y_out), None)
        self.h.check(rc)
        return S_out

    @property
    def launches(self):
        return self.h.launches


class TrialPipeline:
    """The body of the reference's Monte-Carlo loop (plot_errorVSsnr.m:56-67,124-141) for a batch of trials, device-resident and
    entirely through the library's C ABI: wideband_mmwave_channel -> proposed_hbf (pilots, noise, ZC combiner, sampling mask) ->
    tau_Y / tau_Z / rho -> proposed_algorithm -> NMSE.  torch only draws the random numbers and holds the buffers."""

    def __init__(self, shape, device=0, precision="f32", engine=None):
        self.s = shape
        self.eng = engine or AdmmEngine(device, precision)
        self.device, self.precision = self.eng.device, self.eng.precision
        from . import synth
        cd = _CD[self.precision]
        s = shape
        W = synth._zc(s.Nr, self.device, torch.complex128)                                   # createBeamformer(Nr,'ZC') (plot_errorVSsnr.m:124)
        Dr = synth._dft(s.Nr, s.Nr, self.device, torch.complex128)
        self.W = W.T.contiguous().to(cd)[None]                                                # (1, Nr, Nr) column-major
        self.A = (W.conj().T @ Dr).T.contiguous().to(cd)[None]                                # A = W_e' Dr (:132), W_e = W (Mr_e = Nr)
        self.Dt = synth._dft(s.Nt, s.Nt, self.device, torch.complex128).T.contiguous().to(cd)[None]

    def run_from_draws(self, coef, u_r, u_t, noise_unit, sym_idx, mask_rank, sigma2, imax=100, keep=False):
        """The trial-loop body on explicit draws (synth.draw's tensors; used by the parity tests, which hand the same draws to the fp64 restatement)."""
        s, dev = self.s, self.device
        cd = _CD[self.precision]
        # draws in the reference's order (wideband_mmwave_channel.m:19-22): (randn, randn) and (rand, rand) per (tap, ray)
        normals = (torch.view_as_real(coef.to(torch.complex128)) * (2.0 ** 0.5)).contiguous()           # (b, L, Np, 2)
        uniforms = torch.stack([u_r.double(), u_t.double()], dim=-1).contiguous()
        qam = torch.tensor([1 + 1j, -1 + 1j, 1 - 1j, -1 - 1j], dtype=torch.complex128, device=dev) / (2.0 ** 0.5)   # qam4mod.m:7-8
        pilots = qam[sym_idx].transpose(1, 2).contiguous().to(cd)                                        # (b, M, Nt): row k = s_k
        noise = (noise_unit.to(torch.complex128) * torch.sqrt(sigma2.double())[:, None, None]).transpose(1, 2).contiguous().to(cd)   # :60
        perm = (mask_rank.argsort(dim=1).transpose(1, 2) + 1).to(torch.int32).contiguous()               # (b, M, Nr): rows in sampling order
        return self._body(normals, uniforms, pilots, noise, perm, imax, keep)

    def device_draws(self, batch, snr_db, seed, first_trial=0):
        """Draws of the global trials [first_trial, first_trial + batch) from the library's Philox4x32-10 generator (csrc/rng.cu), keyed by
        (seed, global trial): normals, uniforms, pilots, noise, perm in the layouts jstsp_wideband_mmwave_channel / jstsp_measure read, and sigma2."""
        s, dev, h = self.s, self.device, self.eng.h
        cd, dt = _CD[self.precision], _DT[self.precision]
        b, Np = int(batch), s.ncl * s.nray
        snr = torch.as_tensor(snr_db, dtype=torch.float64).to(dev).expand(b) if not torch.is_tensor(snr_db) else snr_db.to(dev).double()
        sigma2 = (10.0 ** (-snr / 10.0)).contiguous()                                                    # plot_errorVSsnr.m:49
        normals = torch.empty(b, s.L, Np, 2, dtype=torch.float64, device=dev); uniforms = torch.empty_like(normals)
        pilots = torch.empty(b, s.M, s.Nt, dtype=cd, device=dev); noise = torch.empty(b, s.M, s.Nr, dtype=cd, device=dev)
        perm = torch.empty(b, s.M, s.Nr, dtype=torch.int32, device=dev)
        self.eng._bind_stream()
        h.check(_lib.lib.jstsp_draw_trials(h.ptr, dt, int(seed), int(first_trial), b, s.Nr, s.Nt, s.L, Np, s.M, _p(sigma2), _p(normals), _p(uniforms), _p(pilots), _p(noise), _p(perm)))
        return normals, uniforms, pilots, noise, perm, sigma2

    def _body(self, normals, uniforms, pilots, noise, perm, imax, keep):
        s, dev, h = self.s, self.device, self.eng.h
        cd, rd, dt = _CD[self.precision], _RD[self.precision], _DT[self.precision]
        b = normals.shape[0]
        Nr, Nt, L, M, P = s.Nr, s.Nt, s.L, s.M, s.P
        self.eng._bind_stream()
        lib = _lib.lib
        H = torch.empty(b, L, Nt, Nr, dtype=cd, device=dev)
        Zbar = torch.empty(b, P, Nr, dtype=cd, device=dev)
        h.check(lib.jstsp_wideband_mmwave_channel(h.ptr, dt, _lib.DEVICE, L, Nr, Nt, s.ncl, s.nray, Nr, Nt, b, _p(normals), _p(uniforms),
                                                  _p(H), _p(Zbar), None, None, None, None))
        md = MeasDesc()
        md.Nr, md.Nt, md.L, md.T, md.Wc, md.Lr, md.psi_mode, md.Tp, md.batch = Nr, Nt, L, M, Nr, s.Mr, 1, M, b
        md.ld_H, md.ld_N, md.ld_Psi, md.ld_W = Nr * Nt * L, Nr * M, Nt * M, 0
        subY = torch.empty(b, M, Nr, dtype=cd, device=dev)
        Omega = torch.empty(b, M, Nr, dtype=rd, device=dev)
        h.check(lib.jstsp_measure(h.ptr, C.byref(md), dt, _lib.DEVICE, _p(H), _p(noise), _p(pilots), _p(self.W), _p(perm),
                                  _p(subY), None, None, _p(Omega), None))                               # proposed_hbf.m:14-42
        tau_Y, tau_Z, rho = (torch.empty(b, dtype=torch.float64, device=dev) for _ in range(3))
        h.check(lib.jstsp_admm_parameters(h.ptr, dt, _lib.DEVICE, Nr, M, Nr, P, b, 6, _p(subY), Nr * M, _p(Zbar), Nr * P, _p(tau_Y), _p(tau_Z), _p(rho)))
        S = self.eng.proposed_algorithm_pilots(subY, Omega, self.A, self.Dt, pilots, L, imax, tau_Y, tau_Z, rho, "approximate")   # :137
        nm = torch.empty(b, dtype=torch.float64, device=dev)
        self.eng._bind_stream()
        h.check(lib.jstsp_nmse(h.ptr, dt, _lib.DEVICE, Nr, P, b, _p(S), Nr * P, _p(Zbar), Nr * P, _p(nm)))    # :138-141
        if keep:
            return dict(nmse=nm, S=S, Zbar=Zbar, subY=subY, Omega=Omega, H=H, tau_Y=tau_Y, tau_Z=tau_Z, rho=rho, pilots=pilots,
                        normals=normals, uniforms=uniforms, noise=noise, perm=perm)
        return nm

    def run(self, batch, snr_db, seed, first_trial=0, imax=100, keep=False):
        """One batch of Monte-Carlo trials, draws included, entirely through the library (no torch arithmetic between the calls)."""
        normals, uniforms, pilots, noise, perm, _ = self.device_draws(batch, snr_db, seed, first_trial)
        return self._body(normals, uniforms, pilots, noise, perm, imax, keep)


def shard_range(n_trials, rank, world):
    """Contiguous trial range [lo, hi) of `rank` (ranks differ by at most one trial)."""
    base, rem = divmod(n_trials, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class MonteCarlo:
    """Accumulates per-trial NMSE on a rank and reduces [sum, count, flagged] over ranks -
    the only cross-rank exchange of a sweep point (mean over trials, plot_errorVSsnr.m:170-178)."""

    def __init__(self, device):
        self.acc = torch.zeros(3, dtype=torch.float64, device=device)

    def add(self, nmse):
        bad = ~torch.isfinite(nmse)
        self.acc[0] += torch.where(bad, torch.zeros_like(nmse), nmse).sum()
        self.acc[1] += (~bad).sum()
        self.acc[2] += bad.sum()

    def reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        s, n, f = (float(x) for x in self.acc.tolist())
        return dict(mean_nmse=s / n if n else float("nan"), trials=int(n), flagged=int(f))
