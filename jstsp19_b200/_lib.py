"""ctypes binding of ``libjstsp_b200.so`` (the C ABI in ``include/jstsp_b200.h``).

There is deliberately no CPU fallback: importing works without a GPU (so the
symbol table can be checked on a CPU box), but creating a handle raises unless a
sm_100-class device is present, and a missing shared library raises at import.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JSTSP_LIB") or os.path.join(_HERE, "libjstsp_b200.so")      # JSTSP_LIB: another build of the same library (the sanitizer variant of `make racecheck-lib`)

F32, F64 = 0, 1
HOST, DEVICE = 0, 1
APPROXIMATE, STD = 0, 1

E_ARG, E_CUDA, E_UNSUPPORTED, E_NOMEM = -1, -2, -3, -4


class JstspError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"jstsp error {code}: {msg}")
        self.code = code


class AdmmDesc(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("M", C.c_int), ("G", C.c_int), ("P", C.c_int),
        ("imax", C.c_int), ("type", C.c_int), ("batch", C.c_int),
        ("ld_subY", C.c_longlong), ("ld_omega", C.c_longlong), ("ld_A", C.c_longlong), ("ld_B", C.c_longlong),
        ("ld_S", C.c_longlong), ("ld_Y", C.c_longlong), ("ld_conv", C.c_longlong),
        ("n_indx", C.c_int), ("ld_indx", C.c_longlong),
    ]


class MeasDesc(C.Structure):
    _fields_ = [("Nr", C.c_int), ("Nt", C.c_int), ("L", C.c_int), ("T", C.c_int), ("Wc", C.c_int), ("Lr", C.c_int),
                ("psi_mode", C.c_int), ("Tp", C.c_int), ("batch", C.c_int),
                ("ld_H", C.c_longlong), ("ld_N", C.c_longlong), ("ld_Psi", C.c_longlong), ("ld_W", C.c_longlong)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing - build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`); "
            "jstsp19_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i, ll, dp = C.c_void_p, C.c_int, C.c_longlong, C.POINTER(C.c_double)
    lib.jstsp_version.restype = C.c_char_p
    lib.jstsp_create.argtypes = [C.POINTER(vp), i]
    lib.jstsp_destroy.argtypes = [vp]
    lib.jstsp_destroy.restype = None
    lib.jstsp_last_error.argtypes = [vp]
    lib.jstsp_last_error.restype = C.c_char_p
    lib.jstsp_set_stream.argtypes = [vp, vp]
    lib.jstsp_synchronize.argtypes = [vp]
    lib.jstsp_launch_count.argtypes = [vp]
    lib.jstsp_launch_count.restype = ll
    lib.jstsp_set_chunk.argtypes = [vp, i]
    lib.jstsp_debug_buffer.argtypes = [vp, vp]
    lib.jstsp_profile.argtypes = [vp, i]
    lib.jstsp_profile_read.argtypes = [vp, i, dp, C.POINTER(ll), C.POINTER(C.c_char_p)]
    lib.jstsp_proposed_algorithm.argtypes = [vp, C.POINTER(AdmmDesc), i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.jstsp_proposed_algorithm_angles.argtypes = [vp, C.POINTER(AdmmDesc), i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.jstsp_proposed_algorithm_psi.argtypes = [vp, C.POINTER(AdmmDesc), i, i, vp, vp, vp, vp, vp, ll, vp, ll, i, i, vp, vp, vp, vp, vp, vp]
    lib.jstsp_proposed_algorithm_pilots.argtypes = [vp, C.POINTER(AdmmDesc), i, i, vp, vp, vp, vp, vp, ll, vp, ll, i, i, vp, vp, vp, vp, vp, vp]
    lib.jstsp_last_path.argtypes = [vp]
    lib.jstsp_last_variant.argtypes = [vp]
    lib.jstsp_nonfinite_count.argtypes = [vp]
    lib.jstsp_nonfinite_count.restype = ll
    lib.jstsp_ls_estimate.argtypes = [vp, i, i, i, i, i, i, i, vp, ll, vp, ll, vp, ll, vp, ll, vp, ll]
    lib.jstsp_capacity.argtypes = [vp, i, i, i, i, i, i, i, vp, ll, vp, ll, vp, ll, vp, vp]
    lib.jstsp_power_model.argtypes = [i, i, i, vp]
    lib.jstsp_capacity_sweep.argtypes = [vp, i, i, i, i, i, i, vp, vp, ll, vp, vp, vp, ll, vp, vp]
    lib.jstsp_draw_trials.argtypes = [vp, i, C.c_ulonglong, ll, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp]
    lib.jstsp_philox4x32_10.argtypes = [vp, vp, vp]
    lib.jstsp_philox4x32_10.restype = None
    lib.jstsp_svt.argtypes = [vp, i, i, i, i, i, vp, ll, vp, vp, ll]
    lib.jstsp_mc_svt.argtypes = [vp, i, i, i, i, i, i, vp, ll, vp, ll, vp, vp, vp, ll]
    lib.jstsp_omp.argtypes = [vp, i, i, i, i, i, i, vp, ll, vp, ll, vp, ll, vp, vp, vp, C.c_double]
    lib.jstsp_omp_kron.argtypes = [vp, i, i, i, i, i, i, i, i, vp, ll, vp, ll, vp, ll, vp, ll, vp, vp, vp, ll, vp, C.c_double]
    lib.jstsp_somp.argtypes = [vp, i, i, i, i, i, i, i, vp, ll, vp, ll, vp, ll, vp, vp, vp, ll, C.c_double]
    lib.jstsp_sparse_admm.argtypes = [vp, i, i, i, i, i, i, vp, ll, vp, ll, vp, ll, vp, ll, vp, ll, vp, ll]
    lib.jstsp_vamp.argtypes = [vp, i, i, i, i, i, i, C.c_double, vp, ll, vp, ll, vp, vp, vp, ll, vp, ll, vp, ll]
    lib.jstsp_wideband_mmwave_channel.argtypes = [vp, i, i, i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.jstsp_measure.argtypes = [vp, C.POINTER(MeasDesc), i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.jstsp_create_beamformer.argtypes = [vp, i, i, i, i, vp, vp]
    lib.jstsp_qam4mod.argtypes = [vp, i, i, i, ll, vp, vp, vp]
    lib.jstsp_nmse.argtypes = [vp, i, i, i, i, i, vp, ll, vp, ll, vp]
    lib.jstsp_admm_parameters.argtypes = [vp, i, i, i, i, i, i, i, i, vp, ll, vp, ll, vp, vp, vp]
    lib.jstsp_log2det_rate.argtypes = [vp, i, i, i, i, i, vp, ll, vp, vp]
    lib.jstsp_mc_admm.argtypes = [vp, i, i, i, i, i, i, vp, ll, vp, ll, vp, ll, vp, vp, vp, ll, vp, ll]
    return lib


lib = _load()

#: every symbol include/jstsp_b200.h declares (checked by tests/test_abi.py)
EXPORTED = [
    "jstsp_create", "jstsp_destroy", "jstsp_last_error", "jstsp_version", "jstsp_set_stream",
    "jstsp_synchronize", "jstsp_launch_count", "jstsp_set_chunk", "jstsp_profile", "jstsp_profile_read", "jstsp_debug_buffer",
    "jstsp_proposed_algorithm", "jstsp_proposed_algorithm_angles", "jstsp_proposed_algorithm_psi", "jstsp_proposed_algorithm_pilots", "jstsp_last_path", "jstsp_last_variant", "jstsp_nonfinite_count", "jstsp_ls_estimate", "jstsp_capacity", "jstsp_power_model",
    "jstsp_svt", "jstsp_mc_svt", "jstsp_mc_admm", "jstsp_omp", "jstsp_omp_kron", "jstsp_somp", "jstsp_sparse_admm", "jstsp_vamp",
    "jstsp_wideband_mmwave_channel", "jstsp_measure", "jstsp_create_beamformer", "jstsp_qam4mod", "jstsp_nmse", "jstsp_admm_parameters", "jstsp_log2det_rate", "jstsp_draw_trials", "jstsp_philox4x32_10", "jstsp_capacity_sweep",
]


class Handle:
    """Owns one ``jstsp_handle`` (one CUDA context / stream / workspace)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib.jstsp_create(C.byref(self._h), int(device))
        if rc != 0:
            raise JstspError(rc, "jstsp_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")
        self.device = device

    def check(self, rc):
        if rc < 0:
            raise JstspError(rc, lib.jstsp_last_error(self._h).decode())
        return rc

    @property
    def ptr(self):
        return self._h

    def set_stream(self, cuda_stream_ptr: int):
        self.check(lib.jstsp_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self.check(lib.jstsp_synchronize(self._h))

    def set_chunk(self, n: int):
        self.check(lib.jstsp_set_chunk(self._h, int(n)))

    def profile(self, enable: int):
        self.check(lib.jstsp_profile(self._h, int(enable)))

    def profile_read(self):
        """{kernel class: (total_ms, launches)} accumulated since profiling was (re)started."""
        out = {}
        slot = 0
        while True:
            ms, n, name = C.c_double(), C.c_longlong(), C.c_char_p()
            rc = lib.jstsp_profile_read(self._h, slot, C.byref(ms), C.byref(n), C.byref(name))
            if rc != 0:
                break
            out[name.value.decode()] = (ms.value, n.value)
            slot += 1
        return out

    def nonfinite_count(self) -> int:
        """Trials of the last solver call with a non-finite output (synchronises the stream; for DEVICE-buffer calls, which return 0)."""
        return int(lib.jstsp_nonfinite_count(self._h))

    @property
    def last_variant(self) -> int:
        """Form of the Psi-domain path of the last pass: 1 = one persistent kernel per pass (csrc/admm_mega.cuh), 0 = four kernels per iteration."""
        return int(lib.jstsp_last_variant(self._h))

    @property
    def last_path(self) -> int:
        """Path of the last proposed_algorithm_psi call: 1 = dense kernels on the materialised B, 2 = Psi-domain tcgen05 kernel."""
        return int(lib.jstsp_last_path(self._h))

    @property
    def launches(self) -> int:
        return int(lib.jstsp_launch_count(self._h))

    def close(self):
        if self._h:
            lib.jstsp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_handle(device: int = 0) -> Handle:
    if device not in _default:
        _default[device] = Handle(device)
    return _default[device]
