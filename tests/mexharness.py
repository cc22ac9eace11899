"""Drives the MEX gateways (jstsp19_b200/mex/*.c) without MATLAB: each gateway is compiled against the
in-repo ``mex.h`` shim into ``jstsp19_b200/mex/build/<name>.so`` (``make mex-shim``) and its ``mexFunction`` is
called through ctypes with fake ``mxArray`` objects.  ``mexCallMATLAB`` (randn / rand / randperm / svd) is served
by a Python callback, so a test can feed the gateway the same draws it feeds the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "jstsp19_b200", "mex", "build")
GATEWAYS = ["wideband_mmwave_channel", "wideband_hybBF_comm_system_training", "proposed_hbf", "hbf", "proposed_algorithm",
            "proposed_algorithm_angles", "proposed_algorithm_psi", "svt", "mc_svt", "mc_admm", "sparse_admm", "OMP", "vamp", "OMP_kron", "jstsp_somp", "proposed_algorithm_pilots", "createBeamformer", "qam4mod", "ls_estimate", "capacity_sweep"]

_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_char_p)


def build():
    subprocess.run(["make", "-C", ROOT, "mex-shim"], check=True, stdout=subprocess.DEVNULL)


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.ident = ident


class Gateway:
    def __init__(self, name):
        path = os.path.join(BUILD, name + ".so")
        if not os.path.exists(path):
            build()
        self.name = name
        self.so = so = C.CDLL(path)
        vp, sz = C.c_void_p, C.c_size_t
        so.mxCreateDoubleMatrix.restype = vp; so.mxCreateDoubleMatrix.argtypes = [sz, sz, C.c_int]
        so.mxCreateNumericArray.restype = vp; so.mxCreateNumericArray.argtypes = [sz, C.POINTER(sz), C.c_int, C.c_int]
        so.mxCreateString.restype = vp; so.mxCreateString.argtypes = [C.c_char_p]
        so.mxGetCell.restype = vp; so.mxGetCell.argtypes = [vp, sz]
        so.mxDestroyArray.argtypes = [vp]
        so.mxGetNumberOfDimensions.restype = sz; so.mxGetNumberOfDimensions.argtypes = [vp]
        so.mxGetDimensions.restype = C.POINTER(sz); so.mxGetDimensions.argtypes = [vp]
        so.mxIsComplex.argtypes = [vp]; so.mxIsCell.argtypes = [vp]
        so.mxGetDoubles.restype = vp; so.mxGetDoubles.argtypes = [vp]
        so.mxGetComplexDoubles.restype = vp; so.mxGetComplexDoubles.argtypes = [vp]
        so.jstsp_shim_last_error_id.restype = C.c_char_p; so.jstsp_shim_last_error_msg.restype = C.c_char_p
        so.jstsp_shim_last_warning_id.restype = C.c_char_p
        so.jstsp_shim_call.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
        so.jstsp_shim_set_callback.argtypes = [_CB]
        self._cb = None

    # ---- numpy <-> fake mxArray ----
    def to_mx(self, v):
        so = self.so
        if isinstance(v, str):
            return so.mxCreateString(v.encode())
        a = np.asarray(v)
        cplx = np.iscomplexobj(a)
        a = np.asarray(a, dtype=np.complex128 if cplx else np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        dims = (C.c_size_t * a.ndim)(*a.shape)
        mx = so.mxCreateNumericArray(a.ndim, dims, 6, 1 if cplx else 0)
        ptr = so.mxGetComplexDoubles(mx) if cplx else so.mxGetDoubles(mx)
        flat = np.asfortranarray(a).reshape(-1, order="F")
        C.memmove(ptr, flat.ctypes.data, flat.nbytes)
        return mx

    def from_mx(self, mx):
        so = self.so
        nd = so.mxGetNumberOfDimensions(mx)
        dims = [so.mxGetDimensions(mx)[i] for i in range(nd)]
        n = int(np.prod(dims))
        if so.mxIsCell(mx):
            return [self.from_mx(so.mxGetCell(mx, i)) for i in range(n)]
        cplx = bool(so.mxIsComplex(mx))
        ptr = so.mxGetComplexDoubles(mx) if cplx else so.mxGetDoubles(mx)
        out = np.empty(n, dtype=np.complex128 if cplx else np.float64)
        if n:
            C.memmove(out.ctypes.data, ptr, out.nbytes)
        return out.reshape(dims, order="F")

    def set_matlab(self, fn):
        """fn(name, [numpy args]) -> list of numpy outputs; serves mexCallMATLAB."""
        def cb(nlhs, plhs, nrhs, prhs, name):
            try:
                args = [self.from_mx(prhs[i]) for i in range(nrhs)]
                outs = fn(name.decode(), args)
                for i in range(nlhs):
                    plhs[i] = self.to_mx(outs[i])
                return 0
            except Exception:      # reported to the gateway as a failed call
                return 1
        self._cb = _CB(cb)
        self.so.jstsp_shim_set_callback(self._cb)

    def __call__(self, nlhs, *args):
        so = self.so
        ins = [self.to_mx(a) for a in args]
        prhs = (C.c_void_p * max(len(ins), 1))(*ins)
        plhs = (C.c_void_p * max(nlhs, 1))()
        fn = C.cast(so.mexFunction, C.c_void_p)
        failed = so.jstsp_shim_call(fn, nlhs, plhs, len(ins), prhs)
        for m in ins:
            so.mxDestroyArray(m)
        if failed:
            raise MexError(so.jstsp_shim_last_error_id().decode(), so.jstsp_shim_last_error_msg().decode())
        outs = [self.from_mx(plhs[i]) for i in range(max(nlhs, 1)) if plhs[i]]
        for i in range(max(nlhs, 1)):
            if plhs[i]:
                so.mxDestroyArray(plhs[i])
        self.warning = so.jstsp_shim_last_warning_id().decode()
        return outs

    def close(self):
        self.so.jstsp_shim_run_atexit()
