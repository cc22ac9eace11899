"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "jstsp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jstsp_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from jstsp19_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (jstsp_[a-z_0-9]+)", out))
    decl = _declared()
    assert decl, "no declarations parsed"
    missing = [s for s in decl if s not in exported]
    assert not missing, missing
    assert sorted(_lib.EXPORTED) == decl          # the ctypes binding covers the whole header


def test_library_is_sm100a_cuda_code():
    from jstsp19_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jstsp19_b200 import _lib
    with pytest.raises(_lib.JstspError):
        _lib.Handle(0)
    import numpy as np
    import jstsp19_b200 as jb
    with pytest.raises(_lib.JstspError):
        jb.svt(np.ones((4, 6), complex), 0.1)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jstsp19_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
