"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "switch_nerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from switch_nerf_b200 import _lib as L
    handle = C.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/switch_nerf_b200.h but not exported"
    assert set(L.EXPORTED_SYMBOLS) == set(declared), "ctypes binding table and header disagree"
    assert L.lib().snb_version() >= 100


def test_no_cpu_path(built_lib):
    """Product code must fail loudly instead of falling back to CPU/PyTorch."""
    from switch_nerf_b200 import _lib as L
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    desc, w, h = L.ModelDesc(4, 256, 7, 3, 2, 12, 4, 48, 16, 128, 0), L.Weights(), C.c_void_p()
    rc = L.lib().snb_model_create(C.byref(desc), C.byref(w), None, C.byref(h))
    assert rc != 0 and b"no CUDA device" in L.lib().snb_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "switch_nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), f"{f} references the oracle"


def test_sass_has_blackwell_instructions(built_lib):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, bulk copy -> UBLKCP (B200_PROFILING.md)."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnem in sass, f"{mnem} missing from SASS"
