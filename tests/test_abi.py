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
    # hidden activations in tensor memory: tcgen05.st -> STTM, and the MMA form that takes its A operand from TMEM
    assert "STTM" in sass, "tcgen05.st (STTM) missing from SASS"
    assert re.search(r"UTCHMMA(\.2CTA)? tmem\[", sass), "no tcgen05.mma with the A operand in tensor memory"
    # the default kernels of the fused path
    for k in ("k_front_ts", "k_back_ts", "k_ep_dispatch", "k_ep_plan"):
        assert k in sass, f"kernel {k} missing"


def test_argument_errors_are_reported_not_crashed(built_lib):
    """Entry points validate their arguments before touching the device: non-zero status + snb_last_error()."""
    from switch_nerf_b200 import _lib as L
    lib = L.lib()
    opts = L.RouteOpts(1.0, 1, 0)
    assert lib.snb_moe_layer_forward(None, None, None, 16, C.byref(opts), None, None, None, None, 0, None) != 0
    assert b"NULL" in lib.snb_last_error()
    assert lib.snb_moe_forward(None, None, 16, None, C.byref(opts), 1, None, None, None, None, None, None, 0, None) != 0
    assert lib.snb_model_attach_a2a(None, None) != 0
    h = C.c_void_p()
    assert lib.snb_a2a_init(5, 4, 8, 1024, 1.0, C.byref(h)) != 0 and b"rank" in lib.snb_last_error()
    assert lib.snb_a2a_init(0, 9, 9, 1024, 1.0, C.byref(h)) != 0
    assert lib.snb_a2a_connect(None, None, 0) != 0
    out = (C.c_uint64 * 6)()
    assert lib.snb_umma_microbench(96, 0, 0, 16, out, None) != 0 and b"N must be" in lib.snb_last_error()
    assert lib.snb_a2a_handle_bytes() == 64
