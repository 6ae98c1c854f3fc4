"""CPU-side checks: the C-ABI library builds, loads, exports every symbol the header declares,
and the product package never reaches into oracle/ (no CPU fallback on the product path)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from benerf_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_are_exported_and_typed(lib):
    from benerf_b200 import _lib
    header = open(os.path.join(ROOT, "include", "benerf_b200.h")).read()
    declared = set(re.findall(r"\b(bnrf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/benerf_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.bnrf_abi_version() == 1


def test_create_fails_loudly_without_gpu(lib):
    import ctypes as C
    import torch
    from benerf_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    cfg = _lib.Cfg(64, 64, 3, 1, 0.0, 1.0, 0, 0)
    rc = lib.bnrf_create(C.byref(ctx), 0, C.byref(cfg))
    assert rc == _lib.ERR_DEVICE and not ctx.value
    assert b"no CPU path" in lib.bnrf_last_error(None)
    with pytest.raises(Exception):
        from benerf_b200.engine import Engine
        Engine()


def test_bad_config_is_rejected(lib):
    import ctypes as C
    from benerf_b200 import _lib
    ctx = C.c_void_p()
    for cfg in (_lib.Cfg(2, 0, 3, 1, 0, 1, 0, 0), _lib.Cfg(64, 64, 2, 1, 0, 1, 0, 0), _lib.Cfg(400, 200, 3, 1, 0, 1, 0, 0)):
        assert lib.bnrf_create(C.byref(ctx), 0, C.byref(cfg)) == _lib.ERR_ARG


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "benerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
                assert "/root/reference" not in src, f"{f} reads the reference tree"


def test_sass_contains_blackwell_tensor_and_tma_instructions(lib):
    import subprocess
    from benerf_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/benerf_b200.h must compile as C (no C++ types, no torch)."""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "benerf_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_bulk_store_and_tile_kernels_are_in_the_library(lib):
    """The backward pass's kernels are part of the shipped .so (no torch / cuBLAS fallback for dgrad / wgrad)."""
    import subprocess
    from benerf_b200 import _lib
    names = subprocess.run(["cuobjdump", "-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for k in ("dgrad_chain_kernel", "tile_wgrad_kernel", "tile_dgrad_kernel", "mlp_tc2_kernel", "adam_step_kernel"):
        assert k in names, k
