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
    assert lib.bnrf_abi_version() == 2


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


@pytest.mark.parametrize("split", [0, 1, 2, 4])
def test_forward_kernel_issue_schedule_invariants(split):
    """The forward kernel (csrc/mlp_tc3.cu) issues its MMAs from a host-built table.  Whatever the split of a layer's tail into
    N-halves, per tile: every GEMM step multiplies every K-block against every output column exactly once, each accumulator half is
    overwritten by its first group and committed by its last, the encoded points are waited for once and released once, a group
    waits for the previous layer's chunks before the first read of its K-block, and the weight stream has the bytes the groups
    consume.  Pure host code: runs without a GPU."""
    import ctypes as C
    from benerf_b200 import _lib
    lib = _lib.load()
    buf = (C.c_int32 * (4 * 64))()
    nbytes = C.c_int64(0)
    n = lib.bnrf_debug_tc3_schedule(split, buf, 64, C.byref(nbytes))
    assert 0 < n <= 64
    groups = [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]
    FIRST, ACC0, ACC1, PE_EMPTY, PE_FULL, WAIT_A = 1, 2, 4, 8, 16, 32
    assert [g[0] for g in groups] == sorted(g[0] for g in groups)                    # steps in order
    total = 0
    for t in range(9):
        gs = [g for g in groups if g[0] == t]
        blocks = ([-1] if t in (0, 5) else []) + ([] if t == 0 else [0, 1, 2, 3])
        for half in (0, 1):
            touching = [g for g in gs if g[1] == 0 or g[1] == 1 + half]
            assert sorted(g[2] for g in touching) == sorted(blocks), (t, half)           # every K-block once per column half
            assert touching[0][3] & FIRST, (t, half)                                     # the first group overwrites the accumulator
            assert not any(g[3] & FIRST for g in touching[1:] if g[1] != 0), (t, half)   # ... and no later one of this half does
            commits = [g for g in touching if g[3] & (ACC0 if half == 0 else ACC1)]
            assert len(commits) == 1 and commits[0] is touching[-1], (t, half)           # committed once, by the last group
        seen = set()
        for g in gs:                                                                      # A_READY is awaited at the first read of a K-block
            if g[2] >= 0:
                assert bool(g[3] & WAIT_A) == (g[2] not in seen), (t, g)
                seen.add(g[2])
            total += 2 * (16384 if (g[1] == 0 and t < 8) else 8192)
    assert sum(1 for g in groups if g[3] & PE_FULL) == 1 and groups[0][3] & PE_FULL and groups[0][2] == -1
    assert sum(1 for g in groups if g[3] & PE_EMPTY) == 1 and [g for g in groups if g[3] & PE_EMPTY][0][:3] == (5, 0, -1)
    assert nbytes.value == total
