"""The C-ABI library loads and exports every symbol include/svdq.h declares (no compute without a GPU);
argument validation returns the documented negative codes before any CUDA call."""
import ctypes as C
import os
import re

import pytest

from svd_quantization_task_merging_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "svdq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svdq_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/svdq.h but not exported by libsvdq.so"
    assert lib.svdq_abi_version() == _native.ABI_VERSION
    assert lib.svdq_k4_scratch_bytes() > 0


def test_invalid_arguments_are_value_errors_without_touching_the_gpu():
    lib = _native.load()
    # n_tasks out of range
    rc = lib.svdq_tv_mask_gram(0, 0, 0, 0, 1, 1024, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0 and b"n_tasks" in lib.svdq_last_error()
    # unknown mask strategy
    rc = lib.svdq_tv_mask_gram(0, 8, 7, 0, 1, 1024, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0 and b"Unknown mask strategy" in lib.svdq_last_error()
    # tile size must be a multiple of 1024
    rc = lib.svdq_tv_mask_gram(0, 8, 0, 0, 1, 1000, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0
    # empty mask list -> ValueError text of the reference (mask_loader.py:424)
    rc = lib.svdq_combine_masks(None, 0, 10, 0, None, None)
    assert rc < 0 and b"Empty mask list" in lib.svdq_last_error()
    with pytest.raises(ValueError):
        _native.check(rc)
    # config validation mirrors (config.py:225-234)
    rc = lib.svdq_param_solve(8, 1, 1, C.c_float(1.5), 0, 10, 4, 2, *([None] * 22))
    assert rc < 0 and b"Energy threshold" in lib.svdq_last_error()
    rc = lib.svdq_param_solve(8, 1, 1, C.c_float(0.9), 0, 10, 9, 2, *([None] * 22))
    assert rc < 0 and b"Low bits" in lib.svdq_last_error()
    # zero-sized work is a no-op, not an error
    assert lib.svdq_rtvq_quantize(None, 0, 4, 2, None, 0, 1, None, None, None, None, None, 0, None) == 0
    assert lib.svdq_unpack_mask(None, 0, None, None) == 0
    # K14 operator entries: shape checks before any launch
    assert lib.svdq_project_scratch_bytes() >= 32 * 8 and lib.svdq_select_chunk_elems() % 1024 == 0
    rc = lib.svdq_basis_project(0, None, 0, 0, 10, None, None, None, None, None)
    assert rc < 0 and b"cols" in lib.svdq_last_error()
    rc = lib.svdq_basis_project(0, None, 4, 33, 10, None, None, None, None, None)
    assert rc < 0 and b"cols" in lib.svdq_last_error()
    rc = lib.svdq_basis_expand(1, None, 0, -1, None, 0, 0, 10, None, None, None, C.c_float(1.0), None, None)
    assert rc < 0 and b"column counts" in lib.svdq_last_error()
    assert lib.svdq_basis_expand(1, None, 0, 3, None, 0, 0, 0, None, None, None, C.c_float(1.0), None, None) == 0   # no rows
    rc = lib.svdq_mask_offsets(None, -1, 0, None, None)
    assert rc < 0
    assert lib.svdq_mask_select(None, 4, None, 0, 0, None, None, None) == 0
    assert lib.svdq_mask_scatter(None, 4, None, 0, 1, None, None, None) == 0


def test_host_mask_packer_matches_numpy_packbits():
    """svdq_host_pack_mask is the host-side transfer encoding of one task mask (element 8k+i -> bit i of byte k)."""
    import numpy as np
    lib = _native.load()
    rng = np.random.default_rng(7)
    for n, threads in [(0, 1), (1, 1), (7, 1), (8, 2), (1023, 1), (70_001, 3), (1 << 20, 4), ((1 << 20) + 5, 16)]:
        src = (rng.random(n) < 0.3).astype(np.uint8)
        if n > 10:
            src[3] = 255                                    # any non-zero byte is a set bit
        dst = np.zeros((n + 7) // 8, np.uint8)
        rc = lib.svdq_host_pack_mask(src.ctypes.data if n else None, n, dst.ctypes.data if n else None, threads)
        assert rc == 0
        assert np.array_equal(dst, np.packbits(src != 0, bitorder="little")), (n, threads)
    assert lib.svdq_host_pack_mask(None, 8, None, 1) < 0
    assert lib.svdq_host_pack_mask(None, 0, None, 0) < 0


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig
    with pytest.raises(_native.NativeLibraryError):
        rtvq.asymmetric_quantization(torch.randn(10), 4)
    with pytest.raises(_native.NativeLibraryError):
        merge_state_dicts({"w": torch.zeros(4)}, {"a": {"w": torch.ones(4)}}, None, SVDHybridConfig(tasks=["a"]))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "svd_quantization_task_merging_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)
