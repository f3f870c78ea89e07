"""The C-ABI library must load without a GPU, export every symbol include/rbis_batch.h declares, and
refuse to compute (there is no CPU fallback).  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "rbis_batch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(rbis_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_are_exported_and_bound(rbis_lib):
    from pronto_b200 import capi

    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(rbis_lib, name), f"librbis_b200.so does not export {name}"
        assert name in capi.PROTOTYPES, f"pronto_b200.capi has no prototype for {name}"
    assert sorted(capi.PROTOTYPES) == declared


def test_struct_layouts_match_header(rbis_lib):
    from pronto_b200 import capi

    assert C.sizeof(capi.Op) == 32
    assert C.sizeof(capi.Config) == 56
    assert C.sizeof(capi.Stream) == 4 * 4 + 9 * 4 + 4 + 3 * 8 + 8
    cfg = capi.Config()
    rbis_lib.rbis_default_config(C.byref(cfg))
    assert cfg.g_val == 9.8 and cfg.chi_tol == 1e-6 and cfg.ctor_folds_chi == 1 and cfg.renormalize_quat == 0


def test_argument_errors_do_not_need_a_gpu(rbis_lib):
    h = C.c_void_p()
    assert rbis_lib.rbis_batch_create(None, 4, None) == -1
    assert rbis_lib.rbis_batch_create(C.byref(h), 0, None) == -1
    assert b"positive" in rbis_lib.rbis_last_error()
    assert rbis_lib.rbis_batch_synchronize(None) == -1
    assert rbis_lib.rbis_batch_num_filters(None) == 0
    assert rbis_lib.rbis_batch_destroy(None) == 0


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pronto_b200 import RBISBatch, capi

    with pytest.raises(capi.RBISError, match="no CUDA device|CUDA"):
        RBISBatch(16)


def test_reduce_chunks_is_fixed_order(rbis_lib):
    from pronto_b200 import reduce_chunks

    rng = np.random.default_rng(1)
    chunks = rng.normal(size=(37, 96)) * 10.0 ** rng.integers(-8, 8, size=(37, 96))
    out = reduce_chunks(chunks)
    exp = np.zeros(96)
    for c in range(37):
        exp = exp + chunks[c]  # ascending chunk order, one rounding per add
    assert np.array_equal(out, exp)
