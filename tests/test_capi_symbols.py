"""The C-ABI library: loads, exports every symbol include/b200ann.h declares, and validates arguments without a GPU.
No compute call is made here."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "b200ann.h").read_text()
    return sorted(set(re.findall(r"ANN_API\s+[\w\s\*]+?\b(ann_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    from the_algorithm_b200 import _capi

    assert declared_symbols() == sorted(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for sym in declared_symbols():
        assert hasattr(lib, sym), f"{sym} declared in b200ann.h but not exported"


def test_header_cites_reference_lines():
    text = (ROOT / "include" / "b200ann.h").read_text()
    for cite in ("BruteForceIndex.scala:29-37", "BruteForceIndex.scala:48-52", "BruteForceIndex.scala:66-91",
                 "ShardApi.scala:72-86", "Api.scala", "ann_common.thrift:16-19"):
        assert cite in text


def test_header_is_plain_c(tmp_path):
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-c", "-o", str(tmp_path / "h.o"),
                    str(ROOT / "tests" / "capi_header_check.c")], check=True)


def test_argument_validation_without_gpu(built_lib):
    from the_algorithm_b200 import _capi

    L = _capi.lib()
    assert L.ann_version() >= 100
    h = ctypes.c_void_p()
    assert L.ann_create(None, ctypes.byref(h)) == _capi.ANN_ERR_NULL_POINTER
    assert b"NULL" in L.ann_last_error()
    bad = _capi.AnnConfig(7, 16, 0, 0, 0)
    assert L.ann_create(ctypes.byref(bad), ctypes.byref(h)) == _capi.ANN_ERR_INVALID_ARGUMENT
    bad = _capi.AnnConfig(0, 0, 0, 0, 0)
    assert L.ann_create(ctypes.byref(bad), ctypes.byref(h)) == _capi.ANN_ERR_INVALID_ARGUMENT
    assert L.ann_size(None, None) == _capi.ANN_ERR_NULL_POINTER
    assert L.ann_append_batch(None, None, None, 1) == _capi.ANN_ERR_NULL_POINTER
    assert L.ann_query_batch(None, None, 1, 1, 1, None, None, None) == _capi.ANN_ERR_NULL_POINTER
    assert L.ann_merge_topk_device(0, None, None, None, 0, 1, 1, None, None, None, None) == _capi.ANN_ERR_INVALID_ARGUMENT
    assert L.ann_merge_topk_device(0, None, None, None, 2, 1, -1, None, None, None, None) == _capi.ANN_ERR_NEGATIVE_K
    assert L.ann_query_seed_device(None, None, 1, 1, 1, None, None) == _capi.ANN_ERR_NULL_POINTER
    assert L.ann_query_finish_device(None, None, 1, 1, 1, None, 0, None, None, None, None) == _capi.ANN_ERR_NULL_POINTER
    assert L.ann_distance_pairs(9, 0, 4, None, None, 1, None, 0) == _capi.ANN_ERR_INVALID_ARGUMENT
    assert L.ann_distance_pairs(0, 0, 4000, None, None, 1, None, 0) == _capi.ANN_ERR_INVALID_ARGUMENT
    assert L.ann_distance_pairs(0, 0, 4, None, None, 1, None, 0) == _capi.ANN_ERR_NULL_POINTER
    assert L.ann_distance_pairs(0, 0, 4, None, None, 0, None, 0) == _capi.ANN_OK          # nothing to do
    assert L.ann_normalize_rows(0, None, 1, None, 0) == _capi.ANN_ERR_INVALID_ARGUMENT
    assert L.ann_normalize_rows(4, None, 1, None, 0) == _capi.ANN_ERR_NULL_POINTER
    L.ann_destroy(None)  # ignored


def test_no_cpu_fallback_when_device_missing(built_lib):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from the_algorithm_b200 import _capi

    h = ctypes.c_void_p()
    cfg = _capi.AnnConfig(2, 16, 0, 0, 0)
    rc = _capi.lib().ann_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc in (_capi.ANN_ERR_NO_DEVICE, _capi.ANN_ERR_CUDA)
    assert not h.value
    import numpy as np

    from the_algorithm_b200.ann.brute_force import BruteForceIndex
    from the_algorithm_b200.ann.common import FuturePool, InnerProduct

    ix = BruteForceIndex(InnerProduct, FuturePool.immediate_pool())
    with pytest.raises(_capi.AnnError):
        ix.append_batch([1], np.ones((1, 16), np.float32))
    # the Metric trait's own entry points compute on the device too: no host arithmetic stands in for them
    from the_algorithm_b200.ann.common import MetricUtil

    with pytest.raises(_capi.AnnError) as e:
        InnerProduct.distance(np.ones(16, np.float32), np.ones(16, np.float32))
    assert e.value.code in (_capi.ANN_ERR_NO_DEVICE, _capi.ANN_ERR_CUDA)
    with pytest.raises(_capi.AnnError):
        MetricUtil.norm(np.ones(16, np.float32))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    for p in (ROOT / "the-algorithm_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".c", ".scala") and p.is_file():
            text = p.read_text(errors="replace")
            assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, p
