"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/crackle_b200.h
declares, parses headers on the host, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden
from crackle_b200 import _capi, codec


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "crackle_b200.h")).read()
    return sorted(set(re.findall(r"CKL_API\s+[\w\s\*]+?\b((?:crackle_b200|ckl)_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for must in ("crackle_b200_compress", "crackle_b200_decompress", "crackle_b200_free", "ckl_compress", "ckl_decompress",
                 "ckl_shard_begin", "ckl_shard_encode", "ckl_shard_finish", "ckl_shard_fetch"):
        assert must in names


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/crackle_b200.h but not exported"
        assert name in _capi.SYMBOLS, f"{name} has no ctypes prototype in crackle_b200/_capi.py"


def test_version_string():
    assert b"sm_100a" in _capi.lib().crackle_b200_version()


@pytest.mark.parametrize("name", golden_names()[:6])
def test_host_header_parse(name):
    g = load_golden(name)
    a = g["input"]
    h = codec.header(bytes(g["ckl_order5"]))
    s = list(a.shape) + [1, 1]
    assert (h["sx"], h["sy"], h["sz"]) == (s[0], s[1], s[2])
    assert h["data_width"] == a.dtype.itemsize
    assert h["format_version"] == 1 and h["label_format"] == 0
    assert h["fortran_order"] == int(bool(g["f_order"]))


def test_header_errors_match_reference_text():
    g = load_golden("island_6x6")
    b = bytearray(bytes(g["ckl_order0"]))
    with pytest.raises(RuntimeError, match="Input too small"):
        codec.header(bytes(b[:10]))
    bad = bytearray(b); bad[0] = ord("x")
    with pytest.raises(RuntimeError, match="Data stream is not valid"):
        codec.header(bytes(bad))
    bad = bytearray(b); bad[8] ^= 0x40
    with pytest.raises(RuntimeError, match="CRC8 check failed"):
        codec.header(bytes(bad))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        codec.Context(0)
    a = np.zeros((4, 4, 1), dtype=np.uint8, order="F")
    out, n = ctypes.c_void_p(), ctypes.c_uint64()
    err = ctypes.create_string_buffer(256)
    rc = _capi.lib().crackle_b200_compress(a.ctypes.data, 1, 4, 4, 1, 1, 0, ctypes.byref(out), ctypes.byref(n), err, 256)
    assert rc == 1 and b"no CUDA device" in err.value
    with pytest.raises(RuntimeError):
        codec.compress(a)


def test_signed_rejected_like_reference():
    with pytest.raises(TypeError, match="Signed integer"):
        codec._as_fortran_volume(np.zeros((2, 2, 2), dtype=np.int32))


def test_counters_and_argument_checks_that_need_no_device():
    """the launch / drain counters are plain host counters; crackle.voxel_connectivity_graph rejects other connectivities in
    Python (operations.py:947-948) before anything touches the device"""
    import crackle_b200 as cb
    assert codec.launch_count() >= 0 and codec.sync_count() >= 0
    g = load_golden("island_6x6")
    with pytest.raises(ValueError, match="Only 4 and 6 connected are supported. Got: 8"):
        cb.voxel_connectivity_graph(bytes(g["ckl_order0"]), connectivity=8)
