"""The pybind11 `fastcrackle` drop-in (crackle_b200/csrc/fastcrackle_module.cpp): same positional signatures as the
reference module (src/fastcrackle.cpp:84-129, :163-210, registered without keyword names :643-644)."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _module():
    from crackle_b200 import build
    build.build_lib()
    build.build_pymodule()
    sys.path.insert(0, os.path.join(ROOT, "crackle_b200"))
    try:
        return importlib.import_module("fastcrackle")
    finally:
        sys.path.pop(0)


def test_module_surface_and_no_cpu_fallback():
    import torch
    m = _module()
    assert callable(m.compress) and callable(m.decompress)
    g = load_golden("voronoi_u64_96x80x5")
    # header problems are reported with the reference's text even without a GPU
    with pytest.raises(RuntimeError, match="Data stream is not valid|too small"):
        m.decompress(b"\x00" * 40, 0, -1, 1, None)
    with pytest.raises(RuntimeError, match="1D"):
        m.decompress(np.zeros((4, 4), np.uint8), 0, -1, 1, None)
    with pytest.raises(RuntimeError, match="allow_pins"):
        m.compress(np.asfortranarray(g["input"]), True, True, 0, False, True, 0, 1)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            m.compress(np.asfortranarray(g["input"]), False, True, 0, False, True, 0, 1)


@pytest.mark.gpu
def test_module_matches_reference_golden():
    m = _module()
    g = load_golden("voronoi_u64_96x80x5")
    a = np.asfortranarray(g["input"])
    for order in (0, 5):
        b = m.compress(a, False, True, order, False, True, 0, 0)            # codec.py:729-733 calls it positionally
        assert isinstance(b, bytes) and b == bytes(g[f"ckl_order{order}"])
        d = m.decompress(b, 0, -1, 0, None)                                # codec.py:670
        assert d.ndim == 1 and d.dtype == a.dtype
        assert np.array_equal(d.reshape(a.shape, order="F"), a)
    lab = int(g["label"])
    mk = m.decompress(bytes(g["ckl_order0"]), 0, -1, 0, lab)
    assert mk.dtype == np.uint8 and np.array_equal(mk.reshape(a.shape, order="F").view(bool), g["mask"].view(bool))
    z = m.decompress(bytes(g["ckl_order0"]), 1, 3, 0, None)
    assert np.array_equal(z.reshape((96, 80, 2), order="F"), a[:, :, 1:3])
