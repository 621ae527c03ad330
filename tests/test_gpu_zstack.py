"""zstack / zsplit / zshatter of streams on the device (ckl_zstack, ckl_zslice) against the reference's own
recipe (crackle/operations.py:424-662, run from the staged package when present) and its own tests
(automated_test.py:448-562): stacking independently compressed slabs == compressing the whole volume."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ref_ops():
    """the reference's Python zstack / zsplit (pure Python over its compiled module), when the staged package is there"""
    pkg = os.path.join(ROOT, "oracle", "_ref", "refpkg")
    if not os.path.exists(pkg):
        return None
    os.environ["CKL_SHIM_BACKEND"] = "ref"
    for p in (os.path.join(ROOT, "oracle", "_ref"), pkg, os.path.join(ROOT, "tests", "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import crackle
    return crackle


def _volumes():
    from crackle_b200 import synth
    rng = np.random.default_rng(11)
    return [synth.jittered_voronoi((97, 130, 9), 13, np.uint32, seed=4, id_bits=20),
            synth.jittered_voronoi((128, 96, 12), 16, np.uint64, seed=1, id_bits=40),
            np.asfortranarray(rng.integers(0, 255, (64, 64, 10)).astype(np.uint8)),
            np.ones((32, 32, 7), dtype=np.uint32, order="F")]


def test_zstack_of_slabs_is_the_monolithic_stream():
    import crackle_b200 as cb
    from oracle import oracle as O
    for v in _volumes():
        sz = v.shape[2]
        whole = O.compress(v, 0)
        for cuts in ([0, sz // 2, sz], [0, 1, 2, sz], list(range(sz + 1))):
            parts = [O.compress(np.asfortranarray(v[:, :, a:b]), 0) for a, b in zip(cuts[:-1], cuts[1:])]
            if O.header(whole)["crack_format"] != O.header(parts[0])["crack_format"] or len({O.header(p)["crack_format"] for p in parts}) > 1:
                with pytest.raises((RuntimeError, ValueError)):
                    cb.zstack(parts)
                continue
            got = cb.zstack(parts)
            assert got == whole, (v.shape, cuts)
    # arrays in, markov streams in (re-coded to order 0 first, like operations.zstack)
    v = _volumes()[0]
    assert cb.zstack([np.asfortranarray(v[:, :, :4]), O.compress(np.asfortranarray(v[:, :, 4:]), 3)]) == O.compress(v, 0)
    with pytest.raises(ValueError, match="same width and height"):
        cb.zstack([O.compress(v, 0), O.compress(np.asfortranarray(v[:50]), 0)])


def test_zsplit_and_zshatter_against_the_reference_recipe():
    import crackle_b200 as cb
    from oracle import oracle as O
    ref = _ref_ops()
    for v in _volumes():
        b = O.compress(v, 0)
        sz = v.shape[2]
        for z in (0, 3, sz - 1):
            before, middle, after = cb.zsplit(b, z)
            assert np.array_equal(cb.decompress(middle), v[:, :, z:z + 1])
            if z > 0:
                assert np.array_equal(cb.decompress(before), v[:, :, :z])
                assert before == O.compress(np.asfortranarray(v[:, :, :z]), 0) or O.header(before)["crack_format"] != O.header(O.compress(np.asfortranarray(v[:, :, :z]), 0))["crack_format"]
            else:
                assert before == b""
            if z + 1 < sz:
                assert np.array_equal(cb.decompress(after), v[:, :, z + 1:])
            else:
                assert after == b""
            if ref is not None and 0 < z < sz - 1:
                rb, rm, ra = ref.zsplit(b, z)
                assert (before, middle, after) == (bytes(rb), bytes(rm), bytes(ra))
        pieces = cb.zshatter(b)
        assert len(pieces) == sz
        if ref is not None:
            assert pieces == [bytes(p) for p in ref.zshatter(b)]
        assert cb.zstack(pieces) == b
        with pytest.raises(ValueError, match="outside the range"):
            cb.zsplit(b, sz)
