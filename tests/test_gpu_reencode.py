"""re-coding a stream's crack codes with another markov order on the GPU (ckl_reencode) against
crackle::reencode_with_markov_order (src/crackle.hpp:860-984) of the compiled reference, byte for byte, and against a
fresh compress at the target order (the reference's own test: automated_test.py:834-847 test_reencode)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _volumes():
    from crackle_b200 import synth
    rng = np.random.default_rng(5)
    empty_mid = synth.jittered_voronoi((64, 48, 6), 12, np.uint32, seed=7, id_bits=20)
    empty_mid[:, :, 2:4] = 9                                           # uniform slices: empty crack codes
    return [synth.jittered_voronoi((97, 130, 9), 13, np.uint32, seed=4, id_bits=20),
            synth.jittered_voronoi((256, 192, 6), 20, np.uint64, seed=1, id_bits=40),
            synth.random_blobs((61, 47, 6), 9, np.uint16, seed=3),
            np.asfortranarray(rng.integers(0, 50, (40, 37, 5)).astype(np.uint8)),           # PERMISSIBLE crack format
            np.asfortranarray(empty_mid),
            np.asfortranarray(np.full((16, 12, 3), 7, dtype=np.uint32))]


@pytest.mark.parametrize("src_order,dst_order", [(0, 1), (0, 5), (5, 0), (3, 5), (2, 2), (1, 0)])
def test_reencode_matches_reference_bytes(src_order, dst_order):
    import crackle_b200 as cb
    from oracle import oracle as O
    ref = O.ref_module()
    for v in _volumes():
        src = O.compress(v, src_order)
        got = cb.reencode(src, dst_order)
        # a fresh encode at the target order is the same stream (same codepoints, same global statistics), except that the
        # reference keeps the requested order in the header even when no slice has a codepoint (crackle.hpp:107-118 applies
        # to compress only)
        fresh = O.compress(v, dst_order)
        if O.header(fresh)["order"] == dst_order:
            assert got == fresh, (v.shape, v.dtype)
        if ref is not None:
            assert got == bytes(ref.reencode_markov(src, dst_order, 1)), (v.shape, v.dtype)
        assert np.array_equal(cb.decompress(got).reshape(v.shape), v)


def test_reencode_v0_stream_and_device_pointers():
    import torch
    import crackle_b200 as cb
    from oracle import oracle as O
    from test_format_v0 import to_v0
    from crackle_b200 import synth
    v = synth.jittered_voronoi((96, 80, 5), 12, np.uint64, seed=2, id_bits=40)
    v0 = to_v0(O.compress(v, 0))
    got = cb.reencode(v0, 4)
    assert cb.header(got)["format_version"] == 0 and cb.header(got)["markov_model_order"] == 4
    assert got == to_v0(O.compress(v, 4))
    # (the reference itself writes a 29-byte version-1 header layout under version byte 0 here -- header.tobytes() has one
    # size, header.hpp:269-273 -- which its own decoder then misreads; the v0 result is therefore pinned by the v0 form of
    # a fresh encode, not by the reference's output)
    # 1024x1024 slab of the bench volume, stream resident on the device
    ctx = cb.Context(0)
    t = synth.jittered_voronoi_torch((1024, 1024, 8), 24, np.uint64, seed=0, id_bits=40, sz_total=1024)
    b0, b5 = ctx.compress(t, 0), ctx.compress(t, 5)
    d = torch.from_numpy(np.frombuffer(b0, dtype=np.uint8).copy()).cuda()
    import ctypes
    from crackle_b200 import _capi
    n = ctypes.c_uint64()
    ctx._check(_capi.lib().ckl_reencode(ctx._h, d.data_ptr(), 1, d.numel(), 5, ctypes.byref(n)))
    assert ctx.result_bytes() == b5
    assert ctx.reencode(b5, 0) == b0
