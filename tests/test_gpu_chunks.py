"""z-chunk pipelining (ckl_ctx_set_chunks): the single-GPU compress / decompress of a volume split into K z-ranges on
child contexts must produce exactly the bytes / voxels of the unchunked path -- the same property the reference's
own zstack test states for independently compressed slabs (automated_test.py:449-487 test_zstack_ones)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture()
def cctx():
    import crackle_b200 as cb
    c = cb.Context(0)
    yield c
    c.close()


def _roundtrip(c, v, order, want):
    b = c.compress(v, order)
    assert b == want
    out = c.decompress(b)
    assert np.array_equal(out.reshape(v.shape, order="F" if v.flags.f_contiguous else "C"), v)
    return b


@pytest.mark.parametrize("K", [2, 3, 7])
@pytest.mark.parametrize("order", [0, 5])
def test_chunked_equals_oracle(cctx, K, order):
    from oracle import oracle as O
    from crackle_b200 import synth
    rng = np.random.default_rng(K)
    vols = [synth.jittered_voronoi((96, 80, 11), 12, np.uint64, seed=K, id_bits=40),
            synth.random_blobs((61, 47, 8), 9, np.uint16, seed=3),
            np.asfortranarray(rng.integers(0, 3, (70, 50, 9)).astype(np.uint8)),          # PERMISSIBLE crack format
            np.zeros((40, 40, 7), np.uint32, order="F"),                                  # no chains at all (order forced to 0)
            np.ascontiguousarray(synth.random_blobs((33, 45, 7), 8, np.uint32, seed=1))]  # C-order header flag
    cctx.set_chunks(K)
    for v in vols:
        _roundtrip(cctx, v, order, O.compress(v, order))


@pytest.mark.parametrize("order", [0, 3])
def test_chunk_guess_of_crack_format_is_corrected(cctx, order):
    # chunk 0 is noise (its own pixel-pair count says PERMISSIBLE), the rest is uniform: the global decision is
    # IMPERMISSIBLE and chunk 0 must be re-encoded; and the other way round
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    a = np.zeros((64, 64, 8), np.uint16, order="F")
    a[:, :, :2] = rng.integers(0, 1000, (64, 64, 2))
    b = np.asfortranarray(rng.integers(0, 1000, (64, 64, 8)).astype(np.uint16))
    b[:, :, 6:] = 7
    cctx.set_chunks(4)
    for v in (a, b):
        _roundtrip(cctx, v, order, O.compress(v, order))


@pytest.mark.parametrize("name", golden_names())
def test_chunked_golden(cctx, name):
    g = load_golden(name)
    a = g["input"]
    v = np.asfortranarray(a) if bool(g["f_order"]) else np.ascontiguousarray(a)
    if v.ndim < 3 or v.shape[2] < 2:
        pytest.skip("single slice")
    cctx.set_chunks(2)
    for order in (0, 5):
        assert cctx.compress(v, order) == bytes(g[f"ckl_order{order}"])
        d = cctx.decompress(bytes(g[f"ckl_order{order}"]))
        assert np.array_equal(d.reshape(a.shape, order="F" if bool(g["f_order"]) else "C"), a)
    m = cctx.decompress(bytes(g["ckl_order0"]), label=int(g["label"]))
    assert np.array_equal(m.reshape(a.shape, order="F" if bool(g["f_order"]) else "C").view(bool), g["mask"].view(bool))


def test_chunked_z_ranges_masks_and_errors(cctx):
    from crackle_b200 import synth
    v = synth.jittered_voronoi((64, 48, 12), 10, np.uint32, seed=2, id_bits=16)
    cctx.set_chunks(3)
    b = cctx.compress(v, 5)
    for z0, z1 in ((0, 1), (3, 9), (11, 12), (0, 12), (5, 100)):
        got = cctx.decompress(b, z0, z1).reshape((64, 48, -1), order="F")
        assert np.array_equal(got, v[:, :, z0:z1])
    lab = int(v[10, 10, 5])
    m = cctx.decompress(b, label=lab).reshape(v.shape, order="F")
    assert np.array_equal(m.view(bool), v == lab)
    bad = bytearray(b); bad[-1] ^= 1
    with pytest.raises(RuntimeError, match="crack code crc mismatch on z=11"):
        cctx.decompress(bytes(bad))
    bad = bytearray(b); bad[-4 * 12 + 4 * 2] ^= 1; bad[-1] ^= 1          # two bad slices: the lowest z is reported
    with pytest.raises(RuntimeError, match="crack code crc mismatch on z=2"):
        cctx.decompress(bytes(bad))


def test_chunked_device_resident_and_automatic(cctx):
    # 256 x 256 x 1024 uint64 (512 MiB).  The automatic policy chunks HOST-resident volumes of this size (4 z-chunks, staggered
    # uploads) and leaves device-resident ones alone; every setting must give the bytes of chunks = 1.
    import torch
    from crackle_b200 import synth
    t = synth.jittered_voronoi_torch((256, 256, 1024), 24, np.uint64, seed=1, id_bits=40)
    cctx.set_chunks(1)
    n1 = cctx.compress_ptr(t.data_ptr(), 1, 8, 256, 256, 1024, True, 0)
    one = cctx.result_bytes()
    cctx.set_chunks(0)
    n0 = cctx.compress_ptr(t.data_ptr(), 1, 8, 256, 256, 1024, True, 0)
    assert n0 == n1 and cctx.result_bytes() == one
    host = t.cpu().numpy()                                   # (sz, sy, sx) C-contiguous == Fortran (sx, sy, sz)
    nh = cctx.compress_ptr(host.ctypes.data, 0, 8, 256, 256, 1024, True, 0)          # automatic: chunked, staggered uploads
    assert nh == n1 and cctx.result_bytes() == one
    out_h = np.empty_like(host)
    buf = np.frombuffer(one, dtype=np.uint8)
    cctx.decompress_into(buf.ctypes.data, 0, buf.size, 0, -1, None, out_h.ctypes.data, 0, out_h.nbytes)   # automatic: chunked
    assert np.array_equal(out_h, host)
    p, n = cctx.result_device()
    out = torch.empty_like(t)
    cctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), t.view(torch.int64))
    cctx.set_chunks(5)
    cctx.compress_ptr(t.data_ptr(), 1, 8, 256, 256, 1024, True, 5)
    five = cctx.result_bytes()
    out.zero_()
    p, n = cctx.result_device()
    cctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)   # 5 staggered chunks, markov order 5
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), t.view(torch.int64))
    cctx.set_chunks(1)
    cctx.compress_ptr(t.data_ptr(), 1, 8, 256, 256, 1024, True, 5)
    assert cctx.result_bytes() == five


@pytest.mark.parametrize("K", [1, 3])
def test_pageable_host_volumes_take_the_staged_copies(cctx, K):
    """numpy arrays are pageable memory: copies of >= 16 MB run through the pinned staging ring (copy_host, ckl_api.cu) on
    eight host threads.  Sizes that are no multiple of the stage or the page size, whole volume and a z-range, a label
    mask, chunked and unchunked: the bytes and voxels must be those of the plain copies (small volumes, the other tests)."""
    from oracle import oracle as O
    from crackle_b200 import synth
    v = synth.jittered_voronoi((331, 277, 53), 18, np.uint32, seed=5, id_bits=30)      # 19.4 MB: staged, ragged tail
    assert v.nbytes >= (16 << 20) and v.nbytes % 4096 != 0
    ref = O.ref_module()
    want = bytes(ref.compress(v, False, True, 0, False, True, 0, 0)) if ref is not None else O.compress(v, 0)
    cctx.set_chunks(K)
    b = cctx.compress(v, 0)
    assert b == want
    out = cctx.decompress(b)
    assert np.array_equal(out.reshape(v.shape, order="F"), v)
    part = cctx.decompress(b, 3, 50)                                                   # 17.2 MB: staged as well
    assert np.array_equal(part.reshape((331, 277, 47), order="F"), v[:, :, 3:50])
    w = synth.jittered_voronoi((256, 256, 40), 16, np.uint64, seed=6, id_bits=40)      # 21 MB, 8-byte labels
    lab = int(w[100, 100, 20])
    bw = cctx.compress(w, 0)
    assert bw == (bytes(ref.compress(w, False, True, 0, False, True, 0, 0)) if ref is not None else O.compress(w, 0))
    assert np.array_equal(cctx.decompress(bw).reshape(w.shape, order="F"), w)
    assert np.array_equal(cctx.decompress(bw, label=lab).reshape(w.shape, order="F").view(bool), w == lab)
