"""Parity tests proper: the CUDA path, called through the C-ABI, against the golden vectors generated from the
compiled reference, against the CPU oracle on seeded inputs, and -- at BASELINE.json sizes -- through
size-independent properties.  Integer/byte work: the bar is bit-exact."""
import ctypes

import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import crackle_b200 as cb
    return cb.default_context()


def _golden_input(g):
    a = g["input"]
    return np.asfortranarray(a) if bool(g["f_order"]) else np.ascontiguousarray(a)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("order", [0, 1, 5])
def test_compress_matches_reference_golden(ctx, name, order):
    g = load_golden(name)
    assert ctx.compress(_golden_input(g), order) == bytes(g[f"ckl_order{order}"])


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("order", [0, 5])
def test_decompress_matches_reference_golden(ctx, name, order):
    import crackle_b200 as cb
    g = load_golden(name)
    a = g["input"]
    b = bytes(g[f"ckl_order{order}"])
    d = cb.decompress(b)
    assert d.dtype == a.dtype
    assert np.array_equal(d.reshape(a.shape), a)
    m = cb.decompress(b, label=int(g["label"]))
    assert m.dtype == bool and np.array_equal(m, g["mask"].view(bool))
    if "z1_2" in g:
        assert np.array_equal(cb.decompress_range(b, 1, 2), g["z1_2"])


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64])
def test_random_volumes_against_oracle(ctx, dtype):
    # mirrors automated_test.py:15-39 (test_compress_decompress_random), bytes checked against the oracle
    from oracle import oracle as O
    from crackle_b200 import synth
    import crackle_b200 as cb
    rng = np.random.default_rng(np.dtype(dtype).itemsize)
    vols = [np.asfortranarray(rng.integers(0, 5, (4, 4, 1)).astype(dtype)),
            np.asfortranarray(rng.integers(0, 255, (256, 255, 1)).astype(dtype)),
            np.asfortranarray(rng.integers(0, 255, (40, 37, 9)).astype(dtype)),
            np.asfortranarray(rng.integers(0, 2, (65, 33, 5)).astype(dtype)),
            synth.random_blobs((61, 47, 6), 9, dtype, seed=3),
            synth.jittered_voronoi((97, 130, 4), 13, dtype, seed=4, id_bits=8 * np.dtype(dtype).itemsize - 1)]
    for v in vols:
        for order in (0, 2, 5):
            b = ctx.compress(v, order)
            assert b == O.compress(v, order), (v.shape, order)
            assert np.array_equal(cb.decompress(b).reshape(v.shape), v)


def test_dense_crack_graphs(ctx):
    # noise volumes whose crack graphs have far more nodes than a segmentation's: exercises the 16-bit shared-memory and
    # the global-memory replay stores (> 8191 / > 16382 nodes per slice) and the global-memory fallback of the band CCL
    from oracle import oracle as O
    import crackle_b200 as cb
    rng = np.random.default_rng(7)
    vols = [np.asfortranarray(rng.integers(0, 3, (110, 110, 2)).astype(np.uint8)),
            np.asfortranarray(rng.integers(0, 3, (180, 170, 2)).astype(np.uint16)),
            np.asfortranarray(rng.integers(0, 2, (512, 520, 1)).astype(np.uint8)),
            np.asfortranarray(rng.integers(0, 2000, (300, 260, 2)).astype(np.uint32))]
    for v in vols:
        for order in (0, 3):
            b = ctx.compress(v, order)
            assert b == O.compress(v, order), (v.shape, order)
            assert np.array_equal(cb.decompress(b).reshape(v.shape), v)


def test_edge_shapes(ctx):
    # automated_test.py:151-225, 263-271: empty, black, uniform, arange, 2D, 1D inputs
    from oracle import oracle as O
    import crackle_b200 as cb
    cases = [np.zeros((0, 0, 0), np.uint32, order="F"), np.zeros((3, 0, 2), np.uint8, order="F"),
             np.zeros((1, 1, 1), np.uint64, order="F"), np.zeros((100, 100, 3), np.uint32, order="F"),
             np.full((31, 33, 2), 2**40 + 7, np.uint64, order="F"),
             np.asfortranarray(np.arange(33 * 32 * 2, dtype=np.uint32).reshape((33, 32, 2), order="F")),
             np.arange(129, dtype=np.uint16), np.asfortranarray(np.arange(64, dtype=np.uint8).reshape(8, 8)),
             np.asfortranarray((np.indices((50, 50))[0] // 10).astype(np.uint8)),
             np.ascontiguousarray(np.random.default_rng(0).integers(0, 4, (20, 30, 5)).astype(np.uint16))]
    for v in cases:
        b = ctx.compress(v, 0)
        assert b == O.compress(v, 0), v.shape
        if v.size:
            d = cb.decompress(b)
            s = list(v.shape) + [1, 1]
            assert np.array_equal(d.reshape((s[0], s[1], s[2])), v.reshape((s[0], s[1], s[2])))
        else:
            assert len(b) == 29


def test_c_order_roundtrip_layout(ctx):
    # automated_test.py:658-675 (test_contiguous_fortran): C-order volumes keep their memory order
    import crackle_b200 as cb
    from crackle_b200 import synth
    v = np.ascontiguousarray(synth.random_blobs((33, 45, 7), 8, np.uint32, seed=1))
    b = ctx.compress(v, 0)
    assert cb.header(b)["fortran_order"] == 0
    d = cb.decompress(b)
    assert d.flags.c_contiguous and np.array_equal(d, v)
    lab = int(v[10, 10, 3])
    assert np.array_equal(cb.decompress(b, label=lab), v == lab)


def test_z_ranges(ctx):
    import crackle_b200 as cb
    from crackle_b200 import synth
    v = synth.jittered_voronoi((64, 48, 12), 10, np.uint32, seed=2, id_bits=16)
    b = ctx.compress(v, 5)
    for z0, z1 in ((0, 1), (3, 9), (11, 12), (0, 12), (5, 100)):
        assert np.array_equal(cb.decompress_range(b, z0, z1), v[:, :, z0:z1])
    with pytest.raises(RuntimeError, match="Invalid range"):
        cb.decompress_range(b, 5, 5)


def test_corruption_errors_match_reference_text(ctx):
    # automated_test.py:731-826 (test_crc_check_one_bit_error)
    import crackle_b200 as cb
    from crackle_b200 import synth
    v = synth.jittered_voronoi((64, 64, 4), 12, np.uint16, seed=5, id_bits=15)
    b = bytearray(ctx.compress(v, 0))
    bad = bytearray(b); bad[-1] ^= 1
    with pytest.raises(RuntimeError, match="crack code crc mismatch on z=3"):
        cb.decompress(bytes(bad))
    bad = bytearray(b); bad[29] ^= 1
    with pytest.raises(RuntimeError, match="grid index crc32c did not match"):
        cb.decompress(bytes(bad))
    bad = bytearray(b); bad[9] ^= 1
    with pytest.raises(RuntimeError, match="CRC8 check failed"):
        cb.decompress(bytes(bad))
    with pytest.raises(RuntimeError, match="Input too small"):
        cb.decompress(bytes(b[:12]))
    # a flipped crack-code bit must be caught by the per-slice crc
    h = cb.header(bytes(b))
    off = 29 + 4 * (h["sz"] + 1) + h["num_label_bytes"] + 12
    bad = bytearray(b); bad[off] ^= 0x10
    with pytest.raises(RuntimeError):
        cb.decompress(bytes(bad))


def test_one_shot_c_abi(ctx):
    # the exact entry points a fastcrackle binding would call
    from crackle_b200 import _capi
    g = load_golden("voronoi_u64_96x80x5")
    a = np.asfortranarray(g["input"])
    L = _capi.lib()
    out, n = ctypes.c_void_p(), ctypes.c_uint64()
    err = ctypes.create_string_buffer(512)
    rc = L.crackle_b200_compress(a.ctypes.data, 8, 96, 80, 5, 1, 5, ctypes.byref(out), ctypes.byref(n), err, 512)
    assert rc == 0, err.value
    b = ctypes.string_at(out.value, n.value)
    L.crackle_b200_free(out)
    assert b == bytes(g["ckl_order5"])
    dec = np.zeros(a.size, dtype=np.uint64)
    buf = np.frombuffer(b, dtype=np.uint8)
    rc = L.crackle_b200_decompress(buf.ctypes.data, buf.size, 0, -1, 0, 0, dec.ctypes.data, dec.nbytes, err, 512)
    assert rc == 0, err.value
    assert np.array_equal(dec.reshape(a.shape, order="F"), a)
    rc = L.crackle_b200_decompress(buf.ctypes.data, buf.size, 0, -1, 0, 0, dec.ctypes.data, 8, err, 512)
    assert rc != 0 and b"too small" in err.value


def test_device_resident_tensors(ctx):
    import torch
    from oracle import oracle as O
    from crackle_b200 import synth
    v = synth.jittered_voronoi((128, 96, 8), 16, np.uint64, seed=6)
    t = synth.jittered_voronoi_torch((128, 96, 8), 16, np.uint64, seed=6)
    assert np.array_equal(t.cpu().numpy().transpose(2, 1, 0), v)          # numpy and torch generators agree
    want = O.compress(v, 0)
    assert ctx.compress(t, 0) == want
    n = ctx.compress_ptr(t.data_ptr(), 1, 8, 128, 96, 8, True, 0)
    p, n2 = ctx.result_device()
    assert n == n2 == len(want)
    out = torch.empty_like(t)
    ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), t.view(torch.int64))
    m = torch.empty(t.shape, dtype=torch.uint8, device="cuda")
    lab = int(v[64, 48, 4])
    ctx.decompress_into(p, 1, n, 0, -1, lab, m.data_ptr(), 1, m.numel())
    torch.cuda.synchronize()
    assert np.array_equal(m.cpu().numpy().transpose(2, 1, 0).astype(bool), v == lab)


@pytest.mark.parametrize("order", [0, 5])
def test_config1_512x512x64_u32_bytes_vs_oracle(ctx, order):
    # BASELINE.json configs[0]: 512x512x64 uint32 ~1k labels, byte-exact against the oracle
    from oracle import oracle as O
    from crackle_b200 import synth
    t = synth.jittered_voronoi_torch((512, 512, 64), 40, np.uint32, seed=0, id_bits=16)
    v = np.asfortranarray(t.cpu().numpy().transpose(2, 1, 0))
    b = ctx.compress(t, order)
    assert b == O.compress(v, order)
    import crackle_b200 as cb
    assert np.array_equal(cb.decompress(b), v)


@pytest.mark.parametrize("order", [0, 5])
def test_config2_512cube_u64_properties(ctx, order):
    # BASELINE.json configs[1]/[3]: 512^3 uint64, ~10k labels.  Full-size checks use size-independent properties:
    # round trip, per-slice crc agreement with the decoder, z-slab bytes == oracle, mask == (volume == label).
    import torch
    from oracle import oracle as O
    from crackle_b200 import synth
    import crackle_b200 as cb
    t = synth.jittered_voronoi_torch((512, 512, 512), 24, np.uint64, seed=0, id_bits=40)
    n = ctx.compress_ptr(t.data_ptr(), 1, 8, 512, 512, 512, True, order)
    p, _ = ctx.result_device()
    out = torch.empty_like(t)
    ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)     # verifies every slice crc
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), t.view(torch.int64))
    stream = ctx.result_bytes()
    h = cb.header(stream)
    assert (h["sx"], h["sy"], h["sz"], h["data_width"], h["stored_data_width"], h["markov_model_order"]) == (512, 512, 512, 8, 8, order)
    # oracle decodes a z-range of the GPU stream identically
    ref = O.decompress(stream, 100, 104)
    assert np.array_equal(ref, t[100:104].cpu().numpy().transpose(2, 1, 0))
    if order == 0:
        # order-0 crack codes are slice-local: a 12-slice slab compressed alone has the same per-slice codes and crcs
        slab = np.asfortranarray(t[200:212].cpu().numpy().transpose(2, 1, 0))
        so, sg = O.sections(O.compress(slab, 0)), O.sections(stream)
        assert so["codes"] == sg["codes"][200:212]
        assert so["slice_crcs"] == sg["slice_crcs"][4 * 200:4 * 212]
    lab = int(t[256, 256, 256].item())
    m = torch.empty(t.shape, dtype=torch.uint8, device="cuda")
    ctx.decompress_into(p, 1, n, 0, -1, lab, m.data_ptr(), 1, m.numel())
    torch.cuda.synchronize()
    assert torch.equal(m.bool(), t.view(torch.int64) == lab)


def test_noise_volumes_permissible_path(ctx):
    # benchmarks/perf.py:68-76 noise cases: randint(0,2000) u32 and binary u8 -> PERMISSIBLE crack format
    from oracle import oracle as O
    import crackle_b200 as cb
    rng = np.random.default_rng(9)
    for v in (np.asfortranarray(rng.integers(0, 2000, (128, 128, 8)).astype(np.uint32)),
              np.asfortranarray(rng.integers(0, 2, (128, 128, 8)).astype(np.uint8))):
        b = ctx.compress(v, 0)
        assert b == O.compress(v, 0)
        assert np.array_equal(cb.decompress(b), v)


def test_config5_slab_2048x2048_u32_dense_labels(ctx):
    # BASELINE.json configs[4] at slab size: 2048x2048 uint32 dense-label slices (cell 16, permutation ids), decompress
    # + single-label extraction.  Slices this large have tens of thousands of crack-graph nodes, so the encoder's replay
    # runs in its 16-bit and global-memory modes; bytes are checked against the oracle.
    import torch
    from oracle import oracle as O
    from crackle_b200 import synth
    import crackle_b200 as cb
    t = synth.jittered_voronoi_torch((2048, 2048, 6), 16, np.uint32, seed=5, id_bits=0, sz_total=1024)
    v = np.asfortranarray(t.cpu().numpy().transpose(2, 1, 0))
    b = ctx.compress(t, 0)
    assert b == O.compress(v, 0)
    n = ctx.compress_ptr(t.data_ptr(), 1, 4, 2048, 2048, 6, True, 0)
    p, _ = ctx.result_device()
    out = torch.empty_like(t)
    ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 4)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), t.view(torch.int32))
    lab = int(v[1024, 1024, 3])
    m = torch.empty(t.shape, dtype=torch.uint8, device="cuda")
    ctx.decompress_into(p, 1, n, 0, -1, lab, m.data_ptr(), 1, m.numel())
    torch.cuda.synchronize()
    assert torch.equal(m.bool(), t.view(torch.int32) == lab)
    assert np.array_equal(cb.decompress_range(b, 2, 4, label=lab), v[:, :, 2:4] == lab)


def test_config3_slab_1024x1024_u64_bytes_vs_oracle(ctx):
    # BASELINE.json configs[2] (the bench volume) on a 24-slice slab: byte-exact against the oracle for orders 0 and 5
    from oracle import oracle as O
    from crackle_b200 import synth
    t = synth.jittered_voronoi_torch((1024, 1024, 24), 24, np.uint64, seed=0, id_bits=40, sz_total=1024)
    v = np.asfortranarray(t.cpu().numpy().transpose(2, 1, 0))
    for order in (0, 5):
        assert ctx.compress(t, order) == O.compress(v, order)


@pytest.mark.parametrize("order", [0, 5])
def test_config2_and_config4_full_size_bytes_vs_compiled_reference(ctx, order):
    # BASELINE.json configs[1] / configs[3] at FULL size: the 512^3 uint64 stream is byte-identical to the stream of the
    # unmodified reference compiled on this box (oracle/_ref, parallel=0 = all host threads); the order-5 model depends on
    # the statistics of the whole volume, so no slab test can stand in for this one.
    from oracle import oracle as O
    from crackle_b200 import synth
    ref = O.ref_module()
    if ref is None:
        pytest.skip("oracle/_ref (compiled reference) is not present on this box")
    t = synth.jittered_voronoi_torch((512, 512, 512), 24, np.uint64, seed=0, id_bits=40)
    v = np.asfortranarray(t.cpu().numpy().transpose(2, 1, 0))
    got = ctx.compress(t, order)
    want = bytes(ref.compress(v, False, True, order, False, True, 0, 0))
    assert len(got) == len(want)
    assert got == want


def _craft_long_escape_run(stream, z=0, pairs=20):
    """Stream surgery: slice z's first chain gets `pairs` x (UP, DOWN) = 'b' followed by `pairs` x (DOWN, UP) = 't' in front
    of its moves -- branches that are pushed and popped without a move, so the cracks (and every crc) stay the same, but
    the code now holds a run of opposite moves far longer than a 16-codepoint word, which the encoders never produce."""
    from oracle import oracle as O
    sec, h = O.sections(stream), O.header(stream)
    assert h["order"] == 0
    code = sec["codes"][z]
    isz = 4 + int.from_bytes(code[:4], "little")
    boc, body = code[:isz], np.frombuffer(code[isz:], dtype=np.uint8)
    bits = np.unpackbits(body, bitorder="little").reshape(-1, 2)
    moves = np.cumsum(bits[:, 0] + 2 * bits[:, 1]) % 4
    new = np.concatenate([np.array([0, 2] * pairs + [2, 0] * pairs), moves])
    diffs = (np.diff(np.concatenate([[0], new]).astype(np.int64)) % 4).astype(np.uint8)
    nb = np.stack([diffs & 1, diffs >> 1], 1).astype(np.uint8).reshape(-1)
    codes = list(sec["codes"])
    codes[z] = boc + np.packbits(nb, bitorder="little").tobytes()
    zi = np.array([len(c) for c in codes], dtype="<u4").tobytes()
    return (sec["header"] + zi + O.crc32c(zi).to_bytes(4, "little") + sec["labels"] + sec["model"] + b"".join(codes) +
            sec["labels_crc"] + sec["slice_crcs"])


def test_crafted_stream_takes_the_serial_decoder(ctx):
    # a valid stream no encoder emits: an opposite-move run longer than a word makes the scan-parallel decoder hand the
    # call to the serial per-slice decoder (k_decode_slices); result identical to the oracle's and the reference's decode
    from oracle import oracle as O
    from crackle_b200 import synth
    import crackle_b200 as cb
    v = synth.jittered_voronoi((96, 80, 3), 16, np.uint32, seed=2, id_bits=20)
    b = O.compress(v, 0)
    for z in (0, 2):
        crafted = _craft_long_escape_run(b, z)
        assert crafted != b and len(crafted) == len(b) + 20
        assert np.array_equal(O.decompress(crafted), v)
        if O.ref_module() is not None:
            assert np.array_equal(O.ref_decompress(crafted), v)
        assert np.array_equal(cb.decompress(crafted).reshape(v.shape), v)
        assert np.array_equal(cb.decompress_range(crafted, z, z + 1), v[:, :, z:z + 1])
    lab = int(v[40, 40, 0])
    assert np.array_equal(cb.decompress(_craft_long_escape_run(b, 0), label=lab), v == lab)


def test_label_crop_and_z_range_for_label(ctx):
    # crackle.decompress(binary, label, crop=True) (codec.py:588-614): only the z-range that holds the label is returned
    import crackle_b200 as cb
    v = np.zeros((40, 30, 12), dtype=np.uint16, order="F")
    v[5:20, 5:20, 3:7] = 7
    v[22:30, 2:9, 6:11] = 9
    b = cb.compress(v)
    assert cb.z_range_for_label(b, 7) == (3, 7)
    assert cb.z_range_for_label(b, 9) == (6, 11)
    assert cb.z_range_for_label(b, 8) == (-1, -1) and cb.z_range_for_label(b, 1 << 20) == (-1, -1)
    m = cb.decompress(b, label=7, crop=True)
    assert m.shape == (40, 30, 4) and np.array_equal(m, v[:, :, 3:7] == 7)
    assert cb.decompress(b, label=8, crop=True).shape == (0, 0, 0)
    full = cb.decompress(b, label=9)
    assert full.shape == v.shape and np.array_equal(full, v == 9)
    assert not cb.decompress(b, label=8).any()


def test_decoder_rejects_markov_orders_it_cannot_hold(ctx):
    import crackle_b200 as cb
    from oracle import oracle as O
    v = np.ones((8, 8, 2), dtype=np.uint8, order="F")
    b = bytearray(O.compress(v, 0))
    fmt = int.from_bytes(b[5:7], "little") | (13 << 9)
    b[5:7] = fmt.to_bytes(2, "little")
    b[28] = O.lib().ckl_oracle_crc8((ctypes.c_char * 23).from_buffer(b, 5), 23)
    with pytest.raises(RuntimeError, match="markov_model_order 13"):
        cb.decompress(bytes(b))


def test_serial_chain_fallback_of_the_decoder():
    # the warp-parallel chain pass hands a slice back to the serial kernel when the reference's push quirk (x == sx) occurs,
    # which no encoder-made stream triggers: the test hook sends every slice down that path (fresh process: read once)
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = ("import numpy as np, crackle_b200 as cb\n"
            "from crackle_b200 import synth\n"
            "from oracle import oracle as O\n"
            "for v in (synth.jittered_voronoi((192, 160, 6), 20, np.uint64, seed=11), synth.random_blobs((61, 47, 6), 9, np.uint16, seed=3),\n"
            "          np.asfortranarray(np.random.default_rng(1).integers(0, 30, (70, 50, 3)).astype(np.uint8))):\n"
            "    for order in (0, 4):\n"
            "        assert np.array_equal(cb.decompress(O.compress(v, order)).reshape(v.shape), v)\n"
            "print('SERIAL CHAIN OK')\n")
    env = dict(os.environ, CKL_TEST_SERIAL_CHAIN="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "SERIAL CHAIN OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
