"""Pins the CPU oracle (oracle/crackle_oracle.c) against (a) the known-answer vectors of SURVEY.md 8(c),
(b) the golden vectors generated from the compiled reference (tests/golden, oracle/make_golden.py) and
(c) the compiled reference itself when oracle/_ref is present."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import oracle as O

# captured from the compiled reference (x-fastest inputs, order as noted)
KATS = {
    "kat1_2x2_perm": (0, "63726b6c0190000200000002000000010000001f11000000000000009c050000008cd000ee0400000000000000"
                         "0102030404000102030100000000992535bdad3d5b83"),
    "spurious1_10x9": (0, "63726b6c0180000a00000009000000010000001f130000000000000090110000004250467c0500000000000000"
                          "00010203040500010204030400000001020102011190303b930119094b28cb29dacec431"),
    "spurious2_5x4": (0, "63726b6c0180000500000004000000010000001f1000000000000000000e000000533a667a03000000000000"
                         "00008ba1040001000204000000010001019234e3798fb69dea3d8071e851b6"),
    "zeros_u16_4x3x2": (0, "63726b6c0181000400000003000000020000001f0d00000000000000e005000000050000008bd66e42010000"
                           "000000000000010100000100000000010000000047a58df4b93a8c28b93a8c28"),
    "island_6x6": (0, "63726b6c0180000600000006000000010000001f0d00000000000000880b00000018a101dc0200000000000000"
                      "0007020001040000000102010232330beabcfb792c7ef976"),
}


@pytest.mark.parametrize("name", sorted(KATS))
def test_kat(name):
    order, hexs = KATS[name]
    g = load_golden(name)
    assert O.compress(g["input"], order).hex() == hexs
    assert bytes(g[f"ckl_order{order}"]).hex() == hexs


def test_kat_order1():
    g = load_golden("spurious1_10x9")
    assert O.compress(g["input"], 1).hex() == (
        "63726b6c0180020a00000009000000010000001f13000000000000005d0f000000eb9023a70500000000000000000102030405"
        "000102040367a80104000000010201022928ad591951424b28cb29dacec431")


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("order", [0, 1, 5])
def test_compress_matches_golden(name, order):
    g = load_golden(name)
    a = g["input"]
    a = np.asfortranarray(a) if bool(g["f_order"]) else np.ascontiguousarray(a)
    assert O.compress(a, order) == bytes(g[f"ckl_order{order}"])


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("order", [0, 5])
def test_decompress_matches_golden(name, order):
    g = load_golden(name)
    a = g["input"]
    b = bytes(g[f"ckl_order{order}"])
    d = O.decompress(b)
    assert np.array_equal(d.reshape(a.shape), a)
    assert np.array_equal(O.decompress(b, label=int(g["label"])), g["mask"])
    if "z1_2" in g:
        assert np.array_equal(O.decompress(b, 1, 2), g["z1_2"])


def test_crc32c_kat():
    # 48 zero bytes -> b93a8c28 (little-endian in the stream => 0x288c3ab9)
    assert O.crc32c(bytes(48)) == int.from_bytes(bytes.fromhex("b93a8c28"), "little")
    assert O.crc32c(b"123456789") == 0xE3069283


def test_empty_volume_is_header_only():
    a = np.zeros((0, 5, 3), dtype=np.uint32, order="F")
    b = O.compress(a)
    assert len(b) == 29 and b[:4] == b"crkl"


def test_corrupt_slice_crc_detected():
    g = load_golden("voronoi_u16_64x64x4")
    b = bytearray(bytes(g["ckl_order0"]))
    b[-1] ^= 1
    with pytest.raises(RuntimeError):
        O.decompress(bytes(b))


@pytest.mark.skipif(O.ref_module() is None, reason="compiled reference (oracle/_ref) not present")
def test_against_compiled_reference_random():
    rng = np.random.default_rng(7)
    from crackle_b200 import synth
    for trial in range(24):
        sx, sy, sz = (int(v) for v in rng.integers(1, 48, 3))
        dt = [np.uint8, np.uint16, np.uint32, np.uint64][trial % 4]
        if trial % 3 == 0:
            a = rng.integers(0, 3, (sx, sy, sz)).astype(dt)
        elif trial % 3 == 1:
            a = rng.integers(0, 250, (sx, sy, sz)).astype(dt)
        else:
            a = synth.random_blobs((sx, sy, sz), 6, dt, seed=trial)
        a = np.asfortranarray(a)
        for order in (0, 2, 5):
            r = O.ref_compress(a, order)
            assert O.compress(a, order) == r
            assert np.array_equal(O.decompress(r), O.ref_decompress(r))
