"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/_fastcrackle_ref, built by
oracle/build_ref.sh from /root/reference).  Run in the build container only:  python oracle/make_golden.py
Each vector stores the input array, the compress arguments and the reference's exact .ckl bytes (plus, for a
few, reference decompress outputs for z-ranges / label masks).  The committed vectors travel to the GPU box,
where /root/reference does not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from crackle_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cases():
    rng = np.random.default_rng(1234)
    yield "kat1_2x2_perm", np.asfortranarray(np.array([[1, 2], [3, 4]], dtype=np.uint8).T)
    # automated_test.py:909-919 and :926-931 (test_spurious_branch_elimination)
    s1 = np.array([
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 1, 1, 2, 2, 0, 0, 0, 0],
        [0, 0, 1, 1, 2, 2, 0, 0, 0, 0],
        [0, 0, 4, 4, 3, 3, 0, 0, 0, 0],
        [0, 0, 4, 4, 3, 3, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    ], dtype=np.uint8).T
    yield "spurious1_10x9", np.asfortranarray(s1)
    s2 = np.array([
        [0, 139, 139, 139, 139],
        [0, 139, 0, 139, 139],
        [0, 161, 0, 0, 161],
        [161, 161, 161, 161, 161],
    ], dtype=np.uint8).T
    yield "spurious2_5x4", np.asfortranarray(s2)
    yield "zeros_u16_4x3x2", np.zeros((4, 3, 2), dtype=np.uint16, order="F")
    isl = np.zeros((6, 6), dtype=np.uint8, order="F"); isl[2:4, 2:4] = 7
    yield "island_6x6", isl
    yield "voronoi_u64_96x80x5", synth.jittered_voronoi((96, 80, 5), 16, np.uint64, seed=1)
    yield "voronoi_u32_130x67x3", synth.jittered_voronoi((130, 67, 3), 12, np.uint32, seed=2, id_bits=16)
    yield "voronoi_u16_64x64x4", synth.jittered_voronoi((64, 64, 4), 9, np.uint16, seed=3, id_bits=16)
    yield "noise2000_u32_40x33x3", np.asfortranarray(rng.integers(0, 2000, (40, 33, 3)).astype(np.uint32))
    yield "binary_noise_u8_50x50x2", np.asfortranarray(rng.integers(0, 2, (50, 50, 2)).astype(np.uint8))
    yield "noise3_u8_33x47x3", np.asfortranarray(rng.integers(0, 3, (33, 47, 3)).astype(np.uint8))
    yield "arange_u32_10x10x3", np.asfortranarray(np.arange(300, dtype=np.uint32).reshape((10, 10, 3), order="F"))
    yield "ones_u64_7x5x3", np.ones((7, 5, 3), dtype=np.uint64, order="F")
    yield "row_1x37", np.asfortranarray(rng.integers(0, 3, (1, 37)).astype(np.uint8))
    yield "col_37x1", np.asfortranarray(rng.integers(0, 3, (37, 1)).astype(np.uint8))
    yield "blobs_c_order_u32_31x29x4", np.ascontiguousarray(synth.random_blobs((31, 29, 4), 9, np.uint32, seed=5))
    big = synth.random_blobs((70, 65, 2), 11, np.uint64, seed=7) * np.uint64((1 << 50) + 12345)
    yield "blobs_u64_wide_70x65x2", np.asfortranarray(big)
    # comb: deep revisit stack
    comb = np.zeros((64, 40), dtype=np.uint8, order="F"); comb[::2, 5:35] = 1; comb[:, 20] = 2
    yield "comb_64x40", comb
    chk = np.asfortranarray((np.add.outer(np.arange(24), np.arange(20)) % 2).astype(np.uint8))
    yield "checker_24x20", chk


def main():
    ref = O.ref_module()
    assert ref is not None, "compiled reference missing: run oracle/build_ref.sh"
    os.makedirs(OUT, exist_ok=True)
    for name, arr in cases():
        rec = {"input": arr, "f_order": np.array(arr.flags.f_contiguous)}
        for order in (0, 1, 5):
            rec[f"ckl_order{order}"] = np.frombuffer(O.ref_compress(arr, order), dtype=np.uint8)
        b = bytes(rec["ckl_order0"])
        if arr.ndim == 3 and arr.shape[2] >= 3:
            rec["z1_2"] = O.ref_decompress(b, 1, 2)
        lab = int(arr.reshape(-1, order="F")[arr.size // 2])
        rec["label"] = np.array(lab, dtype=np.uint64)
        rec["mask"] = O.ref_decompress(b, 0, -1, lab)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, arr.shape, arr.dtype, {k: len(v) for k, v in rec.items() if k.startswith("ckl")})


if __name__ == "__main__":
    main()
