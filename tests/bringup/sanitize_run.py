"""Small end-to-end pass over every entry point of the C-ABI, meant to be run under compute-sanitizer
(tools/sanitize.sh: memcheck, racecheck, initcheck, synccheck).  numpy in / numpy out, no torch -- only the library's own
kernels are instrumented.  Results are checked against the oracle so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import crackle_b200 as cb                      # noqa: E402
from crackle_b200 import synth                 # noqa: E402
from oracle import oracle as O                 # noqa: E402  (the checker)


def main():
    rng = np.random.default_rng(2)
    vols = [synth.jittered_voronoi((96, 80, 5), 12, np.uint64, seed=1, id_bits=40),
            synth.jittered_voronoi((256, 64, 3), 10, np.uint32, seed=2, id_bits=30),     # sx % 256 == 0: bulk-copy edges / paint
            synth.random_blobs((61, 47, 4), 9, np.uint16, seed=3),
            np.asfortranarray(rng.integers(0, 3, (70, 50, 4)).astype(np.uint8)),          # PERMISSIBLE crack format
            np.zeros((40, 40, 3), np.uint32, order="F")]
    small = os.environ.get("SANITIZE_SMALL") == "1"      # racecheck instruments every shared-memory access: a shorter pass
    if small:
        vols = vols[:2]
    for chunks in ((1,) if small else (1, 2)):
        ctx = cb.Context(0)
        ctx.set_chunks(chunks)
        for v in vols:
            for order in (0, 3):
                want = O.compress(v, order)
                b = ctx.compress(v, order)
                assert b == want, (v.shape, order, chunks)
                out = ctx.decompress(b).reshape(v.shape, order="F")
                assert np.array_equal(out, v)
            b = O.compress(v, 0)
            lab = int(v[v.shape[0] // 2, v.shape[1] // 2, 1])
            m = ctx.decompress(b, label=lab).reshape(v.shape, order="F")
            assert np.array_equal(m.view(bool), v == lab)
            part = ctx.decompress(b, 1, 3).reshape(v.shape[:2] + (2,), order="F")
            assert np.array_equal(part, v[:, :, 1:3])
            labels, counts, sums, bbox = ctx.label_stats(b)
            u, c = np.unique(v, return_counts=True)
            assert np.array_equal(labels, u.astype(np.uint64)) and np.array_equal(counts, c.astype(np.uint64))
            for conn in (4, 6):
                g = ctx.voxel_connectivity_graph(b, 0, -1, conn)
                assert g.shape == v.shape and (conn == 4 or v.shape[2] == 1 or (g[:, :, 0] & 32).all())
            ref = O.ref_module()
            if ref is not None:
                assert np.array_equal(g, np.asarray(ref.voxel_connectivity_graph(b, 0, -1, 1, 6)))
            r2 = ctx.reencode(b, 2)
            if ref is not None:
                assert r2 == bytes(ref.reencode_markov(b, 2, 1))
            assert np.array_equal(O.decompress(r2).reshape(v.shape, order="F"), v)
            parts = [O.compress(np.asfortranarray(v[:, :, :2]), 0), O.compress(np.asfortranarray(v[:, :, 2:]), 0)]
            if len({O.header(p)["crack_format"] for p in parts + [b]}) == 1:
                assert ctx.zstack(parts) == b
            assert ctx.zslice(b, 1, 3) is not None
        ctx.close()
    print("sanitize_run ok: launches", cb.codec.launch_count())


if __name__ == "__main__":
    main()
