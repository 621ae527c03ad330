"""GPU bring-up helper: runs every golden vector + a few oracle-checked random volumes through the CUDA path and
prints, per case, which stream section differs.  Usage (on a GPU box): python tests/bringup/gpu_debug.py [name-filter]"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import crackle_b200 as cb  # noqa: E402
from oracle import oracle as O  # noqa: E402
from crackle_b200 import synth  # noqa: E402


def diff_sections(got, want):
    try:
        sg, sw = O.sections(got), O.sections(want)
    except Exception as e:
        return f"unparseable ({e}); len got {len(got)} want {len(want)}"
    msgs = []
    for k in sw:
        if sg[k] != sw[k]:
            if k == "codes":
                for z, (a, b) in enumerate(zip(sg[k], sw[k])):
                    if a != b:
                        msgs.append(f"codes[z={z}] got({len(a)}) {a.hex()[:120]} want({len(b)}) {b.hex()[:120]}")
                        break
                if len(sg[k]) != len(sw[k]):
                    msgs.append(f"codes count {len(sg[k])} vs {len(sw[k])}")
            else:
                msgs.append(f"{k}: got {bytes(sg[k]).hex()[:100]} want {bytes(sw[k]).hex()[:100]}")
    return "; ".join(msgs) if msgs else f"len got {len(got)} want {len(want)}"


def main():
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    gdir = os.path.join(ROOT, "tests", "golden")
    nfail = 0
    ctx = cb.default_context()
    cases = []
    for f in sorted(os.listdir(gdir)):
        if f.endswith(".npz") and filt in f:
            g = np.load(os.path.join(gdir, f))
            a = g["input"]
            a = np.asfortranarray(a) if bool(g["f_order"]) else np.ascontiguousarray(a)
            cases.append((f[:-4], a, {o: bytes(g[f"ckl_order{o}"]) for o in (0, 1, 5)}))
    rng = np.random.default_rng(5)
    extra = [("voronoi_u64_256x256x8", synth.jittered_voronoi((256, 256, 8), 24, np.uint64, seed=0)),
             ("voronoi_u32_300x200x5", synth.jittered_voronoi((300, 200, 5), 20, np.uint32, seed=1, id_bits=16)),
             ("noise_u8_100x90x4", np.asfortranarray(rng.integers(0, 2, (100, 90, 4)).astype(np.uint8))),
             ("noise_u32_64x64x3", np.asfortranarray(rng.integers(0, 2000, (64, 64, 3)).astype(np.uint32)))]
    for name, a in extra:
        if filt in name:
            cases.append((name, a, {o: O.compress(a, o) for o in (0, 1, 5)}))
    for name, a, want in cases:
        for order in (0, 1, 5):
            tag = f"{name} order{order}"
            try:
                got = ctx.compress(a, order)
                if got != want[order]:
                    nfail += 1
                    print("FAIL compress", tag, diff_sections(got, want[order]))
                else:
                    print("ok   compress", tag, len(got))
            except Exception as e:
                nfail += 1
                print("EXC  compress", tag, repr(e))
                traceback.print_exc()
            try:
                d = cb.decompress(want[order])
                if not np.array_equal(d.reshape(a.shape), a):
                    nfail += 1
                    bad = np.argwhere(d.reshape(a.shape) != a)
                    print("FAIL decompress", tag, "mismatches", len(bad), "first", bad[:3].tolist())
                else:
                    print("ok   decompress", tag)
            except Exception as e:
                nfail += 1
                print("EXC  decompress", tag, repr(e))
        try:
            lab = int(a.reshape(-1, order="F")[a.size // 2])
            m = cb.decompress(want[0], label=lab)
            if not np.array_equal(m.reshape(a.shape), a == lab):
                nfail += 1
                print("FAIL mask", name)
        except Exception as e:
            nfail += 1
            print("EXC  mask", name, repr(e))
    print("TOTAL FAILURES", nfail)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
