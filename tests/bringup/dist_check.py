"""multi-GPU parity check (run under torchrun): z-sharded compress == single-GPU compress of the whole volume,
orders 0 and 5, checked on EVERY rank (each holds the complete stream after the one padded all-gather); every rank then
decodes its own z-range of its own copy."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import crackle_b200 as cb  # noqa: E402
from crackle_b200 import synth  # noqa: E402
from crackle_b200.dist import ShardedCodec  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cb.Context(local)
    job = ShardedCodec(ctx, dist)
    mono = cb.Context(local)
    ok = True
    for shape, cell, dt, bits in (((256, 192, 8 * world + 3), 20, np.uint64, 40), ((130, 97, 2 * world), 12, np.uint16, 15)):
        sx, sy, szt = shape
        per = szt // world
        z0 = rank * per
        z1 = szt if rank == world - 1 else z0 + per
        vol = synth.jittered_voronoi_torch((sx, sy, z1 - z0), cell, dt, seed=1, id_bits=bits, z0=z0, sz_total=szt)
        whole = synth.jittered_voronoi_torch(shape, cell, dt, seed=1, id_bits=bits)
        torch.cuda.synchronize()
        for order in (0, 5):
            s = job.compress(vol, z0, szt, order)
            if True:
                want = mono.compress(whole, order)
                got = bytes(s.cpu().numpy().tobytes())
                same = got == want
                print(f"shape {shape} {np.dtype(dt).name} order {order}: sharded == monolithic: {same} ({len(got)} bytes)", flush=True)
                ok &= same
                if not same:
                    from oracle import oracle as O
                    sg, sw = O.sections(got), O.sections(want)
                    for k in sw:
                        if sg[k] != sw[k]:
                            if k == "codes":
                                for z, (a, b) in enumerate(zip(sg[k], sw[k])):
                                    if a != b:
                                        print(f"  codes[z={z}] got({len(a)}) {a.hex()[:160]}\n             want({len(b)}) {b.hex()[:160]}", flush=True)
                            else:
                                i = next((j for j in range(min(len(sg[k]), len(sw[k]))) if sg[k][j] != sw[k][j]), -1)
                                print(f"  section {k}: len {len(sg[k])} vs {len(sw[k])}, first diff at {i}: {bytes(sg[k][i:i+16]).hex()} vs {bytes(sw[k][i:i+16]).hex()}", flush=True)
            out = torch.empty_like(vol)
            try:
                job.decompress_shard(s, z0, z1, out)
            except RuntimeError as e:
                print(f"rank {rank}: decode error {e}", flush=True)
            torch.cuda.synchronize()
            good = torch.equal(out.view(torch.uint8), vol.view(torch.uint8))
            if not good:
                print(f"rank {rank}: decode mismatch shape {shape} order {order}", flush=True)
            ok &= good
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
