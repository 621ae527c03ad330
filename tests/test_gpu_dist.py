"""z-sharded compress on real CUDA contexts.  On a 1-GPU box the two "ranks" are two threads with their own ckl_ctx on
cuda:0 and an in-process stand-in for torch.distributed; with >= 2 GPUs the real NCCL path is run under torchrun."""
import os
import queue
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


class ThreadGroup:
    """Minimal blocking collectives between threads (all_gather / all_reduce / send / recv / broadcast)."""

    class ReduceOp:
        SUM = "sum"

    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world
        self.q = {(s, d): queue.Queue() for s in range(world) for d in range(world)}
        self.local = threading.local()

    def view(self, rank):
        g = self

        class V:
            ReduceOp = ThreadGroup.ReduceOp

            def get_rank(self_):
                return rank

            def get_world_size(self_):
                return g.world

            def all_gather(self_, out, t):
                torch.cuda.synchronize()
                g.slots[rank] = t.clone()
                g.bar.wait()
                for i in range(g.world):
                    out[i].copy_(g.slots[i])
                torch.cuda.synchronize()
                g.bar.wait()

            def all_gather_into_tensor(self_, out, t):
                torch.cuda.synchronize()
                g.slots[rank] = t.clone()
                g.bar.wait()
                n = t.numel()
                for i in range(g.world):
                    out[i * n:(i + 1) * n].copy_(g.slots[i])
                torch.cuda.synchronize()
                g.bar.wait()

            def all_reduce(self_, t, op=None):
                torch.cuda.synchronize()
                g.slots[rank] = t.clone()
                g.bar.wait()
                acc = g.slots[0].clone()
                for i in range(1, g.world):
                    acc += g.slots[i]
                t.copy_(acc)
                torch.cuda.synchronize()
                g.bar.wait()

            def send(self_, t, dst):
                torch.cuda.synchronize()
                g.q[(rank, dst)].put(t.clone())

            def recv(self_, t, src):
                t.copy_(g.q[(src, rank)].get(timeout=60))
                torch.cuda.synchronize()

            def broadcast(self_, t, src):
                torch.cuda.synchronize()
                if rank == src:
                    g.slots[src] = t.clone()
                g.bar.wait()
                if rank != src:
                    t.copy_(g.slots[src])
                torch.cuda.synchronize()
                g.bar.wait()
        return V()


def _mixed_volume(shape, world):
    """rank 0's slab is 60 % noise (its own statistics say PERMISSIBLE), the rest is blocky: the global decision is
    IMPERMISSIBLE, so rank 0 must re-encode after the metadata exchange"""
    sx, sy, sz = shape
    g = torch.Generator().manual_seed(5)
    vol = (torch.arange(sz)[:, None, None] // 2 * 7 + torch.arange(sy)[None, :, None] // 16 * 3 + torch.arange(sx)[None, None, :] // 16 + 1)
    vol = vol.to(torch.int64).contiguous()
    per = sz // world
    noise = torch.randint(1, 1 << 30, (per, sy, sx), generator=g, dtype=torch.int64)
    m = torch.rand((per, sy, sx), generator=g) < 0.6
    vol[:per][m] = noise[m]
    return vol.view(torch.uint64).cuda()


@pytest.mark.parametrize("order", [0, 5])
@pytest.mark.parametrize("world,kind", [(2, "voronoi"), (3, "voronoi"), (2, "mixed"), (3, "mixed")])
def test_sharded_equals_monolithic_threads(order, world, kind):
    import crackle_b200 as cb
    from crackle_b200 import synth
    from crackle_b200.dist import ShardedCodec
    from oracle import oracle as O
    shape = (160, 128, 4 * world + 1)
    szt = shape[2]
    whole = synth.jittered_voronoi_torch(shape, 14, np.uint64, seed=8, id_bits=40) if kind == "voronoi" else _mixed_volume(shape, world)
    want = cb.default_context().compress(whole, order)
    assert want == O.compress(np.asfortranarray(whole.cpu().numpy().transpose(2, 1, 0)), order)
    group = ThreadGroup(world)
    results, errors = {}, []

    def run(rank):
        try:
            ctx = cb.Context(0)
            per = szt // world
            z0 = rank * per
            z1 = szt if rank == world - 1 else z0 + per
            vol = whole[z0:z1].contiguous()
            job = ShardedCodec(ctx, group.view(rank))
            s = job.compress(vol, z0, szt, order)
            results["stream", rank] = bytes(s.cpu().numpy().tobytes())     # every rank holds the complete stream
            results["collectives", rank] = job.collectives
            out = torch.empty_like(vol)
            job.decompress_shard(s, z0, z1, out)
            torch.cuda.synchronize()
            results[rank] = torch.equal(out.view(torch.uint8), vol.view(torch.uint8))
        except Exception as e:  # noqa: BLE001
            errors.append((rank, repr(e)))
            try:
                group.bar.abort()
            except Exception:
                pass

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=180)
    assert not errors, errors
    for r in range(world):
        assert results["stream", r] == want
        # metadata (+1 when a shard re-encodes), unique tables, packed blocks; order > 0 adds statistics + code sizes
        assert results["collectives", r] == 4 + (kind == "mixed") + 2 * (order > 0)
    assert all(results[r] for r in range(world))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs for the NCCL path")
def test_sharded_nccl_torchrun():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "bringup", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
