"""world_size-2 gloo tests (CPU) of the z-sharded host logic in crackle_b200/dist.py: global decisions from the
per-shard summaries, label-table merge, piece gathering and stream assembly.  The per-shard compute is supplied by
an oracle-backed fake backend (tests may use the oracle; the product backend is CUDA and is covered by -m gpu tests).
The reference's own proof of this property is test_zstack_ones (automated_test.py:449-487)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle as O


class OracleShardBackend:
    device = torch.device("cpu")

    def begin(self, vol):
        self.vol = np.asfortranarray(vol)
        sx, sy, sz = self.vol.shape
        self.shape, self.width = (sx, sy, sz), self.vol.dtype.itemsize
        flat = self.vol.reshape(-1, order="F")
        return dict(max_label=int(flat.max()), pairs=int((flat[1:] == flat[:-1]).sum()), first_voxel=int(flat[0]),
                    last_voxel=int(flat[-1]), voxels=flat.size)

    def encode(self, permissible, stored_width, order):
        assert order == 0, "the CPU fake covers order 0 (order > 0 needs the global model; covered on GPU)"
        b = O.compress(self.vol, 0)
        h = O.header(b)
        assert h["crack_format"] == int(permissible), "test data must not flip crack format per slab"
        self.sec = O.sections(b)
        lab = self.sec["labels"]
        sx, sy, sz = self.shape
        nu = int.from_bytes(lab[:8], "little")
        sw = h["stored_width"]
        uniq = np.frombuffer(lab[8:8 + nu * sw], dtype=f"<u{sw}").astype(np.uint64)
        cw = 1 if sx * sy <= 0xFF else 2 if sx * sy <= 0xFFFF else 4
        self.nz = np.frombuffer(lab[8 + nu * sw:8 + nu * sw + cw * sz], dtype=f"<u{cw}").astype(np.uint64)
        kw = 1 if nu <= 0xFF else 2 if nu <= 0xFFFF else 4
        keys = np.frombuffer(lab[8 + nu * sw + cw * sz:], dtype=f"<u{kw}")
        self.mapping = uniq[keys]
        self.uniq = uniq
        ncp = sum(len(c) for c in self.sec["codes"])
        return nu, len(self.mapping), ncp

    def unique(self):
        return torch.from_numpy(self.uniq.view(np.int64).copy())

    def sort_unique(self, t, key_bytes=8):
        return torch.from_numpy(np.unique(t.numpy().view(np.uint64)).view(np.int64).copy())

    def finish(self, guniq, gstats):
        g = guniq.numpy().view(np.uint64)
        kw = 1 if len(g) <= 0xFF else 2 if len(g) <= 0xFFFF else 4
        self.keys = np.searchsorted(g, self.mapping).astype(f"<u{kw}").tobytes()
        self.codes = b"".join(self.sec["codes"])
        return dict(keys_bytes=len(self.keys), codes_bytes=len(self.codes), sz_local=self.shape[2])

    def info(self):
        cs = sum(len(c) for c in self.sec["codes"])
        return dict(n_unique_local=len(self.uniq), n_components=len(self.mapping), n_codepoints=cs, codes_bytes_order0=cs,
                    sz_local=self.shape[2], runs=0, keys_bytes=0, codes_bytes=0)

    def pack(self, buf):
        """layout of ckl_shard_pack: N_z u32[sz] | code sizes u32[sz] | crcs u32[sz] | keys | codes"""
        cs = np.array([len(c) for c in self.sec["codes"]], dtype="<u4")
        cr = np.frombuffer(self.sec["slice_crcs"], dtype="<u4")
        blob = self.nz.astype("<u4").tobytes() + cs.tobytes() + cr.tobytes() + self.keys + self.codes
        buf[: len(blob)] = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy())
        return len(blob)

    def assemble(self, gathered, blocks, guniq, data_width, stored, permissible, fortran_order, order, sx, sy):
        """numpy restatement of ckl_shard_assemble (crackle.hpp:171-216, labels.hpp:123-152)"""
        from crackle_b200.dist import header_bytes, byte_width
        G = gathered.numpy().tobytes()
        nz, cs, cr, keys, codes = b"", b"", b"", b"", b""
        cw = byte_width(sx * sy)
        sz = 0
        for off, szl, ncomp, kb, cb in blocks:
            small = np.frombuffer(G[off:off + 12 * szl], dtype="<u4")
            nz += small[:szl].astype(f"<u{cw}").tobytes()
            cs += small[szl:2 * szl].tobytes()
            cr += small[2 * szl:].tobytes()
            keys += G[off + 12 * szl: off + 12 * szl + kb]
            codes += G[off + 12 * szl + kb: off + 12 * szl + kb + cb]
            sz += szl
        g = guniq.numpy().view(np.uint64)
        labels = len(g).to_bytes(8, "little") + g.astype(f"<u{stored}").tobytes() + nz + keys
        head = header_bytes(data_width, stored, permissible, fortran_order, order, sx, sy, sz, len(labels))
        out = (head + cs + O.crc32c(cs).to_bytes(4, "little") + labels + codes + O.crc32c(labels).to_bytes(4, "little") + cr)
        return torch.from_numpy(np.frombuffer(out, dtype=np.uint8).copy())

    def empty_bytes(self, n):
        return torch.zeros(n, dtype=torch.uint8)


def _volume(kind):
    from crackle_b200 import synth
    if kind == "voronoi_u64":
        return synth.jittered_voronoi((48, 40, 10), 9, np.uint64, seed=3)
    if kind == "blobs_u16":
        return synth.random_blobs((37, 29, 7), 6, np.uint16, seed=2)
    v = np.ones((16, 16, 6), dtype=np.uint32, order="F")     # test_zstack_ones: uniform volume
    return v


def _worker(rank, world, port, kind, split, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crackle_b200.dist import ShardedCodec
        vol = _volume(kind)
        bounds = [0] + list(split) + [vol.shape[2]]
        z0, z1 = bounds[rank], bounds[rank + 1]
        codec = ShardedCodec(None, dist, backend=OracleShardBackend())
        out = codec.compress(np.asfortranarray(vol[:, :, z0:z1]), z0, vol.shape[2], 0)
        q.put((rank, bytes(out.numpy().tobytes()), codec.collectives))      # EVERY rank holds the complete stream
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind,split", [("voronoi_u64", (5,)), ("voronoi_u64", (1,)), ("blobs_u16", (4,)), ("ones", (3,))])
def test_two_rank_sharded_compress_is_byte_identical(kind, split):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, split, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = O.compress(_volume(kind), 0)
    assert sorted(r for r, _, _ in got) == [0, 1]
    for _, stream, ncoll in got:
        assert stream == want
        assert ncoll == 4            # metadata, unique tables, code sizes, packed blocks: no per-peer send/recv, no broadcast


def test_header_helper_matches_oracle():
    from crackle_b200.dist import header_bytes
    v = _volume("voronoi_u64")
    b = O.compress(v, 0)
    h = O.header(b)
    assert header_bytes(8, h["stored_width"], h["crack_format"], 1, 0, 48, 40, 10, h["num_label_bytes"]) == b[:29]
