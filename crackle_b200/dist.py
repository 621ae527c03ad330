"""z-sharded multi-GPU compress / decompress: one process per GPU, torch.distributed for the plumbing.

Slices are independent, so the volume shards along z (rank r owns a contiguous z-slab).  Volume-proportional data
never leaves its GPU; the only exchanges are the small global couplings of the stream
(the reference's own recipe for stitching independently compressed slabs is
crackle/operations.py:258-295 `_zstack_flat_labels` and :424-548 `zstack`):

  * max label -> stored width, pixel pairs -> crack format       (all_gather of a 64-byte summary)
  * sorted unique label table                                     (all_gather of per-shard unique sets, merge+unique)
  * markov statistics                                             (all_reduce SUM of uint32[4^order*4], wraps mod 2^32)
  * N_z / code sizes / slice crcs                                 (all_gather of 3 small arrays)
  * keys + crack codes -> one stream on rank 0                    (point-to-point send/recv straight into place)

The per-shard compute is behind a small backend interface so the host logic can be exercised on CPU with gloo
(tests/test_dist_cpu.py supplies an oracle-backed fake); the product backend is `CudaShardBackend` (C-ABI)."""
import ctypes
import os
import time

import numpy as np
import torch

from . import _capi


def byte_width(x: int) -> int:      # lib.hpp:236-247
    return 1 if x <= 0xFF else 2 if x <= 0xFFFF else 4 if x <= 0xFFFFFFFF else 8


def crc8_header(b: bytes) -> int:   # crc.hpp:23-37
    c = 0xFF
    for v in b:
        c ^= v
        for _ in range(8):
            c = ((c >> 1) ^ 0xE7) if (c & 1) else (c >> 1)
    return c


def header_bytes(data_width, stored_width, crack_format, fortran, order, sx, sy, sz, num_label_bytes) -> bytes:
    """29-byte v1 header (header.hpp:206-267)."""
    lg = {1: 0, 2: 1, 4: 2, 8: 3}
    fmt = lg[data_width] | (lg[stored_width] << 2) | (int(crack_format) << 4) | (int(bool(fortran)) << 7) | ((order & 15) << 9)
    b = bytearray(b"crkl")
    b += bytes([1]) + fmt.to_bytes(2, "little") + int(sx).to_bytes(4, "little") + int(sy).to_bytes(4, "little")
    b += int(sz).to_bytes(4, "little") + bytes([31]) + int(num_label_bytes).to_bytes(8, "little")
    b += bytes([crc8_header(bytes(b[5:28]))])
    return bytes(b)


def _i64(v) -> int:
    """uint64 value -> the int64 with the same bit pattern (collectives carry int64)."""
    return int(np.array(int(v), dtype=np.uint64).view(np.int64))


def model_bytes(order: int) -> int:  # header.hpp:284-297
    return 0 if order == 0 else (4 ** order * 5 + 4) // 8


class CudaShardBackend:
    """Per-GPU shard stages through the C-ABI (ckl_shard_*)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.L = _capi.lib()
        self.device = torch.device("cuda", ctx.device)

    def _check(self, rc):
        self.ctx._check(rc)

    def begin(self, vol):
        """vol: CUDA tensor (sz_local, sy, sx) of an unsigned dtype."""
        s = _capi.ShardSummary()
        sz, sy, sx = vol.shape
        self.shape = (sx, sy, sz)
        self.width = vol.element_size()
        self._check(self.L.ckl_shard_begin(self.ctx._h, vol.data_ptr(), 1, self.width, sx, sy, sz, ctypes.byref(s)))
        return dict(max_label=s.max_label, pairs=s.pairs, first_voxel=s.first_voxel, last_voxel=s.last_voxel, voxels=s.voxels)

    def encode(self, permissible, stored_width, order):
        nu, nc, ncp = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        self.order = order
        self._check(self.L.ckl_shard_encode(self.ctx._h, int(permissible), int(stored_width), int(order), ctypes.byref(nu),
                                            ctypes.byref(nc), ctypes.byref(ncp)))
        self.n_unique_local, self.ncomp = nu.value, nc.value
        return nu.value, nc.value, ncp.value

    def unique(self):
        t = torch.empty(max(self.n_unique_local, 1), dtype=torch.int64, device=self.device)
        self._check(self.L.ckl_shard_unique(self.ctx._h, t.data_ptr(), 1))
        return t[: self.n_unique_local]

    def stats(self):
        t = torch.empty(4 ** self.order * 4, dtype=torch.int32, device=self.device)
        self._check(self.L.ckl_shard_stats(self.ctx._h, t.data_ptr(), 1))
        return t

    def sort_unique(self, t, key_bytes=8):
        n = ctypes.c_uint64()
        t = t.contiguous()
        self._check(self.L.ckl_sort_unique_u64(self.ctx._h, t.data_ptr(), t.numel(), key_bytes, ctypes.byref(n)))
        return t[: n.value]

    def finish(self, global_unique, global_stats):
        p = _capi.ShardPieces()
        self._check(self.L.ckl_shard_finish(self.ctx._h, global_unique.data_ptr(), 1, global_unique.numel(),
                                            global_stats.data_ptr() if global_stats is not None else None, 1, ctypes.byref(p)))
        self.pieces = p
        return dict(keys_bytes=p.keys_bytes, codes_bytes=p.codes_bytes, sz_local=p.sz_local)

    def small_pieces(self):
        """-> (components_per_slice u64[sz], code_sizes u32[sz], slice_crcs u32[sz]) as numpy"""
        sz = int(self.pieces.sz_local)
        nz = np.zeros(sz, dtype=np.uint64)
        cs = np.zeros(sz, dtype=np.uint32)
        cr = np.zeros(sz, dtype=np.uint32)
        self._check(self.L.ckl_shard_fetch(self.ctx._h, None, nz.ctypes.data, cs.ctypes.data, cr.ctypes.data, None, 0))
        return nz, cs, cr

    def big_pieces(self, keys_dst, codes_dst):
        """copies keys / codes into the given uint8 CUDA tensors (views into the final stream on rank 0)"""
        self._check(self.L.ckl_shard_fetch(self.ctx._h, keys_dst.data_ptr() if keys_dst is not None and keys_dst.numel() else None,
                                           None, None, None,
                                           codes_dst.data_ptr() if codes_dst is not None and codes_dst.numel() else None, 1))

    def stored_model(self):
        n = ctypes.c_uint64()
        self._check(self.L.ckl_shard_model(self.ctx._h, None, 0, 0, ctypes.byref(n)))
        buf = np.zeros(max(n.value, 1), dtype=np.uint8)
        if n.value:
            self._check(self.L.ckl_shard_model(self.ctx._h, buf.ctypes.data, 0, n.value, ctypes.byref(n)))
        return buf[: n.value].tobytes()

    def crc32c(self, t):
        out = ctypes.c_uint32()
        self._check(self.L.ckl_crc32c(self.ctx._h, t.data_ptr(), 1, t.numel(), ctypes.byref(out)))
        return out.value

    def empty_bytes(self, n):
        return torch.empty(n, dtype=torch.uint8, device=self.device)

    def to_device(self, np_bytes):
        return torch.from_numpy(np.frombuffer(np_bytes, dtype=np.uint8).copy()).to(self.device)


class ShardedCodec:
    """compress(): every rank passes its z-slab; rank 0 gets the complete .ckl stream (uint8 tensor on the backend's
    device), byte-identical to compressing the whole volume at once.  Other ranks get None."""

    def __init__(self, ctx_or_backend, dist, backend=None):
        self.dist = dist
        self.be = backend if backend is not None else CudaShardBackend(ctx_or_backend)
        self.ctx = ctx_or_backend
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    # -- helpers ---------------------------------------------------------------------------------------------
    def _dev(self):
        return self.be.device

    def _all_gather_i64(self, vals):
        """fixed-size all_gather of a few int64 per rank: one collective into one tensor, one device->host copy"""
        t = torch.tensor(vals, dtype=torch.int64, device=self._dev())
        out = torch.empty(self.world * t.numel(), dtype=torch.int64, device=self._dev())
        self.dist.all_gather_into_tensor(out, t)
        a = out.cpu().numpy().view(np.uint64).reshape(self.world, t.numel())
        return [a[r] for r in range(self.world)]

    def _all_gather_var(self, t, counts, to_host=False):
        """all_gather of 1-D tensors of different lengths (padded to the max); to_host: numpy views of ONE host copy."""
        m = max(max(counts), 1)
        pad = torch.zeros(m, dtype=t.dtype, device=self._dev())
        pad[: t.numel()] = t
        out = torch.empty(self.world * m, dtype=t.dtype, device=self._dev())
        self.dist.all_gather_into_tensor(out, pad)
        if to_host:
            out = out.cpu().numpy()
        return [out[r * m: r * m + c] for r, c in enumerate(counts)]

    # -- compress --------------------------------------------------------------------------------------------
    def _mark(self, name):
        """CKL_DIST_PROF=1: wall-clock phase times of compress() (device synchronised), printed by rank 0."""
        if not self._prof:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t = time.perf_counter()
        self._marks.append((name, (t - self._t0) * 1e3))
        self._t0 = t

    def compress(self, vol, z0, sz_total, markov_model_order=0, fortran_order=True):
        be, dist, W, R = self.be, self.dist, self.world, self.rank
        self._prof = os.environ.get("CKL_DIST_PROF") == "1"
        self._marks, self._t0 = [], time.perf_counter()
        s = be.begin(vol)
        self._mark("begin")
        sx, sy, sz_local = be.shape
        data_width = be.width
        # (1) global scalars from the per-shard summaries
        summ = self._all_gather_i64([_i64(s["max_label"]), _i64(s["pairs"]), _i64(s["first_voxel"]), _i64(s["last_voxel"]),
                                     _i64(s["voxels"]), sz_local])
        max_label = max(int(a[0]) for a in summ)
        pairs = sum(int(a[1]) for a in summ)
        for r in range(1, W):                      # the flat-index pair straddling each shard boundary (lib.hpp:249-256)
            pairs += int(summ[r][2] == summ[r - 1][3])
        voxels = sum(int(a[4]) for a in summ)
        sz_all = [int(a[5]) for a in summ]
        assert sum(sz_all) == sz_total, "shards do not cover the volume"
        permissible = pairs < voxels // 2          # crackle.hpp:50-55
        stored = byte_width(max_label)             # crackle.hpp:233-235
        self._mark("summaries")
        # (2) per-shard encode
        nu_local, ncomp_local, ncp_local = be.encode(permissible, stored, markov_model_order)
        self._mark("encode")
        cnt = self._all_gather_i64([nu_local, ncomp_local, ncp_local])
        nu_all = [int(c[0]) for c in cnt]
        ncomp_all = [int(c[1]) for c in cnt]
        order = markov_model_order
        if order > 0 and sum(int(c[2]) for c in cnt) == 0:
            order = 0                              # crackle.hpp:107-118
        # (3) global sorted unique label table: identical merge on every rank
        parts = self._all_gather_var(be.unique(), nu_all)
        guniq = be.sort_unique(torch.cat(parts) if W > 1 else parts[0].clone(), stored)
        nu = int(guniq.numel())
        self._mark("unique_merge")
        # (4) global markov statistics
        gstats = None
        if order > 0:
            gstats = be.stats().clone()
            dist.all_reduce(gstats, op=dist.ReduceOp.SUM)     # int32 two's complement add == uint32 wrap (markov.hpp:210-213)
        # (5) per-shard pieces against the global table / model
        pc = be.finish(guniq, gstats)
        self._mark("finish")
        nz, code_sizes, crcs = be.small_pieces()
        sizes = self._all_gather_i64([pc["keys_bytes"], pc["codes_bytes"]])
        keys_all = [int(a[0]) for a in sizes]
        codes_all = [int(a[1]) for a in sizes]
        small = torch.from_numpy(np.concatenate([nz.view(np.int64), code_sizes.astype(np.int64), crcs.astype(np.int64)])).to(self._dev())
        smalls = self._all_gather_var(small, [3 * z for z in sz_all], to_host=True)
        self._mark("small_gathers")
        # (6) gather keys and codes on rank 0 straight into their place in the stream
        kw, cw = byte_width(nu), byte_width(sx * sy)
        labels_bytes = 8 + nu * stored + sz_total * cw + sum(keys_all)
        off_lab = 29 + 4 * (sz_total + 1)
        off_keys = off_lab + 8 + nu * stored + sz_total * cw
        off_model = off_lab + labels_bytes
        off_codes = off_model + model_bytes(order)
        total = off_codes + sum(codes_all) + 4 + 4 * sz_total
        if R == 0:
            final = be.empty_bytes(total)
            kpos, cpos = off_keys, off_codes
            be.big_pieces(final[kpos:kpos + keys_all[0]], final[cpos:cpos + codes_all[0]])
            kpos += keys_all[0]
            cpos += codes_all[0]
            for r in range(1, W):
                if keys_all[r]:
                    dist.recv(final[kpos:kpos + keys_all[r]], src=r)
                if codes_all[r]:
                    dist.recv(final[cpos:cpos + codes_all[r]], src=r)
                kpos += keys_all[r]
                cpos += codes_all[r]
        else:
            kt, ct = be.empty_bytes(max(keys_all[R], 1)), be.empty_bytes(max(codes_all[R], 1))
            be.big_pieces(kt[: keys_all[R]], ct[: codes_all[R]])
            if keys_all[R]:
                dist.send(kt[: keys_all[R]], dst=0)
            if codes_all[R]:
                dist.send(ct[: codes_all[R]], dst=0)
            return None
        self._mark("keys_codes_gather")
        # (7) rank 0: the small sections (crackle.hpp:171-216, labels.hpp:123-152)
        sm_np = smalls
        nz_g = np.concatenate([p[: z].view(np.uint64) for p, z in zip(sm_np, sz_all)])
        cs_g = np.concatenate([p[z: 2 * z].astype(np.uint32) for p, z in zip(sm_np, sz_all)])
        cr_g = np.concatenate([p[2 * z: 3 * z].astype(np.uint32) for p, z in zip(sm_np, sz_all)])
        zidx = cs_g.astype("<u4").tobytes()
        uniq_np = guniq.cpu().numpy().view(np.uint64)
        head = header_bytes(data_width, stored, int(permissible), fortran_order, order, sx, sy, sz_total, labels_bytes)
        zcrc = be.crc32c(be.to_device(zidx))
        front = (head + zidx + int(zcrc).to_bytes(4, "little") + int(nu).to_bytes(8, "little") +
                 uniq_np.astype(f"<u{stored}").tobytes() + nz_g.astype(f"<u{cw}").tobytes())
        assert len(front) == off_keys
        final[:off_keys] = be.to_device(front)
        if order > 0:
            final[off_model:off_codes] = be.to_device(be.stored_model())
        lcrc = be.crc32c(final[off_lab:off_model])
        tail = int(lcrc).to_bytes(4, "little") + cr_g.astype("<u4").tobytes()
        final[total - len(tail):] = be.to_device(tail)
        self._mark("assemble")
        if self._prof:
            print("CKL_DIST_PROF " + " ".join(f"{k}={v:.2f}" for k, v in self._marks), flush=True)
        return final

    # -- decompress ------------------------------------------------------------------------------------------
    def broadcast_stream(self, stream):
        """rank 0's stream -> every rank (uint8 tensor on the backend's device)."""
        n = torch.tensor([stream.numel() if self.rank == 0 else 0], dtype=torch.int64, device=self._dev())
        self.dist.broadcast(n, src=0)
        if self.rank != 0:
            stream = self.be.empty_bytes(int(n.item()))
        self.dist.broadcast(stream, src=0)
        return stream

    def decompress_shard(self, stream, z_start, z_end, out, label=None):
        """every rank decodes its own z-range of the (replicated) stream into its own output shard; no collective."""
        self.ctx.decompress_into(stream.data_ptr(), 1, stream.numel(), z_start, z_end, label, out.data_ptr(), 1,
                                 out.numel() * out.element_size())
