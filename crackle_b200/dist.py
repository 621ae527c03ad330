"""z-sharded multi-GPU compress / decompress: one process per GPU, torch.distributed for the plumbing.

Slices are independent, so the volume shards along z (rank r owns a contiguous z-slab).  Volume-proportional data
never leaves its GPU; the only exchanges are the small global couplings of the stream
(the reference's own recipe for stitching independently compressed slabs is
crackle/operations.py:258-295 `_zstack_flat_labels` and :424-548 `zstack`):

  * max label -> stored width, pixel pairs -> crack format, counts    (ONE all_gather of an 80-byte record, after the encode:
                                                                       every shard encodes with the crack format its own
                                                                       statistics suggest and re-encodes in the rare case the
                                                                       global decision differs)
  * sorted unique label table                                          (all_gather of per-shard unique sets, merge+unique on device)
  * markov statistics                                                  (all_reduce SUM of uint32[4^order*4], wraps mod 2^32)
  * N_z / code sizes / slice crcs / keys / crack codes                 (ONE padded all_gather of each shard's packed block)

After the last all_gather every rank holds every piece and assembles the complete, header-consistent stream on its own
GPU (ckl_shard_assemble: no host round trip, no rank-0 bottleneck, no broadcast before a sharded decompress).

The per-shard compute is behind a small backend interface so the host logic can be exercised on CPU with gloo
(tests/test_dist_cpu.py supplies an oracle-backed fake); the product backend is `CudaShardBackend` (C-ABI)."""
import ctypes
import os
import sys
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
    """Per-GPU shard stages through the C-ABI (ckl_shard_*).  The context is bound to torch's current CUDA stream so the
    library's kernels, torch ops and the NCCL collectives are ordered by the stream, not by host synchronisation."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.L = _capi.lib()
        self.device = torch.device("cuda", ctx.device)
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        self.ctx._check(rc)

    def begin(self, vol):
        """vol: CUDA tensor (sz_local, sy, sx) of an unsigned dtype, or a pinned / pageable HOST numpy array or tensor of that
        shape (uploaded inside the call)."""
        s = _capi.ShardSummary()
        if isinstance(vol, np.ndarray):
            sz, sy, sx = vol.shape
            ptr, on_dev, width = vol.ctypes.data, 0, vol.dtype.itemsize
        else:
            sz, sy, sx = vol.shape
            ptr, on_dev, width = vol.data_ptr(), int(vol.is_cuda), vol.element_size()
        self.shape = (sx, sy, sz)
        self.width = width
        self._check(self.L.ckl_shard_begin(self.ctx._h, ptr, on_dev, width, sx, sy, sz, ctypes.byref(s)))
        return dict(max_label=s.max_label, pairs=s.pairs, first_voxel=s.first_voxel, last_voxel=s.last_voxel, voxels=s.voxels)

    def encode(self, permissible, stored_width, order):
        nu, nc, ncp = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        self.order = order
        self._check(self.L.ckl_shard_encode(self.ctx._h, int(permissible), int(stored_width), int(order), ctypes.byref(nu),
                                            ctypes.byref(nc), ctypes.byref(ncp)))
        self.n_unique_local, self.ncomp = nu.value, nc.value
        return nu.value, nc.value, ncp.value

    def encode_async(self, permissible, stored_width, order):
        """queues the encode stage; returns when the CCL / label chain is through (the tracer may still be running)."""
        nu, nc = ctypes.c_uint64(), ctypes.c_uint64()
        self.order = order
        self._check(self.L.ckl_shard_encode_async(self.ctx._h, int(permissible), int(stored_width), int(order), ctypes.byref(nu),
                                                  ctypes.byref(nc)))
        self.n_unique_local, self.ncomp = nu.value, nc.value
        return nu.value, nc.value

    def encode_wait(self):
        """joins the tracer; -> (codepoints, order-0 code bytes) of the shard"""
        ncp, cb = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.L.ckl_shard_encode_wait(self.ctx._h, ctypes.byref(ncp), ctypes.byref(cb)))
        return ncp.value, cb.value

    def info(self):
        c = _capi.ShardCounts()
        self._check(self.L.ckl_shard_info(self.ctx._h, ctypes.byref(c)))
        return {n: getattr(c, n) for n, _ in _capi.ShardCounts._fields_}

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

    def pack(self, buf):
        """this shard's block (N_z | code sizes | crcs | keys | codes) into the uint8 CUDA tensor `buf`"""
        n = ctypes.c_uint64()
        self._check(self.L.ckl_shard_pack(self.ctx._h, buf.data_ptr(), buf.numel(), ctypes.byref(n)))
        return n.value

    def assemble(self, gathered, blocks, guniq, data_width, stored, permissible, fortran_order, order, sx, sy):
        """complete stream from the gathered blocks, built on this rank's GPU; returns a zero-copy uint8 view of the
        context-owned result buffer (valid until the context's next compress)."""
        arr = (_capi.ShardBlock * len(blocks))()
        for i, b in enumerate(blocks):
            arr[i].offset, arr[i].sz_local, arr[i].n_components, arr[i].keys_bytes, arr[i].codes_bytes = b
        n = ctypes.c_uint64()
        self._check(self.L.ckl_shard_assemble(self.ctx._h, gathered.data_ptr(), arr, len(blocks), guniq.data_ptr(), 1, guniq.numel(),
                                              int(data_width), int(stored), int(permissible), int(bool(fortran_order)), int(order),
                                              sx, sy, ctypes.byref(n)))
        return self.result_view()

    def result_view(self):
        p, n = self.ctx.result_device()
        return torch.as_tensor(_DevBytes(p, n), device=self.device)

    # pieces fetched separately (kept for callers that place them themselves)
    def small_pieces(self):
        """-> (components_per_slice u64[sz], code_sizes u32[sz], slice_crcs u32[sz]) as numpy"""
        sz = int(self.pieces.sz_local)
        nz = np.zeros(sz, dtype=np.uint64)
        cs = np.zeros(sz, dtype=np.uint32)
        cr = np.zeros(sz, dtype=np.uint32)
        self._check(self.L.ckl_shard_fetch(self.ctx._h, None, nz.ctypes.data, cs.ctypes.data, cr.ctypes.data, None, 0))
        return nz, cs, cr

    def stored_model(self):
        n = ctypes.c_uint64()
        self._check(self.L.ckl_shard_model(self.ctx._h, None, 0, 0, ctypes.byref(n)))
        buf = np.zeros(max(n.value, 1), dtype=np.uint8)
        if n.value:
            self._check(self.L.ckl_shard_model(self.ctx._h, buf.ctypes.data, 0, n.value, ctypes.byref(n)))
        return buf[: n.value].tobytes()

    def empty_bytes(self, n):
        return torch.empty(n, dtype=torch.uint8, device=self.device)


class _DevBytes:
    """zero-copy __cuda_array_interface__ wrapper of a device byte range"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def block_bytes(sz_local, keys_bytes, codes_bytes):
    """size of one shard's packed block (layout of ckl_shard_pack)"""
    return 12 * sz_local + keys_bytes + codes_bytes


class ShardedCodec:
    """compress(): every rank passes its z-slab and EVERY rank gets the complete .ckl stream (uint8 tensor on the
    backend's device), byte-identical to compressing the whole volume at once."""


    def __init__(self, ctx_or_backend, dist, backend=None):
        self.dist = dist
        self.be = backend if backend is not None else CudaShardBackend(ctx_or_backend)
        self.ctx = ctx_or_backend
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.collectives = 0

    # -- helpers ---------------------------------------------------------------------------------------------
    def _dev(self):
        return self.be.device

    def _all_gather_i64(self, vals):
        """fixed-size all_gather of a few int64 per rank: one collective into one tensor, one device->host copy"""
        if self.world == 1:
            return [np.array([int(v) for v in vals], dtype=np.int64).view(np.uint64)]
        t = torch.tensor(vals, dtype=torch.int64, device=self._dev())
        out = torch.empty(self.world * t.numel(), dtype=torch.int64, device=self._dev())
        self.dist.all_gather_into_tensor(out, t)
        self.collectives += 1
        a = out.cpu().numpy().view(np.uint64).reshape(self.world, t.numel())
        return [a[r] for r in range(self.world)]

    def _all_gather_var(self, t, counts):
        """all_gather of 1-D tensors of different lengths (padded to the max) -> list of device views"""
        if self.world == 1:
            return [t]
        m = max(max(counts), 1)
        pad = torch.empty(m, dtype=t.dtype, device=self._dev())
        pad[: t.numel()] = t
        out = torch.empty(self.world * m, dtype=t.dtype, device=self._dev())
        self.dist.all_gather_into_tensor(out, pad)
        self.collectives += 1
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

    def _encode_async(self, permissible, stored_width, order):
        """-> (n_unique_local, n_components); backends without the split stage (test stand-ins) encode in one go"""
        be = self.be
        if hasattr(be, "encode_async"):
            return be.encode_async(permissible, stored_width, order)
        nu, nc, ncp = be.encode(permissible, stored_width, order)
        return nu, nc

    def _encode_wait(self):
        """-> (n_codepoints, codes_bytes_order0) of this shard"""
        be = self.be
        if hasattr(be, "encode_wait"):
            return be.encode_wait()
        info = be.info()
        return info["n_codepoints"], info["codes_bytes_order0"]

    def compress(self, vol, z0, sz_total, markov_model_order=0, fortran_order=True):
        be, dist, W, R = self.be, self.dist, self.world, self.rank
        self._prof = os.environ.get("CKL_DIST_PROF") == "1"
        self._marks, self._t0 = [], time.perf_counter()
        s = be.begin(vol)
        sx, sy, sz_local = be.shape
        data_width = be.width
        # (1) encode with the crack format this shard's own statistics suggest (crackle.hpp:50-55).  The stage is only QUEUED:
        # the call returns when the CCL / label chain is through, and the first metadata exchange and the merge of the unique
        # tables run beside the tracer's serial chain replay; the global decision is checked right after that exchange.
        guess = int(s["pairs"]) < int(s["voxels"]) // 2
        nu_local, ncomp_local = self._encode_async(guess, byte_width(int(s["max_label"])), markov_model_order)
        self._mark("begin+encode_async")
        redone = False
        while True:
            meta = self._all_gather_i64([_i64(s["max_label"]), _i64(s["pairs"]), _i64(s["first_voxel"]), _i64(s["last_voxel"]),
                                         _i64(s["voxels"]), sz_local, nu_local, ncomp_local])
            max_label = max(int(a[0]) for a in meta)
            pairs = sum(int(a[1]) for a in meta)
            for r in range(1, W):                  # the flat-index pair straddling each shard boundary (lib.hpp:249-256)
                pairs += int(meta[r][2] == meta[r - 1][3])
            voxels = sum(int(a[4]) for a in meta)
            sz_all = [int(a[5]) for a in meta]
            assert sum(sz_all) == sz_total, "shards do not cover the volume"
            permissible = pairs < voxels // 2          # crackle.hpp:50-55
            stored = byte_width(max_label)             # crackle.hpp:233-235
            wrong = [r for r in range(W) if (int(meta[r][1]) < int(meta[r][4]) // 2) != permissible]
            if not wrong or redone:
                break
            if R in wrong:                             # rare: this shard's guess differs from the global decision
                self._encode_wait()
                s = be.begin(vol)
                nu_local, ncomp_local = self._encode_async(permissible, stored, markov_model_order)
            redone = True
        nu_all = [int(a[6]) for a in meta]
        ncomp_all = [int(a[7]) for a in meta]
        self._mark("meta")
        # (2) global sorted unique label table: identical merge on every rank
        parts = self._all_gather_var(be.unique(), nu_all)
        guniq = be.sort_unique(torch.cat(parts) if W > 1 else parts[0].clone(), stored)
        nu = int(guniq.numel())
        self._mark("unique_merge")
        # (2b) join the tracer; the code sizes travel in a second small exchange
        ncp_local, codes0_local = self._encode_wait()
        meta2 = self._all_gather_i64([ncp_local, codes0_local])
        order = markov_model_order
        if order > 0 and sum(int(a[0]) for a in meta2) == 0:
            order = 0                              # crackle.hpp:107-118
        self._mark("encode_wait+meta2")
        # (3) global markov statistics
        gstats = None
        if order > 0:
            gstats = be.stats().clone()
            if W > 1:
                dist.all_reduce(gstats, op=dist.ReduceOp.SUM)     # int32 two's complement add == uint32 wrap (markov.hpp:210-213)
                self.collectives += 1
        # (4) per-shard pieces against the global table / model
        pc = be.finish(guniq, gstats)
        kw = byte_width(nu)
        keys_all = [n * kw for n in ncomp_all]
        if order > 0:
            codes_all = [int(a[0]) for a in self._all_gather_i64([pc["codes_bytes"]])]
        else:
            codes_all = [int(a[1]) for a in meta2]
        assert pc["keys_bytes"] == keys_all[R] and pc["codes_bytes"] == codes_all[R]
        self._mark("finish")
        # (5) ONE padded all_gather of the packed blocks; every rank then holds every piece
        sizes = [block_bytes(z, k, c) for z, k, c in zip(sz_all, keys_all, codes_all)]
        P = (max(sizes) + 15) // 16 * 16
        if W > 1:
            gathered = be.empty_bytes(W * P)
            mine = be.empty_bytes(P)
            be.pack(mine)
            dist.all_gather_into_tensor(gathered, mine)
            self.collectives += 1
        else:
            gathered = be.empty_bytes(P)
            be.pack(gathered)
        self._mark("gather")
        # (6) the complete stream, assembled on this rank's device (crackle.hpp:171-216, labels.hpp:123-152)
        blocks = [(r * P, sz_all[r], ncomp_all[r], keys_all[r], codes_all[r]) for r in range(W)]
        final = be.assemble(gathered, blocks, guniq, data_width, stored, int(permissible), fortran_order, order, sx, sy)
        self._mark("assemble")
        if self._prof and R == 0:
            sys.stderr.write("CKL_DIST_PROF " + " ".join(f"{k}={v:.2f}" for k, v in self._marks) + "\n")
        return final

    # -- decompress ------------------------------------------------------------------------------------------
    def broadcast_stream(self, stream):
        """kept for callers that hold the stream on rank 0 only (e.g. read from storage): rank 0's stream -> every rank.
        compress() already leaves the stream on every rank, so a compress -> decompress pipeline does not need it."""
        if self.world == 1:
            return stream
        n = torch.tensor([stream.numel() if self.rank == 0 else 0], dtype=torch.int64, device=self._dev())
        self.dist.broadcast(n, src=0)
        if self.rank != 0:
            stream = self.be.empty_bytes(int(n.item()))
        self.dist.broadcast(stream, src=0)
        return stream

    def decompress_shard(self, stream, z_start, z_end, out, label=None):
        """every rank decodes its own z-range of the (replicated) stream into its own output shard; no collective.
        out: CUDA tensor, or a HOST numpy array / tensor (downloaded inside the call)."""
        if isinstance(out, np.ndarray):
            optr, on_dev, nbytes = out.ctypes.data, 0, out.nbytes
        else:
            optr, on_dev, nbytes = out.data_ptr(), int(out.is_cuda), out.numel() * out.element_size()
        self.ctx.decompress_into(stream.data_ptr(), 1, stream.numel(), z_start, z_end, label, optr, on_dev, nbytes)
