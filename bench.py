#!/usr/bin/env python
"""bench.py -- compress + decompress throughput of the per-z-slice crackle hot path on B200.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c4|c5] [--scaling strong|weak]

Workloads (BASELINE.json `configs`):
  c3 (default)  1024^3 uint64 jittered-Voronoi segmentation, order 0 -- the configuration the metric is quoted on.
                N > 1: ONE 1024^3 volume z-sharded over the N ranks (strong scaling, rank r owns slices
                [r*1024/N, (r+1)*1024/N)); --scaling weak gives every rank its own 1024^3 slab of a 1024x1024x1024N volume.
  c2 / c4       512^3 uint64 (~10k labels), markov order 0 / 5
  c5            2048x2048x1024 uint32 dense labels (cell 16): decompress-only plus decompress(label=...) mask extraction,
                z-sharded over the N ranks
A "step" = one compress of the resident volume followed by one decompress of the resulting stream (c5: one full decode plus
one single-label mask decode); every voxel is processed twice per step; value = 2*V_total / step time in GVox/s.
Inputs are far larger than L2 (8.6 GB at 1024^3 uint64), so no explicit L2 flush is needed between iterations.
One JSON line on stdout (rank 0)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    #        shape               dtype      cell order mode
    "c3": ((1024, 1024, 1024), "uint64", 24, 0, "roundtrip"),
    "c2": ((512, 512, 512), "uint64", 24, 0, "roundtrip"),
    "c4": ((512, 512, 512), "uint64", 24, 5, "roundtrip"),
    "c5": ((2048, 2048, 1024), "uint32", 16, 0, "decode"),
}


def metric_name(shape, dtype, mode):
    sx, sy, sz = shape
    dims = f"{sx}^3" if sx == sy == sz else f"{sx}x{sy}x{sz}"
    if mode == "decode":
        return f"decompress + decompress(label=) GVox/s ({dtype} {dims})"
    return f"compress+decompress GVox/s ({dtype} {dims})"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = one volume of the workload's shape z-sharded over the ranks (BASELINE configs[2]); "
                         "weak = every rank owns a slab of the workload's shape")
    ap.add_argument("--shape", default=None, help="override the workload's sx,sy,sz")
    ap.add_argument("--cell", type=int, default=None)
    ap.add_argument("--order", type=int, default=None, help="markov_model_order")
    ap.add_argument("--cpu-slices", type=int, default=128, help="z-slab size of the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pipeline", default="auto", choices=["auto", "on", "off"],
                    help="e2e: double-buffer the steps (compress of step i+1 beside decompress of step i); auto = on at N = 1")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numpy-e2e", action="store_true", help="skip the pageable-numpy end-to-end measurement (N = 1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the byte-parity checks of the warm-up (round trip is always checked)")
    ap.add_argument("--prof", action="store_true", help="print per-stage timings to stderr")
    ap.add_argument("--chunks", type=int, default=0, help="z-chunk pipelining: 0 = library default (host-resident volumes only), 1 = off, K = force")
    a = ap.parse_args()
    shape, dtype, cell, order, mode = WORKLOADS[a.workload]
    if a.shape:
        shape = tuple(int(v) for v in a.shape.split(","))
    a.shape, a.dtype, a.mode = shape, dtype, mode
    a.cell = cell if a.cell is None else a.cell
    a.order = order if a.order is None else a.order
    return a


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Samples inside the timed window [t0, t1]; when the window is shorter than the sampling period the samples
        taken since the sampler started (warm-up + timed steps, the same kernels back to back) are used and `window` says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        window = "timed"
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1 + 0.12)]
        if len(rows) < 2:
            rows, window = [r for (_, r) in self.rows], "warmup+timed"
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def ref_times(vol, order, reps, mode="roundtrip", label=None):
    """Reference CPU path (oracle/_ref, all host cores) on `vol`; returns (t_first, t_second, kind, cores): compress and
    decompress times, or (mode "decode") full-decode and label-mask-decode times."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    ref = O.ref_module()
    ta = tb = 1e30
    if ref is not None:
        kind = "reference"
        b = None
        for _ in range(reps):
            if mode == "decode":
                if b is None:
                    b = ref.compress(vol, False, True, order, False, True, 0, 0)
                t0 = time.perf_counter(); ref.decompress(b, 0, -1, 0, None); t1 = time.perf_counter()
                ref.decompress(b, 0, -1, 0, label); t2 = time.perf_counter()
            else:
                t0 = time.perf_counter(); b = ref.compress(vol, False, True, order, False, True, 0, 0); t1 = time.perf_counter()
                ref.decompress(b, 0, -1, 0, None); t2 = time.perf_counter()
            ta, tb = min(ta, t1 - t0), min(tb, t2 - t1)
    else:
        kind, cores = "port", 1
        b = None
        for _ in range(reps):
            if mode == "decode":
                if b is None:
                    b = O.compress(vol, order)
                t0 = time.perf_counter(); O.decompress(b); t1 = time.perf_counter()
                O.decompress(b, label=label); t2 = time.perf_counter()
            else:
                t0 = time.perf_counter(); b = O.compress(vol, order); t1 = time.perf_counter()
                O.decompress(b); t2 = time.perf_counter()
            ta, tb = min(ta, t1 - t0), min(tb, t2 - t1)
    return ta, tb, kind, cores


def np_dtype(name):
    return np.dtype(name)


def host_sample(args, zs):
    """first `zs` slices of the workload as an F-ordered numpy array (GPU synth when a device is there: same integers)."""
    from crackle_b200 import synth
    sx, sy, sz = args.shape
    bits = 40 if args.dtype == "uint64" else 30
    try:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError
        t = synth.jittered_voronoi_torch((sx, sy, zs), args.cell, np_dtype(args.dtype), seed=0, id_bits=bits, device="cuda", sz_total=sz)
        vol = np.asfortranarray(t.cpu().numpy().transpose(2, 1, 0))
        del t
    except Exception:
        vol = synth.jittered_voronoi((sx, sy, zs), args.cell, np_dtype(args.dtype), seed=0, id_bits=bits, sz_total=sz)
    return vol


def workload_text(args, world, scaling):
    sx, sy, sz = args.shape
    what = ("decompress-only plus decompress(label=) mask extraction" if args.mode == "decode" else "compress then decompress")
    shard = ""
    if world > 1:
        shard = (f"; ONE volume z-sharded over {world} GPUs ({sz // world} slices per rank)" if scaling == "strong"
                 else f"; every rank owns a {sx}x{sy}x{sz} slab of a {sx}x{sy}x{sz * world} volume")
    return (f"{sx}x{sy}x{sz} {args.dtype} jittered-Voronoi segmentation (cell {args.cell}), flat labels, markov order {args.order}; "
            f"{what}, device-resident{shard}")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on a bounded z-slab of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sx, sy, sz = args.shape
    zs = min(args.cpu_slices if args.dtype == "uint64" and sx <= 1024 else max(8, args.cpu_slices // 4), sz)
    vol = host_sample(args, zs)
    V = vol.size
    label = int(vol[sx // 2, sy // 2, zs // 2])
    times = []
    for i in range(args.warmup + args.steps):
        ta, tb, kind, cores = ref_times(vol, args.order, 1, args.mode, label)
        if i >= args.warmup:
            times.append(ta + tb)
    t = float(np.mean(times))
    val = 2 * V / t / 1e9
    sample = f"{sx}x{sy}x{zs} z-slab of the {sx}x{sy}x{sz} workload"
    line = {"metric": metric_name(args.shape, args.dtype, args.mode), "value": val, "unit": "GVox/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "u64" if args.dtype == "uint64" else "u32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_text(args, 1, args.scaling).replace(", device-resident", ""), "sample": sample},
            "cpu_baseline": {"value": val, "unit": "GVox/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "GVox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import crackle_b200 as cb
    from crackle_b200 import synth
    from crackle_b200 import dist as cdist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scaling = args.scaling if world > 1 else "weak"
    sx, sy, sz = args.shape
    dt = np_dtype(args.dtype)
    width = dt.itemsize
    tdt = getattr(torch, args.dtype)
    bits = 40 if args.dtype == "uint64" else 30
    if scaling == "strong":
        sz_total = sz
        z0, z1 = rank * sz // world, (rank + 1) * sz // world
    else:
        sz_total = sz * world
        z0, z1 = rank * sz, (rank + 1) * sz
    szl = z1 - z0
    Vl = sx * sy * szl                 # voxels of this rank's slab
    V = sx * sy * sz_total             # voxels of the whole job
    ctx = cb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_chunks(args.chunks)
    job = cdist.ShardedCodec(ctx, dist) if world > 1 else None

    vol = synth.jittered_voronoi_torch((sx, sy, szl), args.cell, dt, seed=0, id_bits=bits, device="cuda", z0=z0, sz_total=sz_total)
    out = torch.empty_like(vol)
    torch.cuda.synchronize()
    label = int(vol[szl // 2, sy // 2, sx // 2].item())
    mask = torch.empty(vol.shape, dtype=torch.uint8, device="cuda") if args.mode == "decode" else None

    def compress_step():
        if world > 1:
            return job.compress(vol, z0=z0, sz_total=sz_total, markov_model_order=args.order)
        ctx.compress_ptr(vol.data_ptr(), 1, width, sx, sy, szl, True, args.order)
        p, n = ctx.result_device()
        return torch.as_tensor(cdist._DevBytes(p, n), device="cuda")

    def decompress_step(s, lab=None, dst=None):
        dst = out if dst is None else dst
        ctx.decompress_into(s.data_ptr(), 1, s.numel(), z0, z1, lab, dst.data_ptr(), 1, dst.numel() * dst.element_size())

    state = {}
    if args.mode == "decode":
        state["s"] = compress_step().clone()          # untimed: the stream the decode-only workload starts from

        def step():
            decompress_step(state["s"])
            decompress_step(state["s"], label, mask)
            return state["s"]
    else:
        def step():
            s = compress_step()
            decompress_step(s)
            return s

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.5)                           # let nvidia-smi come up before the GPU is loaded
    s = None
    for _ in range(max(1, args.warmup)):          # at least one untimed pass: it is also the round-trip check
        s = step()
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.uint8), vol.view(torch.uint8)), "round trip mismatch"
    if mask is not None:
        assert torch.equal(mask.view(torch.bool), vol.view(torch.int64 if width == 8 else torch.int32) == label), "label mask mismatch"
    ckl_bytes = int(s.numel())
    parity = {"round_trip": "voxel-exact on every rank"}

    # byte parity inside the same run: (a) a 64-slice slab against the reference compiled on this box
    # (oracle/_ref, the checker); (b) N > 1, strong scaling: the sharded stream == a single-GPU compress of the whole volume
    if not args.no_parity:
        from oracle import oracle as O
        ref = O.ref_module()
        zs = min(64 if sx * sy <= 1024 * 1024 else 16, szl)
        if rank == 0:
            chk = cb.Context(local)
            hv = np.asfortranarray(vol[:zs].cpu().numpy().transpose(2, 1, 0))
            got = chk.compress(vol[:zs].contiguous(), args.order)
            want = ref.compress(hv, False, True, args.order, False, True, 0, 0) if ref is not None else O.compress(hv, args.order)
            assert got == bytes(want), "byte parity against the reference failed on the warm-up slab"
            parity["slab"] = f"{sx}x{sy}x{zs} slab: {len(got)} bytes == {'oracle/_ref (compiled reference)' if ref is not None else 'oracle port'}"
            if world > 1 and scaling == "strong" and args.mode == "roundtrip":
                whole = synth.jittered_voronoi_torch((sx, sy, sz_total), args.cell, dt, seed=0, id_bits=bits, device="cuda")
                torch.cuda.synchronize()
                chk.compress_ptr(whole.data_ptr(), 1, width, sx, sy, sz_total, True, args.order)
                p, n = chk.result_device()
                mono = torch.as_tensor(cdist._DevBytes(p, n), device="cuda")
                torch.cuda.synchronize()
                assert n == s.numel() and torch.equal(mono, s), "sharded stream differs from the single-GPU stream"
                parity["sharded"] = f"{world}-rank stream ({n} bytes) == single-GPU compress of the whole volume (rank 0)"
                del whole, mono
            chk.close()
            del chk
            torch.cuda.empty_cache()
        if dist:
            dist.barrier()

    ctx.prof_enable(True)
    launches0 = cb.codec.launch_count()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    tw1 = time.time()
    if dist:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = cb.codec.launch_count() - launches0
    clk = clocks.stop(tw0, tw1) if rank == 0 else None
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    if dist:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    value = 2.0 * V / (ms_step * 1e-3) / 1e9

    drains = []                       # host-side stream drains per call (the library's ckl_sync_count), in the order timed() is called

    def timed(fn):
        """per-call time of `fn` over args.steps calls: CUDA events on the stream, barrier both sides, max over ranks"""
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0 = cb.codec.sync_count()
        a.record(stream)
        for _ in range(args.steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        drains.append((cb.codec.sync_count() - s0) / args.steps)
        tt = a.elapsed_time(b) / args.steps
        if dist:
            x = torch.tensor([tt], device="cuda", dtype=torch.float64)
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
            tt = float(x.item())
        return tt

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured" if "hbm_gbs" in peaks else "fallback"
    stage_ms = {k: v[0] / max(1, v[1]) for k, v in prof.items()}

    # Algorithmic bytes: compress reads the volume once and writes the stream; decompress reads the stream and
    # writes the volume once (1 byte per voxel for a label mask).  All ranks together; the roofline peak is N x one GPU's.
    if args.mode == "decode":
        alg_a, alg_b = ckl_bytes * world + V * width, ckl_bytes * world + V
        names = ("decompress", "decompress_label")
        t_a = timed(lambda: decompress_step(state["s"]))
        t_b = timed(lambda: decompress_step(state["s"], label, mask))
    else:
        alg_a, alg_b = V * width + ckl_bytes * world, ckl_bytes * world + V * width
        names = ("compress", "decompress")
        t_a = timed(compress_step)
        s = compress_step()
        t_b = timed(lambda: decompress_step(s))
    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")))
    except Exception:
        pass
    tr = traffic_tab if (args.shape == (1024, 1024, 1024) and world == 1 and args.mode == "roundtrip" and args.order == 0) else {}
    alg = alg_a + alg_b
    ach = alg / (ms_step * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "whole step: every kernel of one %s + one %s call" % names, "achieved": ach, "peak": peak * world,
            "peak_kind": peak_kind + (f" x {world} GPUs" if world > 1 else ""), "unit": "GB/s", "frac": ach / (peak * world),
            "traffic": tr.get("step", {}).get("dram_bytes"), "algorithmic_bytes_per_step": alg,
            names[0]: {"ms": t_a, "algorithmic_bytes": alg_a, "achieved": alg_a / (t_a * 1e-3) / 1e9, "frac": alg_a / (t_a * 1e-3) / 1e9 / (peak * world)},
            names[1]: {"ms": t_b, "algorithmic_bytes": alg_b, "achieved": alg_b / (t_b * 1e-3) / 1e9, "frac": alg_b / (t_b * 1e-3) / 1e9 / (peak * world)},
            "note": "whole-call fractions: algorithmic bytes of the call / its CUDA-event time / measured HBM peak.  The kernels between the two "
                    "full-width streaming kernels work on 2 bits per voxel and are issue- or latency-bound, not HBM-bound; roofline_kernels "
                    "lists the single-kernel stages with their OWN algorithmic bytes (rank 0, live inside the timed step)."}
    # single-kernel stages: own algorithmic bytes of the kernel on rank 0's slab
    own = {"edges": ("k_edges_tma (cp.async.bulk staged)" if sx % 256 == 0 and width >= 4 else "k_edges", Vl * width + Vl // 4, "hbm"),
           "d_paint": ("k_paint_tma (cp.async.bulk stores)" if width == 4 and sx % 256 == 0 else "k_paint_band" if sx % 256 == 0 else "k_paint_rows",
                       Vl * width + Vl // 8, "hbm"),
           "trace_replay": ("k_replay", None, "shared-memory latency (serial walk per slice); not an HBM kernel")}
    kern = {}
    for k, (kname, b, bound) in own.items():
        if k in stage_ms:
            e = {"kernel": kname, "kernel_ms": stage_ms[k], "bound": bound, "traffic": tr.get(k, {}).get("dram_bytes")}
            if b:
                e.update({"algorithmic_bytes": b, "achieved": b / (stage_ms[k] * 1e-3) / 1e9, "frac": b / (stage_ms[k] * 1e-3) / 1e9 / peak})
            kern[k] = e

    line = {"metric": metric_name(args.shape, args.dtype, args.mode), "value": value, "unit": "GVox/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "u64" if width == 8 else "u32", "data": "synthetic", "impl": "b200",
            "config": {"workload": workload_text(args, world, scaling), "name": args.workload,
                       "l2": "inputs (volume) are far larger than the 126 MB L2; no flush needed",
                       "voxels_per_step": 2 * V, "ckl_bytes": ckl_bytes, "slices_per_rank": szl, "parity": parity},
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_kernels": kern,
            f"{names[0]}_gvox_s": V / (t_a * 1e-3) / 1e9, f"{names[0]}_ms": t_a, f"{names[0]}_hbm_frac": roof[names[0]]["frac"],
            f"{names[1]}_gvox_s": V / (t_b * 1e-3) / 1e9, f"{names[1]}_ms": t_b, f"{names[1]}_hbm_frac": roof[names[1]]["frac"]}
    line["host_drains_per_call"] = {names[0]: drains[0], names[1]: drains[1],
                                    "what": "cudaStreamSynchronize calls the library issues inside one call (rank 0; ckl_sync_count)"}
    if job is not None:
        line["collectives_per_compress"] = "metadata all_gather + unique-table all_gather (both beside the tracer's chain replay) + code-size all_gather + ONE padded all_gather of the packed blocks" + \
                                           (" + statistics all_reduce + code-size all_gather (order > 0)" if args.order > 0 else "")

    # end to end through the public host API: pinned HOST buffers, H2D + D2H inside the timed region, every rank through its
    # own PCIe link; wall clock between barriers, max over ranks
    if not args.no_e2e:
        hvol = torch.empty(vol.shape, dtype=vol.dtype, pin_memory=True)
        hvol.copy_(vol)
        hout = torch.empty(vol.shape, dtype=vol.dtype, pin_memory=True)
        hmask = torch.empty(vol.shape, dtype=torch.uint8, pin_memory=True) if args.mode == "decode" else None
        hstream = torch.empty(ckl_bytes + 64, dtype=torch.uint8, pin_memory=True)
        if args.mode == "decode":
            hstream[:ckl_bytes].copy_(state["s"])
        torch.cuda.synchronize()

        def e2e_step():
            if args.mode == "decode":
                ctx.decompress_into(hstream.data_ptr(), 0, ckl_bytes, z0, z1, None, hout.data_ptr(), 0, hout.numel() * width)
                ctx.decompress_into(hstream.data_ptr(), 0, ckl_bytes, z0, z1, label, hmask.data_ptr(), 0, hmask.numel())
                return ckl_bytes
            if world > 1:
                st = job.compress(hvol, z0=z0, sz_total=sz_total, markov_model_order=args.order)
                n = st.numel()
                if rank == 0:                      # the stream is the product of compress: rank 0 delivers it to the host
                    ctx.result_to(hstream.data_ptr(), 0, hstream.numel())
                job.decompress_shard(st, z0, z1, hout)
                return n
            n = ctx.compress_ptr(hvol.data_ptr(), 0, width, sx, sy, szl, True, args.order)
            ctx.result_to(hstream.data_ptr(), 0, hstream.numel())
            ctx.decompress_into(hstream.data_ptr(), 0, n, 0, -1, None, hout.data_ptr(), 0, hout.numel() * width)
            return n

        def timed_wall(fn, reps):
            """wall clock of `reps` calls of fn between barriers (device drained both sides), max over ranks, per call"""
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                r = fn()
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0) / reps
            if dist:
                x = torch.tensor([te], device="cuda", dtype=torch.float64)
                dist.all_reduce(x, op=dist.ReduceOp.MAX)
                te = float(x.item())
            return te, r

        e2e_step()
        ne = max(1, min(args.steps, 3))
        te, n = timed_wall(e2e_step, ne)
        assert torch.equal(hout.view(torch.uint8), hvol.view(torch.uint8))
        if args.mode == "decode":
            h2d, d2h = 2 * ckl_bytes * world, V * width + V
        else:
            h2d, d2h = V * width + (n if world == 1 else 0), n + V * width
        api = ("ckl_decompress with pinned HOST stream and output buffers" if args.mode == "decode" else
               "ckl_compress / ckl_decompress with pinned HOST buffers (what fastcrackle.compress/decompress bind)" if world == 1 else
               "ShardedCodec.compress / decompress_shard with each rank's slab in pinned HOST memory (ckl_shard_* + ckl_decompress)")
        chunks = ("library default: 4 z-chunks on child contexts for host-resident volumes, so H2D / D2H copies of one chunk "
                  "overlap the kernels of the others" if args.chunks == 0 else args.chunks) if world == 1 else "one slab per rank"
        serial = {"value": 2.0 * V / te / 1e9, "unit": "GVox/s", "ms_per_step": te * 1e3, "steps": ne,
                  "schedule": "one call after the other on one host thread: PCIe carries one direction at a time"}
        line["e2e"] = {"value": serial["value"], "unit": "GVox/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": te * 1e3, "api": api, "chunks": chunks, "schedule": serial["schedule"]}

        # The same calls double-buffered over the steps, as a service that streams volumes through the codec runs them: the
        # compress call of step i+1 (H2D-bound) beside the decompress call of step i (D2H-bound) on a second context and a
        # second host thread, so both directions of the PCIe link are busy.  Every step still uploads its volume, downloads its
        # stream and its decoded volume; K steps are timed from the first upload to the last download (fill and drain included).
        pipelined = args.e2e_pipeline == "on" or (args.e2e_pipeline == "auto" and world == 1)
        if pipelined and args.mode == "roundtrip":
            ctx2 = cb.Context(local)                  # the decompress side: own stream, own workspaces
            ctx2.set_chunks(args.chunks)
            hs = [hstream, torch.empty(hstream.numel(), dtype=torch.uint8, pin_memory=True)]
            ds = [torch.empty(ckl_bytes + 64, dtype=torch.uint8, device="cuda") for _ in range(2)] if world > 1 else None
            hout.zero_()
            fails = []

            def comp(i):
                if world > 1:
                    st = job.compress(hvol, z0=z0, sz_total=sz_total, markov_model_order=args.order)
                    m = int(st.numel())
                    ds[i & 1][:m].copy_(st)           # the context's result buffer is reused by the next compress
                    if rank == 0:
                        ctx.result_to(hs[i & 1].data_ptr(), 0, hs[i & 1].numel())
                    torch.cuda.current_stream().synchronize()
                    return m
                m = ctx.compress_ptr(hvol.data_ptr(), 0, width, sx, sy, szl, True, args.order)
                ctx.result_to(hs[i & 1].data_ptr(), 0, hs[i & 1].numel())
                return m

            def decomp(i, m):
                try:
                    if world > 1:
                        ctx2.decompress_into(ds[i & 1].data_ptr(), 1, m, z0, z1, None, hout.data_ptr(), 0, hout.numel() * width)
                    else:
                        ctx2.decompress_into(hs[i & 1].data_ptr(), 0, m, 0, -1, None, hout.data_ptr(), 0, hout.numel() * width)
                except Exception as e:                # surfaced after the join
                    fails.append(e)

            def pipeline(K):
                m = comp(0)
                for i in range(1, K):
                    th = threading.Thread(target=decomp, args=(i - 1, m))
                    th.start()
                    m2 = comp(i)
                    th.join()
                    m = m2
                decomp(K - 1, m)
                if fails:
                    raise fails[0]
                return m

            pipeline(2)                               # untimed: workspaces of the second context
            K = max(2, args.steps)
            tp, _ = timed_wall(lambda: pipeline(K), 1)
            tp /= K
            assert torch.equal(hout.view(torch.uint8), hvol.view(torch.uint8)), "pipelined e2e: round trip mismatch"
            line["e2e"].update({"value": 2.0 * V / tp / 1e9, "ms_per_step": tp * 1e3, "steps": K,
                                "schedule": f"double-buffered over {K} steps: the compress call of step i+1 runs beside the decompress call of "
                                            "step i (second context, second host thread), both PCIe directions busy; every step uploads its "
                                            "volume and downloads its stream and its decoded volume; fill and drain inside the timed region",
                                "serial": serial})
            ctx2.close()
            del ctx2, hs, ds
        # What a caller of the reference's Python interface sees (crackle.compress(ndarray) -> bytes, decompress(bytes) ->
        # ndarray): PAGEABLE numpy memory, the stream returned as a bytes object, a fresh output array per call.
        if world == 1 and args.mode == "roundtrip" and not args.no_numpy_e2e:
            try:
                arr = np.asfortranarray(np.array(hvol.numpy().transpose(2, 1, 0)))      # pageable copy, F order
                del hout
                hout = None
                b = ctx.compress(arr, args.order)
                t0 = time.perf_counter()
                b = ctx.compress(arr, args.order)
                t1 = time.perf_counter()
                back = ctx.decompress(b)
                t2 = time.perf_counter()
                assert len(b) == ckl_bytes and np.array_equal(back.reshape(arr.shape, order="F"), arr)
                line["e2e"]["numpy_api"] = {"value": 2.0 * V / (t2 - t0) / 1e9, "unit": "GVox/s", "compress_ms": (t1 - t0) * 1e3,
                                            "decompress_ms": (t2 - t1) * 1e3,
                                            "api": "Context.compress(ndarray) -> bytes / Context.decompress(bytes) -> ndarray (the calls behind "
                                                   "crackle_b200.compress / decompress): pageable host memory, one step"}
                del arr, back, b
            except MemoryError as e:
                line["e2e"]["numpy_api"] = {"skipped": f"host memory: {e}"}
        del hvol, hout, hstream

    if not args.no_cpu and rank == 0:
        zs = min(args.cpu_slices if sx * sy <= 1024 * 1024 else max(8, args.cpu_slices // 4), szl)
        sample = np.asfortranarray(vol[:zs].cpu().numpy().transpose(2, 1, 0))
        ta, tb, kind, cores = ref_times(sample, args.order, 2, args.mode, label)
        line["cpu_baseline"] = {"value": 2.0 * sample.size / (ta + tb) / 1e9, "unit": "GVox/s", "cores": cores, "kind": kind,
                                "sample": f"first {zs} z-slices ({sx}x{sy}x{zs}) of rank 0's slab, all host threads",
                                f"{names[0]}_gvox_s": sample.size / ta / 1e9, f"{names[1]}_gvox_s": sample.size / tb / 1e9}
    if rank == 0:
        if args.prof:
            sys.stderr.write(json.dumps(stage_ms, indent=1) + "\n")
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
