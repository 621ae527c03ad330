#!/usr/bin/env python
"""bench.py -- compress + decompress throughput of the per-z-slice crackle hot path on B200.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]
A "step" = one compress of the resident volume followed by one decompress of the resulting stream (each voxel is
processed twice per step); value = 2*V*N / step time in GVox/s.  Inputs (8.6 GB per GPU at 1024^3 uint64) are larger
than L2, so no explicit L2 flush is needed between iterations.  One JSON line on stdout (rank 0)."""
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

METRIC = "compress+decompress GVox/s (uint64 1024^3)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="1024,1024,1024", help="per-GPU slab sx,sy,sz")
    ap.add_argument("--cell", type=int, default=24)
    ap.add_argument("--order", type=int, default=0, help="markov_model_order")
    ap.add_argument("--cpu-slices", type=int, default=128, help="z-slab size of the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--prof", action="store_true", help="print per-stage timings to stderr")
    ap.add_argument("--chunks", type=int, default=0, help="z-chunk pipelining: 0 = library default (host-resident volumes only), 1 = off, K = force")
    return ap.parse_args()


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
        self.window = window
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


def ref_times(vol, order, reps):
    """Reference CPU path (oracle/_ref, all host cores) on `vol`; returns (t_compress, t_decompress, kind, cores)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    ref = O.ref_module()
    tc = td = 1e30
    if ref is not None:
        kind = "reference"
        for _ in range(reps):
            t0 = time.perf_counter(); b = ref.compress(vol, False, True, order, False, True, 0, 0); t1 = time.perf_counter()
            ref.decompress(b, 0, -1, 0, None); t2 = time.perf_counter()
            tc, td = min(tc, t1 - t0), min(td, t2 - t1)
    else:
        kind, cores = "port", 1
        for _ in range(reps):
            t0 = time.perf_counter(); b = O.compress(vol, order); t1 = time.perf_counter()
            O.decompress(b); t2 = time.perf_counter()
            tc, td = min(tc, t1 - t0), min(td, t2 - t1)
    return tc, td, kind, cores


def run_reference(args, shape):
    """--impl reference: the reference's own CPU implementation on a bounded z-slab of the same workload."""
    from crackle_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sx, sy, sz = shape
    zs = min(args.cpu_slices, sz)
    try:
        import torch
        if torch.cuda.is_available():
            t = synth.jittered_voronoi_torch((sx, sy, zs), args.cell, np.uint64, seed=0, id_bits=40, device="cuda", sz_total=sz)
            vol = np.asfortranarray(t.cpu().numpy().transpose(2, 1, 0))
            del t
        else:
            raise RuntimeError
    except Exception:
        vol = synth.jittered_voronoi((sx, sy, zs), args.cell, np.uint64, seed=0, id_bits=40, sz_total=sz)
    V = vol.size
    times = []
    for i in range(args.warmup + args.steps):
        tc, td, kind, cores = ref_times(vol, args.order, 1)
        if i >= args.warmup:
            times.append(tc + td)
    t = float(np.mean(times))
    val = 2 * V / t / 1e9
    sample = f"{sx}x{sy}x{zs} z-slab of the {sx}x{sy}x{sz} workload"
    line = {"metric": METRIC, "value": val, "unit": "GVox/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{sx}x{sy}x{sz} uint64 jittered-Voronoi segmentation (cell {args.cell}), flat labels, "
                                   f"markov order {args.order}; compress then decompress", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "GVox/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "GVox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    shape = tuple(int(v) for v in args.shape.split(","))
    if args.impl == "reference":
        return run_reference(args, shape)

    import torch
    import crackle_b200 as cb
    from crackle_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sx, sy, sz = shape
    V = sx * sy * sz
    ctx = cb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_chunks(args.chunks)

    # each rank owns a z-slab of a (sx, sy, sz*world) volume: weak scaling, per-GPU work fixed
    vol = synth.jittered_voronoi_torch(shape, args.cell, np.uint64, seed=0, id_bits=40, device="cuda", z0=rank * sz,
                                       sz_total=sz * world)
    out = torch.empty_like(vol)
    torch.cuda.synchronize()

    if world > 1:
        from crackle_b200 import dist as cdist
        job = cdist.ShardedCodec(ctx, dist)

        def step():
            stream_t = job.compress(vol, z0=rank * sz, sz_total=sz * world, markov_model_order=args.order)
            stream_t = job.broadcast_stream(stream_t)
            job.decompress_shard(stream_t, rank * sz, (rank + 1) * sz, out)
            return stream_t
    else:
        def step():
            n = ctx.compress_ptr(vol.data_ptr(), 1, 8, sx, sy, sz, True, args.order)
            p, n = ctx.result_device()
            ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)
            return n

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.5)                           # let nvidia-smi come up before the GPU is loaded
    for _ in range(max(1, args.warmup)):          # at least one untimed pass: it is also the round-trip check
        step()
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), vol.view(torch.int64)), "round trip mismatch"
    ckl_bytes = ctx.result_device()[1] if world == 1 else None

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
    value = 2.0 * V * world / (ms_step * 1e-3) / 1e9

    # per-stage (CUDA events inside the library, same stream) -> roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured" if "hbm_gbs" in peaks else "fallback"
    stage_ms = {k: v[0] / max(1, v[1]) for k, v in prof.items()}
    # roofline: the dominant stage (largest CUDA-event time inside the library, measured on the stream it runs on) against
    # the algorithmic bytes of one launch (compress reads the volume and writes the stream, decompress the
    # reverse); `traffic` = dram bytes of that stage's top kernel from the committed ncu capture (profiles/), per launch
    # stages that are exactly ONE kernel launch (the others bundle several kernels and, on the low-priority stream, include
    # time spent waiting for SMs): the dominant kernel is picked among these
    one_kernel = {"edges": "k_edges<u64,16>", "trace_replay": "k_replay<0>", "d_paint": "k_paint_rows<u64,false>"}
    cand = {k: v for k, v in stage_ms.items() if k in one_kernel}
    dom = max(cand, key=cand.get) if cand else None
    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r1_kernel_traffic.json")))
    except Exception:
        pass
    roof = None
    streaming = {}
    if dom and ckl_bytes:
        alg = V * 8 + ckl_bytes
        ach = alg / (stage_ms[dom] * 1e-3) / 1e9
        tr = traffic_tab.get(dom, {}) if sx * sy * sz == 1024 ** 3 else {}
        roof = {"bound": "hbm", "kernel": one_kernel[dom], "stage": dom, "side": "decompress" if dom.startswith("d_") else "compress",
                "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                "traffic": tr.get("dram_bytes"), "algorithmic_bytes_per_launch": alg, "kernel_ms": stage_ms[dom],
                "note": ("the single-kernel stage with the largest live CUDA-event time in the timed region (events on the stream the kernel "
                         "is launched on; other kernels run beside it on the second stream).  `achieved` = algorithmic bytes of the whole "
                         "compress (or decompress) call / that kernel's time.  k_replay is the serial crack-graph walk, one warp per slice, "
                         "bound by dependent shared-memory latency and not by HBM; the full-width streaming kernels are under roofline_streaming")}
        for k in ("edges", "d_paint"):
            if k in stage_ms:
                a2 = alg / (stage_ms[k] * 1e-3) / 1e9
                streaming[k] = {"achieved": a2, "frac": a2 / peak, "kernel_ms": stage_ms[k],
                                "traffic": (traffic_tab.get(k, {}) if sx * sy * sz == 1024 ** 3 else {}).get("dram_bytes")}

    line = {"metric": METRIC, "value": value, "unit": "GVox/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "impl": "b200",
            "config": {"workload": f"{sx}x{sy}x{sz} uint64 jittered-Voronoi segmentation per GPU (cell {args.cell}, 40-bit ids), "
                                   f"flat labels, markov order {args.order}; compress then decompress, device-resident",
                       "l2": "inputs (8 B/voxel volume) are far larger than the 126 MB L2; no flush needed",
                       "voxels_per_step": 2 * V * world, "ckl_bytes": ckl_bytes},
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "gpu_launches": int(launches), "clocks": clk}
    if roof:
        line["roofline"] = roof
        line["roofline_streaming"] = streaming

    # separate compress / decompress throughput (device-resident), for the record
    if world == 1:
        for name, fn in (("compress", lambda: ctx.compress_ptr(vol.data_ptr(), 1, 8, sx, sy, sz, True, args.order)),
                         ("decompress", lambda: ctx.decompress_into(*ctx.result_device()[:1], 1, ctx.result_device()[1], 0, -1, None,
                                                                    out.data_ptr(), 1, out.numel() * 8))):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(args.steps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            t = a.elapsed_time(b) / args.steps
            line[f"{name}_gvox_s"] = V / (t * 1e-3) / 1e9
            line[f"{name}_ms"] = t
            line[f"{name}_hbm_frac"] = (V * 8 + ckl_bytes) / (t * 1e-3) / 1e9 / peak

    # end to end through the public host API: pinned host buffers, H2D + D2H inside the timed region
    if not args.no_e2e and world == 1:
        hvol = torch.empty(vol.shape, dtype=vol.dtype, pin_memory=True)
        hvol.copy_(vol)
        hout = torch.empty(vol.shape, dtype=vol.dtype, pin_memory=True)
        hstream = torch.empty(ckl_bytes + 64, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            n = ctx.compress_ptr(hvol.data_ptr(), 0, 8, sx, sy, sz, True, args.order)
            ctx.result_to(hstream.data_ptr(), 0, hstream.numel())
            ctx.decompress_into(hstream.data_ptr(), 0, n, 0, -1, None, hout.data_ptr(), 0, hout.numel() * 8)
            return n
        e2e_step()
        t0 = time.perf_counter()
        ne = max(1, min(args.steps, 3))
        for _ in range(ne):
            n = e2e_step()
        torch.cuda.synchronize()
        te = (time.perf_counter() - t0) / ne
        assert torch.equal(hout.view(torch.int64), hvol.view(torch.int64))
        line["e2e"] = {"value": 2.0 * V / te / 1e9, "unit": "GVox/s", "h2d_bytes_per_step": int(V * 8 + n),
                       "d2h_bytes_per_step": int(n + V * 8), "ms_per_step": te * 1e3,
                       "api": "ckl_compress / ckl_decompress with pinned HOST buffers (what fastcrackle.compress/decompress bind)",
                       "chunks": "library default: 4 z-chunks on child contexts for host-resident volumes, so H2D / D2H copies of one chunk "
                                 "overlap the kernels of the others" if args.chunks == 0 else args.chunks}
        del hvol, hout, hstream

    if not args.no_cpu and world == 1 and rank == 0:
        zs = min(args.cpu_slices, sz)
        sample = np.asfortranarray(vol[:zs].cpu().numpy().transpose(2, 1, 0))
        tc, td, kind, cores = ref_times(sample, args.order, 2)
        line["cpu_baseline"] = {"value": 2.0 * sample.size / (tc + td) / 1e9, "unit": "GVox/s", "cores": cores, "kind": kind,
                                "sample": f"first {zs} z-slices ({sx}x{sy}x{zs}) of the same volume, all host threads",
                                "compress_gvox_s": sample.size / tc / 1e9, "decompress_gvox_s": sample.size / td / 1e9}
    if rank == 0:
        if args.prof:
            sys.stderr.write(json.dumps(stage_ms, indent=1) + "\n")
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
