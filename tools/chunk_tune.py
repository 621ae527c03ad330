#!/usr/bin/env python
"""GPU tuning aid: compress / decompress time of the device-resident bench volume for several z-chunk counts."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crackle_b200 as cb
from crackle_b200 import synth

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1024,1024,1024").split(","))
Ks = [int(k) for k in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
order = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sx, sy, sz = shape
vol = synth.jittered_voronoi_torch(shape, 24, np.uint64, seed=0, id_bits=40, device="cuda")
out = torch.empty_like(vol)
ctx = cb.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ref = None
for K in Ks:
    ctx.set_chunks(K)
    def comp():
        return ctx.compress_ptr(vol.data_ptr(), 1, 8, sx, sy, sz, True, order)
    def dec():
        p, n = ctx.result_device()
        ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)
    for _ in range(2):
        comp(); dec()
    torch.cuda.synchronize()
    b = ctx.result_bytes()
    if ref is None:
        ref = b
    ok = (b == ref) and torch.equal(out.view(torch.int64), vol.view(torch.int64))
    res = {}
    for name, fn in (("compress", comp), ("decompress", dec)):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(5):
            fn()
        e.record(stream)
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(e) / 5
    print(json.dumps({"K": K, "ok": bool(ok), "compress_ms": round(res["compress"], 3), "decompress_ms": round(res["decompress"], 3),
                      "gvox_s": round(2 * sx * sy * sz / ((res["compress"] + res["decompress"]) * 1e-3) / 1e9, 1)}), flush=True)
