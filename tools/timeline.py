#!/usr/bin/env python
"""GPU debugging aid: per-stage begin/end times (CKL_TIMELINE=1) of one compress and one decompress for a chunk count."""
import sys, os
import numpy as np
os.environ["CKL_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crackle_b200 as cb
from crackle_b200 import synth
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1024,1024,1024").split(","))
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sx, sy, sz = shape
vol = synth.jittered_voronoi_torch(shape, 24, np.uint64, seed=0, id_bits=40, device="cuda")
out = torch.empty_like(vol)
ctx = cb.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_chunks(K)
for i in range(3):
    if i == 2:
        ctx.prof_enable(True)
        sys.stderr.write(f"=== compress K={K}\n")
    n = ctx.compress_ptr(vol.data_ptr(), 1, 8, sx, sy, sz, True, 0)
    p, n = ctx.result_device()
    if i == 2:
        sys.stderr.write(f"=== decompress K={K}\n")
    ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)
torch.cuda.synchronize()
