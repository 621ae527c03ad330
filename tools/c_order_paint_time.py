"""GPU: time a C-order decompress (the tiled-transpose paint) of the bench volume against the Fortran-order one."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import crackle_b200 as cb
from crackle_b200 import synth

sz = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = cb.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
t = synth.jittered_voronoi_torch((1024, 1024, sz), 24, np.uint64, seed=0, id_bits=40, sz_total=1024)
out = torch.empty_like(t)
for forder in (True, False):
    n = ctx.compress_ptr(t.data_ptr(), 1, 8, 1024, 1024, sz, forder, 0)
    p, n = ctx.result_device()
    ctx.prof_enable(True)
    for _ in range(3):
        ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ctx.decompress_into(p, 1, n, 0, -1, None, out.data_ptr(), 1, out.numel() * 8)
    b.record(); torch.cuda.synchronize()
    prof = ctx.prof_read(); ctx.prof_enable(False)
    if forder:
        ok = torch.equal(out.view(torch.int64), t.view(torch.int64))
    else:   # C order: element (x,y,z) at z + sz*(y + sy*x)
        ok = torch.equal(out.view(torch.int64).view(1024, 1024, sz), t.view(torch.int64).permute(2, 1, 0))
    print("fortran" if forder else "C", "decompress ms", a.elapsed_time(b) / 5, "paint ms", prof["d_paint"][0] / prof["d_paint"][1], "ok", ok)
