"""Deterministic synthetic segmentation volumes.

`jittered_voronoi` is the exact Voronoi diagram of one jittered seed per cell^3 grid cell, computed with
integer arithmetic only so that the numpy (CPU) and torch (GPU) twins produce identical voxels.  The random
draws (seed offsets, ids) always come from numpy's default_rng on the host: they are tiny."""
import numpy as np


def _seeds(shape, cell, seed, id_bits, z0=0, sz_total=None):
    sx, sy, sz = shape
    szt = sz if sz_total is None else sz_total
    rng = np.random.default_rng(seed)
    gx, gy, gz = (sx + cell - 1) // cell, (sy + cell - 1) // cell, (szt + cell - 1) // cell
    off = rng.integers(0, cell, size=(gx + 2, gy + 2, gz + 2, 3), dtype=np.int64)
    base = np.stack(np.meshgrid(np.arange(-1, gx + 1), np.arange(-1, gy + 1), np.arange(-1, gz + 1),
                                indexing="ij"), -1).astype(np.int64) * cell
    seeds = base + off
    nid = (gx + 2) * (gy + 2) * (gz + 2)
    if id_bits <= 16:
        ids = (rng.permutation(nid) + 1).astype(np.uint64)
    else:
        ids = rng.integers(1, 1 << id_bits, size=nid, dtype=np.uint64)
    return seeds, ids.reshape(gx + 2, gy + 2, gz + 2)


def jittered_voronoi(shape, cell, dtype, seed=0, id_bits=40, z0=0, sz_total=None):
    """numpy twin; returns an F-ordered (sx,sy,sz) array.  `z0`/`sz_total` select a z-slab of a taller
    volume (used for z-sharded multi-GPU inputs)."""
    sx, sy, sz = shape
    seeds, ids = _seeds(shape, cell, seed, id_bits, z0, sz_total)
    out = np.empty(shape, dtype=dtype, order="F")
    X = np.arange(sx, dtype=np.int64)[:, None]
    Y = np.arange(sy, dtype=np.int64)[None, :]
    cx, cy = X // cell + 1, Y // cell + 1
    big = np.iinfo(np.int64).max
    for zl in range(sz):
        z = zl + z0
        cz = z // cell + 1
        best = np.full((sx, sy), big, dtype=np.int64)
        lab = np.zeros((sx, sy), np.uint64)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    s = seeds[cx + dx, cy + dy, cz + dz]
                    d = (s[..., 0] - X) ** 2 + (s[..., 1] - Y) ** 2 + (s[..., 2] - z) ** 2
                    i = ids[cx + dx, cy + dy, cz + dz]
                    m = (d < best) | ((d == best) & (i < lab))
                    best = np.where(m, d, best)
                    lab = np.where(m, i, lab)
        out[:, :, zl] = lab.astype(dtype)
    return out


def jittered_voronoi_torch(shape, cell, dtype, seed=0, id_bits=40, device="cuda", z0=0, sz_total=None, zchunk=16):
    """torch twin (same integer math).  Returns a torch tensor of shape (sz, sy, sx), C-contiguous, i.e. the
    SAME MEMORY as an F-ordered (sx,sy,sz) numpy array: element (x,y,z) lives at x + sx*(y + sy*z)."""
    import torch
    sx, sy, sz = shape
    seeds_np, ids_np = _seeds(shape, cell, seed, id_bits, z0, sz_total)
    seeds = torch.from_numpy(seeds_np).to(device)
    ids = torch.from_numpy(ids_np.astype(np.int64)).to(device)     # ids < 2^63
    tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16, np.dtype(np.uint32): torch.uint32,
           np.dtype(np.uint64): torch.uint64}[np.dtype(dtype)]
    out = torch.empty((sz, sy, sx), dtype=tdt, device=device)
    X = torch.arange(sx, dtype=torch.int64, device=device)[None, None, :]
    Y = torch.arange(sy, dtype=torch.int64, device=device)[None, :, None]
    cx, cy = X // cell + 1, Y // cell + 1
    big = torch.iinfo(torch.int64).max
    for zs in range(0, sz, zchunk):
        ze = min(sz, zs + zchunk)
        Z = torch.arange(zs + z0, ze + z0, dtype=torch.int64, device=device)[:, None, None]
        cz = Z // cell + 1
        best = torch.full((ze - zs, sy, sx), big, dtype=torch.int64, device=device)
        lab = torch.zeros((ze - zs, sy, sx), dtype=torch.int64, device=device)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    s = seeds[cx + dx, cy + dy, cz + dz]
                    d = (s[..., 0] - X) ** 2 + (s[..., 1] - Y) ** 2 + (s[..., 2] - Z) ** 2
                    i = ids[cx + dx, cy + dy, cz + dz]
                    m = (d < best) | ((d == best) & (i < lab))
                    best = torch.where(m, d, best)
                    lab = torch.where(m, i, lab)
        out[zs:ze] = lab.to(tdt) if tdt != torch.uint64 else lab.view(torch.uint64)
    return out


def random_blobs(shape, nlabels, dtype, seed=0):
    """cheap smooth-ish segmentation for tests: nearest of `nlabels` uniform random seeds (brute force)."""
    rng = np.random.default_rng(seed)
    sx, sy, sz = shape
    pts = rng.integers(0, [sx, sy, sz], size=(nlabels, 3)).astype(np.int64)
    ids = (rng.permutation(nlabels) + 1).astype(np.uint64)
    g = np.stack(np.meshgrid(np.arange(sx), np.arange(sy), np.arange(sz), indexing="ij"), -1).astype(np.int64)
    d = ((g[..., None, :] - pts[None, None, None]) ** 2).sum(-1)
    return np.asfortranarray(ids[d.argmin(-1)].astype(dtype))
