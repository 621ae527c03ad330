#!/usr/bin/env python
"""Structure of the reference walk (crackcodes.hpp:390-450) on one slice of the bench volume, measured on the CPU.

The walk is a lexicographic depth-first search of the crack graph (edge priority right, left, down, up; every edge is
erased when taken).  This script contracts the graph to its junctions, replays the walk and prints what decides whether
the serial replay kernel (k_replay) could be split into independent pieces:
  * length of the first greedy trail (root to the first dead end) and of the sub-searches hanging off it
  * depth of the revisit stack, lengths of the runs of consecutive 't's
DESIGN.md section 4 quotes the numbers.  usage: tools/dfs_structure.py [sx,sy] [cell] [z]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crackle_b200.synth import jittered_voronoi

DX = [1, -1, 0, 0]; DY = [0, 0, 1, -1]; OPP = [1, 0, 3, 2]      # right, left, down, up


def adjacency(img):
    sx, sy = img.shape
    r = np.zeros((sy + 1, sx + 1), bool); d = np.zeros((sy + 1, sx + 1), bool)
    d[:sy, 1:sx] = (img[1:, :] != img[:-1, :]).T
    r[1:sy, :sx] = (img[:, 1:] != img[:, :-1]).T
    l = np.zeros_like(r); l[:, 1:] = r[:, :-1]
    u = np.zeros_like(d); u[1:, :] = d[:-1, :]
    return r.astype(np.uint8) | (l.astype(np.uint8) << 1) | (d.astype(np.uint8) << 2) | (u.astype(np.uint8) << 3)


def contract(adj):
    deg = np.array([bin(i).count("1") for i in range(16)])[adj]
    start = np.argwhere(adj > 0)[0]                                   # raster-minimum vertex: the chain start
    node = (deg == 1) | (deg >= 3)
    node[start[0], start[1]] = True
    ys, xs = np.nonzero(node)
    ids = {(int(y), int(x)): i for i, (y, x) in enumerate(zip(ys, xs))}
    far = -np.ones((len(ys), 4), int); fdir = -np.ones((len(ys), 4), int)
    for i, (y, x) in enumerate(zip(ys, xs)):
        for k in range(4):
            if not (int(adj[y, x]) >> k) & 1:
                continue
            cy, cx, ck = int(y) + DY[k], int(x) + DX[k], k
            while (cy, cx) not in ids:
                rem = int(adj[cy, cx]) & ~(1 << OPP[ck])
                ck = rem.bit_length() - 1
                cy += DY[ck]; cx += DX[ck]
            far[i, k] = ids[(cy, cx)]; fdir[i, k] = OPP[ck]
    return far, fdir, adj[ys, xs].astype(int)


def walk(far, fdir, adj):
    adj = adj.copy(); ev = []; cursor = 0
    while True:
        while cursor < len(adj) and adj[cursor] == 0:
            cursor += 1
        if cursor >= len(adj):
            return ev
        node, stack = cursor, []
        while True:
            a = int(adj[node])
            if a == 0:
                ev.append(("T", len(stack)))
                if not stack:
                    break
                node = stack.pop(); continue
            if a & (a - 1):
                ev.append(("B", len(stack))); stack.append(node)
            k = (a & -a).bit_length() - 1
            ev.append(("E", k))
            adj[node] &= ~(1 << k)
            f = far[node, k]; adj[f] &= ~(1 << fdir[node, k]); node = f


if __name__ == "__main__":
    sx, sy = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1024,1024").split(","))
    cell = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    z = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    img = jittered_voronoi((sx, sy, 1), cell, np.uint64, seed=0, z0=z, sz_total=max(z + 1, 1024))[:, :, 0]
    far, fdir, a0 = contract(adjacency(img))
    ev = walk(far, fdir, a0)
    chains, cur = [], []
    for e in ev:
        cur.append(e)
        if e[0] == "T" and e[1] == 0:
            chains.append(cur); cur = []
    big = max(chains, key=len)
    iters = lambda seq: sum(1 for e in seq if e[0] != "B")             # one serial step per move and per 't'
    i0 = next(i for i, e in enumerate(big) if e[0] == "T")
    first = big[:i0]
    floor = sum(1 for e in first if e[0] == "B")
    sizes, cur = [], 0
    for e in big[i0:]:
        if e[0] == "T" and e[1] == floor:                              # the stack holds first-trail nodes only: a top-level pop
            if cur:
                sizes.append(cur)
            cur, floor = 1, floor - 1
        elif e[0] != "B":
            cur += 1
    if cur:
        sizes.append(cur)
    bursts, cur = [], 0
    for e in ev:
        if e[0] == "T":
            cur += 1
        elif cur:
            bursts.append(cur); cur = 0
    print(f"slice {sx}x{sy} cell {cell}: {len(a0)} junction / end nodes, {len(chains)} chains, {iters(ev)} serial steps "
          f"({sum(1 for e in ev if e[0] == 'E')} moves, {sum(1 for e in ev if e[0] == 'T')} t)")
    print(f"first greedy trail: {iters(first)} steps, {sum(1 for e in first if e[0] == 'B')} branch points")
    s = np.sort(np.array(sizes))[::-1]
    print(f"sub-searches hanging off the first trail: {len(s)}; largest {s[0]} steps = {100.0 * s[0] / max(1, s.sum()):.1f} % of the rest; next {s[1:4].tolist()}")
    print(f"revisit stack depth: max {max(e[1] for e in ev if e[0] != 'E')}; runs of consecutive t: {len(bursts)}, mean {np.mean(bursts):.2f}, max {max(bursts)}")
