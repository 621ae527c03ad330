"""CPU prototype of the contracted-graph chain tracer (design validation for ckl_trace.cu).

Pipeline mirrored by the CUDA kernels:
  1. crack planes -> per-vertex edge bits (r, d, u) + node mask n
       node = static degree in {1,3,4}, or an (R,D)-only corner whose horizontal run to the right does not end at
       a vertex with an up edge (every crack-graph component's minimum vertex is such a vertex or has degree 1)
  2. nodes numbered in raster order
  3. one walker per (node, direction) slot follows degree-2 vertices to the far node -> seFar, seLen
  4. serial replay of the reference walk (crackcodes.hpp:390-450) on the contracted graph -> event list
  5. parallel post: codepoint offsets, initial-branch reversal, escape coding, chain order -> codepoints
Checked byte-for-byte against the oracle's per-slice crack code.  Bring-up tool only (imports oracle/)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

R, L, D, U = 0, 1, 2, 3
DX = [1, -1, 0, 0]
DY = [0, 0, 1, -1]
OPP = [1, 0, 3, 2]
CODE = {R: 1, L: 3, D: 2, U: 0}      # crackcodes.hpp:20-26: UP=0 RIGHT=1 DOWN=2 LEFT=3
EV_E, EV_B, EV_T, EV_S = 0, 1, 2, 3


def planes(img, perm):
    """EH[y][x]: right edge of vertex (x,y); EV[y][x]: down edge of vertex (x,y).  img[x,y]."""
    sx, sy = img.shape
    EH = np.zeros((sy + 1, sx + 1), bool)
    EV = np.zeros((sy + 1, sx + 1), bool)
    dv = img[1:, :] != img[:-1, :]      # x>=1
    dh = img[:, 1:] != img[:, :-1]      # y>=1
    if perm:
        dv, dh = ~dv, ~dh
    EV[:sy, 1:sx] = dv.T
    EH[1:sy, :sx] = dh.T
    return EH, EV


def trace_slice(img, perm, check_local_top=True):
    sx, sy = img.shape
    sxe = sx + 1
    EH, EV = planes(img, perm)
    r = EH
    l = np.zeros_like(r); l[:, 1:] = r[:, :-1]
    d = EV
    u = np.zeros_like(d); u[1:, :] = d[:-1, :]
    deg = r.astype(int) + l + d + u
    node = (deg == 1) | (deg >= 3)
    # (R,D) corners: keep as node unless the horizontal run to the right ends at a vertex with an up edge
    corner = r & d & ~l & ~u
    for y, x in zip(*np.nonzero(corner)):
        xx = x + 1
        while r[y, xx] and not d[y, xx] and not u[y, xx]:
            xx += 1
        # CUDA version only looks inside the 32-bit word (conservative = keep); emulate that
        if check_local_top and (xx >> 5) == (x >> 5) and u[y, xx]:
            continue
        node[y, x] = True
    ys, xs = np.nonzero(node)            # raster order (row-major over y then x)
    nid = -np.ones(node.shape, int)
    nid[ys, xs] = np.arange(len(ys))
    nn = len(ys)
    adjbits = [r, l, d, u]
    # 3. path walk
    seFar = -np.ones((nn, 4), int)
    seFk = np.zeros((nn, 4), int)
    seLen = np.zeros((nn, 4), int)
    for i in range(nn):
        for k in range(4):
            if not adjbits[k][ys[i], xs[i]]:
                continue
            x, y, kk, n = xs[i], ys[i], k, 0
            while True:
                x += DX[kk]; y += DY[kk]; n += 1
                cf = OPP[kk]
                if node[y, x]:
                    break
                # degree 2: exit = other set bit among R, D, U else L
                a = [q for q in (R, D, U) if q != cf and adjbits[q][y, x]]
                kk = a[0] if a else L
            seFar[i, k] = nid[y, x]; seFk[i, k] = cf; seLen[i, k] = n
    # 4. replay
    adj = np.array([sum((1 << k) for k in range(4) if seFar[i, k] >= 0) for i in range(nn)], int)
    events = []          # [type, slot]
    chains = []          # dict(begin, end, t2f, adjStart)
    cursor = 0
    while True:
        while cursor < nn and adj[cursor] == 0:
            cursor += 1
        if cursor >= nn:
            break
        node_i = cursor
        begin = len(events)
        stack = []
        nB = 0
        firstT, t2, t2f = True, False, 0
        justPopped, poppedB = False, 0
        adjStart = xs[node_i] + sxe * ys[node_i]
        while True:
            a = int(adj[node_i])
            if a == 0:
                if firstT:
                    firstT = False
                    if nB == 1 and events[begin][0] == EV_B:
                        t2, t2f = True, len(events) - begin
                        adjStart = xs[node_i] + sxe * ys[node_i]
                if justPopped and not (t2 and poppedB == begin):
                    events[poppedB][0] = EV_S
                    events.append([EV_S, 0])
                else:
                    events.append([EV_T, 0])
                if not stack:
                    break
                node_i, poppedB = stack.pop()
                justPopped = True
                continue
            justPopped = False
            if a & (a - 1):
                stack.append((node_i, len(events)))
                events.append([EV_B, 0])
                nB += 1
            k = (a & -a).bit_length() - 1
            adj[node_i] &= ~(1 << k)
            far, fk = seFar[node_i, k], seFk[node_i, k]
            adj[far] &= ~(1 << fk)
            events.append([EV_E, node_i * 4 + k])
            node_i = far
        chains.append(dict(begin=begin, end=len(events), t2f=t2f if t2 else 0, adjStart=adjStart))
    # 5. post (per chain; every step below is a scan / independent per event on the GPU)
    out_chains = []
    for c in chains:
        ev = events[c["begin"]:c["end"]]
        ne = len(ev)
        # effective type (remove_initial_branch turns event 0 and t2f into 's')
        typ = [e[0] for e in ev]
        if c["t2f"]:
            typ[0] = EV_S; typ[c["t2f"]] = EV_S
        ln = [seLen[e[1] >> 2, e[1] & 3] if t == EV_E else (2 if t in (EV_B, EV_T) else 0) for e, t in zip(ev, typ)]
        off = np.concatenate([[0], np.cumsum(ln)])
        ncp = off[-1]
        cp = -np.ones(ncp, int)
        fcp = off[c["t2f"]] if c["t2f"] else 0     # codepoints in [0, fcp) are the reversed+flipped initial branch
        for j, (e, t) in enumerate(zip(ev, typ)):
            if t == EV_S:
                continue
            if t == EV_E:
                i, k = e[1] >> 2, e[1] & 3
                x, y, kk = xs[i], ys[i], k
                rev = c["t2f"] and j < c["t2f"]
                for q in range(ln[j]):
                    pos = off[j] + q
                    if rev:
                        cp[fcp - 1 - pos] = CODE[OPP[kk]]
                    else:
                        cp[pos] = CODE[kk]
                    x += DX[kk]; y += DY[kk]
                    cf = OPP[kk]
                    if node[y, x]:
                        break
                    a = [w for w in (R, D, U) if w != cf and adjbits[w][y, x]]
                    kk = a[0] if a else L
                continue
            # b / t: look back over events for the previous kept symbol
            tcount = 0
            prev = None        # previous kept move direction (as dir index) or None at chain start
            jj = j
            while jj > 0:
                jj -= 1
                if typ[jj] == EV_S:
                    continue
                if typ[jj] == EV_T:
                    tcount += 1
                    continue
                if typ[jj] == EV_B:
                    prev = "b"
                    break
                # E event: its last move (or, inside the reversed range, the flipped FIRST move of the first E)
                i, k = ev[jj][1] >> 2, ev[jj][1] & 3
                if c["t2f"] and jj < c["t2f"]:
                    # reversed range ends with flip(first move of the first E event after begin)
                    i0, k0 = ev[1][1] >> 2, ev[1][1] & 3
                    prev = OPP[k0]
                else:
                    prev = OPP[seFk[i, k]]
                break
            first_in_chain = (j == 0)
            if t == EV_B:
                alt = first_in_chain or (tcount == 0 and prev == D)
                a2 = (3, 1) if alt else (0, 2)
            else:
                alt = (1 if prev == U else 0) ^ (tcount & 1)
                a2 = (1, 3) if alt else (2, 0)
            cp[off[j]], cp[off[j] + 1] = a2
        assert (cp >= 0).all()
        out_chains.append((c["adjStart"], cp))
    out_chains.sort(key=lambda t: t[0])
    return out_chains, dict(nodes=nn, events=len(events), edges=int(EH.sum() + EV.sum()))


def byte_width(x):
    return 1 if x <= 0xFF else 2 if x <= 0xFFFF else 4 if x <= 0xFFFFFFFF else 8


def slice_code(img, perm, **kw):
    sx, sy = img.shape
    sxe = sx + 1
    chains, info = trace_slice(img, perm, **kw)
    xw, yw = byte_width(sx + 1), byte_width(sy + 1)
    rows = {}
    for s, _ in chains:
        rows.setdefault(int(s) // sxe, []).append(int(s) % sxe)
    body = bytearray()
    body += len(rows).to_bytes(yw, "little")
    py = 0
    for y in sorted(int(v) for v in rows):
        body += (y - py).to_bytes(yw, "little"); py = y
        body += len(rows[y]).to_bytes(xw, "little")
        px = 0
        for x in rows[y]:
            body += (x - px).to_bytes(xw, "little"); px = x
    out = bytearray(len(body).to_bytes(4, "little")) + body
    last, acc, pos = 0, 0, 0
    for _, cp in chains:
        for c in cp:
            dd = (int(c) - last) & 3; last = int(c)
            acc |= dd << pos; pos += 2
            if pos == 8:
                out.append(acc); acc = 0; pos = 0
    if pos:
        out.append(acc)
    return bytes(out), info


def voronoi(sx, sy, n, rng):
    pts = rng.integers(0, [sx, sy], size=(n, 2))
    xx, yy = np.meshgrid(np.arange(sx), np.arange(sy), indexing="ij")
    d = (xx[..., None] - pts[:, 0]) ** 2 + (yy[..., None] - pts[:, 1]) ** 2
    return (np.argmin(d, axis=2) + 1).astype(np.uint32)


def main():
    rng = np.random.default_rng(1)
    cases = []
    for t in range(40):
        sx, sy = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        kind = t % 5
        if kind == 0:
            img = voronoi(sx, sy, int(rng.integers(1, 30)), rng)
        elif kind == 1:
            img = rng.integers(0, 3, size=(sx, sy)).astype(np.uint32)
        elif kind == 2:
            img = rng.integers(0, 2000, size=(sx, sy)).astype(np.uint32)
        elif kind == 3:
            img = np.zeros((sx, sy), np.uint32)
            for _ in range(int(rng.integers(1, 12))):       # islands / nested blobs
                x0, y0 = int(rng.integers(0, sx)), int(rng.integers(0, sy))
                w, h = int(rng.integers(1, 20)), int(rng.integers(1, 20))
                img[x0:x0 + w, y0:y0 + h] = rng.integers(1, 5)
        else:
            img = voronoi(sx, sy, int(rng.integers(2, 60)), rng)
            img[rng.random(img.shape) < 0.05] = 0
        cases.append(img)
    cases.append(voronoi(200, 160, 60, rng))
    cases.append(voronoi(257, 129, 40, rng))
    nbad = 0
    for img in cases:
        for perm in (0, 1):
            want = O.slice_crack_code(img, perm)
            got, info = slice_code(img, perm)
            ok = got == want
            nbad += not ok
            print(img.shape, "perm", perm, "ok" if ok else "MISMATCH", info, len(want))
    print("bad:", nbad)
    return nbad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
