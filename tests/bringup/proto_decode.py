"""CPU prototype of the scan-parallel crack-code decoder (design validation for ckl_decode.cu).

Order-0 stream of one slice -> crack planes, formulated the way the CUDA kernels do it:
  A. per 16-field word: prefix sum mod 4 inside the word (packed 2-bit adds), word totals scanned -> absolute moves;
     escape pairs: opp[i] = move[i] is the opposite of move[i-1]; second-of-pair S[i] = opp[i] & ~S[i-1], solved per
     32-field window with the add-carry trick; events = S positions (type b if move in {RIGHT,DOWN} else t)
  B. per segment (codepoints between two events, minus the dropped first-of-pair): displacement sum
  C. serial over events only (~4 % of the codepoints): chain starts from the BOC index, revisit stack, positions
  D. per segment: walk the moves from the segment start, set crack bits
Checked against the crack planes computed directly from the labels.  Bring-up tool only (imports oracle/)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O  # noqa: E402
import proto_trace as PT  # noqa: E402

L5 = 0x55555555
HA = 0xAAAAAAAA
M32 = 0xFFFFFFFF


def add4(a, b):
    """per-field (2-bit) addition mod 4 of two packed words"""
    s = a ^ b
    c = ((a & b & L5) << 1) & M32
    return s ^ c


def word_prefix(x):
    x = add4(x, (x << 2) & M32)
    x = add4(x, (x << 4) & M32)
    x = add4(x, (x << 8) & M32)
    x = add4(x, (x << 16) & M32)
    return x


def compact16(v):
    """bits at even positions (one per field) of a 32-bit word -> 16-bit mask"""
    r = 0
    for k in range(16):
        r |= ((v >> (2 * k)) & 1) << k
    return r


def byte_width(x):
    return 1 if x <= 0xFF else 2 if x <= 0xFFFF else 4 if x <= 0xFFFFFFFF else 8


def read_boc(code, sx, sy):
    xw, yw = byte_width(sx + 1), byte_width(sy + 1)
    isz = int.from_bytes(code[:4], "little")
    p = 4
    ny = int.from_bytes(code[p:p + yw], "little"); p += yw
    starts = []
    y = 0
    for _ in range(ny):
        y += int.from_bytes(code[p:p + yw], "little"); p += yw
        nx = int.from_bytes(code[p:p + xw], "little"); p += xw
        x = 0
        for _ in range(nx):
            x += int.from_bytes(code[p:p + xw], "little"); p += xw
            starts.append((x, y))
    return starts, 4 + isz


def decode_slice(code, sx, sy):
    starts, body_off = read_boc(code, sx, sy)
    body = code[body_off:]
    nwords = (len(body) + 3) // 4
    padded = body + b"\0" * (nwords * 4 - len(body))
    words = np.frombuffer(padded, dtype="<u4").astype(np.int64) if nwords else np.zeros(0, np.int64)
    ncp = len(body) * 4
    # ---- A: moves + events, one "thread" per word
    incl = [word_prefix(int(w)) for w in words]
    tot = [(x >> 30) & 3 for x in incl]
    e = np.concatenate([[0], np.cumsum(tot)])[:-1] & 3 if nwords else []
    Mw, Sw = [], []
    fallback = False
    for wi in range(nwords):
        M = add4(incl[wi], (int(e[wi]) * L5) & M32)
        # previous word's moves (recomputed locally, as the kernel does)
        if wi > 0:
            eprev = (int(e[wi]) - tot[wi - 1]) & 3
            Mp = add4(incl[wi - 1], (eprev * L5) & M32)
            Xp = Mp ^ (((Mp << 2) & M32) | eprev)
            Op = compact16((Xp >> 1) & ~Xp & L5)
            if wi - 1 == 0:
                Op &= ~1
        else:
            Op = 0
        X = M ^ (((M << 2) & M32) | int(e[wi]))
        Oc = compact16((X >> 1) & ~X & L5)
        if wi == 0:
            Oc &= ~1                       # the first codepoint has no predecessor
        Oc2 = Oc
        if Oc == 0xFFFF or (Op == 0xFFFF and (Oc & 1)):
            fallback = True
        Ocomb = Op | (Oc2 << 16)
        st = Ocomb & ~((Ocomb << 1) & M32)
        t = (Ocomb + (st & L5)) & M32
        ev_runs = Ocomb & ~t
        odd_runs = Ocomb & ~ev_runs
        S = ((ev_runs & L5) | (odd_runs & HA)) >> 16
        Mw.append(M); Sw.append(S)
    assert not fallback
    move = lambda i: (Mw[i >> 4] >> (2 * (i & 15))) & 3   # noqa: E731
    evidx = [wi * 16 + k for wi in range(nwords) for k in range(16) if (Sw[wi] >> k) & 1 and wi * 16 + k < ncp]
    DXY = {0: (0, -1), 1: (1, 0), 2: (0, 1), 3: (-1, 0)}
    # ---- B: segment sums
    seg = []
    for j, ei in enumerate(evidx):
        lo = 0 if j == 0 else evidx[j - 1] + 1
        dx = dy = 0
        for i in range(lo, ei - 1):
            a, b = DXY[move(i)]
            dx += a; dy += b
        seg.append((dx, dy))
    # ---- C: serial over events
    segstart = []
    openc, ci, stack = 0, 0, []
    x = y = 0
    used = len(evidx)
    for j, ei in enumerate(evidx):
        if openc == 0:
            if ci >= len(starts):
                used = j
                break
            x, y = starts[ci]; ci += 1
            openc = 1; stack = []
        segstart.append((x, y))
        x += seg[j][0]; y += seg[j][1]
        m = move(ei)
        if m in (0, 3):       # t
            openc -= 1
            if stack:
                loc = stack.pop()
                y = loc // sx; x = loc - y * sx
        else:
            openc += 1
            stack.append(x + sx * y)
    # ---- D: marking
    EH = np.zeros((sy + 1, sx + 1), bool)
    EV = np.zeros((sy + 1, sx + 1), bool)
    for j in range(used):
        lo = 0 if j == 0 else evidx[j - 1] + 1
        x, y = segstart[j]
        for i in range(lo, evidx[j] - 1):
            m = move(i)
            if m == 0:
                if 0 < x < sx: EV[y - 1, x] = True
                y -= 1
            elif m == 2:
                if 0 < x < sx: EV[y, x] = True
                y += 1
            elif m == 3:
                if 0 < y < sy: EH[y, x - 1] = True
                x -= 1
            else:
                if 0 < y < sy: EH[y, x] = True
                x += 1
    return EH, EV, dict(ncp=ncp, events=len(evidx), chains=len(starts))


def main():
    rng = np.random.default_rng(2)
    nbad = 0
    for t in range(40):
        sx, sy = int(rng.integers(2, 120)), int(rng.integers(2, 120))
        kind = t % 4
        if kind == 0:
            img = PT.voronoi(sx, sy, int(rng.integers(1, 40)), rng)
        elif kind == 1:
            img = rng.integers(0, 3, size=(sx, sy)).astype(np.uint32)
        elif kind == 2:
            img = rng.integers(0, 2000, size=(sx, sy)).astype(np.uint32)
        else:
            img = PT.voronoi(sx, sy, int(rng.integers(2, 60)), rng)
            img[rng.random(img.shape) < 0.05] = 0
        for perm in (0, 1):
            code = O.slice_crack_code(img, perm)
            EH, EV, info = decode_slice(code, sx, sy)
            wEH, wEV = PT.planes(img, perm)
            ok = np.array_equal(EH, wEH) and np.array_equal(EV, wEV)
            nbad += not ok
            print(img.shape, "perm", perm, "ok" if ok else "MISMATCH", info)
    print("bad:", nbad)
    return nbad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
