// ckl_decode.cu -- decompress-side kernels: crack code -> crack bit-planes, label map resolution, paint.
//
// Reference behaviour reproduced:
//   read_boc_index                                   src/crackcodes.hpp:283-316
//   packed_codepoints_to_symbols (order 0 FSM)       src/crackcodes.hpp:523-603
//   markov::decode_codepoints + codepoints_to_symbols   src/markov.hpp:268-323, src/crackcodes.hpp:606-676
//   decode_{permissible,impermissible}_crack_code    src/crackcodes.hpp:706-862
//   labels::decode_flat                              src/labels.hpp:453-506
//   paint                                            src/crackle.hpp:617-656
// The voxel connectivity graph is kept as two crack bit-planes instead of a byte per pixel: the CCL only reads the
// "-x" and "-y" passable bits, which are exactly "no vertical crack at column x" / "no horizontal crack above y".
#include "ckl_internal.cuh"

static u32 grid1(u64 n, u32 bs, u32 cap_blocks) {
  u64 b = (n + bs - 1) / bs;
  if (b < 1) b = 1;
  if (b > cap_blocks) b = cap_blocks;
  return (u32)b;
}

__device__ __forceinline__ u64 ld_le(const u8* p, int w) {
  u64 v = 0;
  for (int i = 0; i < w; i++) v |= (u64)p[i] << (8 * i);
  return v;
}

struct DecodeState {
  const Geom* g;
  u32 *EV, *EH;
  u32 x, y;
  u32* stack;
  u32 sp, cap;
  bool bad;
};
__device__ __forceinline__ void mark_move(DecodeState& s, u32 m) {
  const Geom& g = *s.g;
  const u32 W = g.W;
  if (m == 0) {          // up: vertical crack at column x, row y-1
    if (s.y == 0) { s.bad = true; return; }
    if (s.x > 0 && s.x < g.sx) atomicOr(&s.EV[(u64)(s.y - 1) * W + (s.x >> 5)], 1u << (s.x & 31));
    s.y--;
  } else if (m == 2) {   // down: vertical crack at column x, row y
    if (s.y >= g.sy) { s.bad = true; return; }
    if (s.x > 0 && s.x < g.sx) atomicOr(&s.EV[(u64)s.y * W + (s.x >> 5)], 1u << (s.x & 31));
    s.y++;
  } else if (m == 3) {   // left: horizontal crack above pixel (x-1, y)
    if (s.x == 0) { s.bad = true; return; }
    if (s.y > 0 && s.y < g.sy) atomicOr(&s.EH[(u64)s.y * W + ((s.x - 1) >> 5)], 1u << ((s.x - 1) & 31));
    s.x--;
  } else {               // right: horizontal crack above pixel (x, y)
    if (s.x >= g.sx) { s.bad = true; return; }
    if (s.y > 0 && s.y < g.sy) atomicOr(&s.EH[(u64)s.y * W + (s.x >> 5)], 1u << (s.x & 31));
    s.x++;
  }
}

// beginning-of-chain index iterator
struct BocIter {
  const u8* p;
  u64 idx, end;
  int xw, yw;
  u32 ny, yi, nx, y, x, sxe;
  __device__ bool next(u32& vx, u32& vy) {
    while (nx == 0) {
      if (yi >= ny || idx + yw + xw > end) return false;
      y += (u32)ld_le(p + idx, yw); idx += yw;
      nx = (u32)ld_le(p + idx, xw); idx += xw;
      yi++;
      x = 0;
    }
    if (idx + xw > end) return false;
    x += (u32)ld_le(p + idx, xw); idx += xw;
    nx--;
    vx = x; vy = y;
    return true;
  }
};

// escape-pair state machine shared by both stream formats.  Returns false when decoding of the slice is finished.
struct Fsm {
  int last_move;     // 255 = NONE
  int pend;          // pending plain move not yet applied (-1 = none)
  u32 open;          // branches_taken
};
__device__ __forceinline__ bool fsm_feed(Fsm& f, DecodeState& s, BocIter& it, u32 move) {
  if (f.open == 0) {
    u32 vx, vy;
    if (!it.next(vx, vy)) return false;
    if (vx > s.g->sx || vy > s.g->sy) { s.bad = true; return false; }
    s.x = vx; s.y = vy; s.sp = 0;
    f.open = 1; f.pend = -1; f.last_move = 255;
  }
  if ((int)(move ^ (u32)f.last_move) != 2) {
    if (f.pend >= 0) mark_move(s, (u32)f.pend);
    f.pend = (int)move;
    f.last_move = (int)move;
    return !s.bad;
  }
  // opposite of the pending move: the pair is an escape.  UP / LEFT second -> 't', DOWN / RIGHT second -> 'b'
  f.pend = -1;
  f.last_move = 255;
  if (move == 0 || move == 3) {
    f.open--;
    if (s.sp > 0) {
      const u32 loc = s.stack[--s.sp];                 // reference quirk: pushed as x + sx*y (crackcodes.hpp:772,850)
      s.y = loc / s.g->sx;
      s.x = loc - s.y * s.g->sx;
    }
  } else {
    f.open++;
    if (s.sp >= s.cap) { s.bad = true; return false; }
    s.stack[s.sp++] = s.x + s.g->sx * s.y;
  }
  return !s.bad;
}

// LSB-first bit reader over a byte range: aligned 32-bit loads, one word prefetched ahead of use so the load
// latency overlaps the decode of the previous 16+ symbols; bits past the end read as zero.
struct BitReader {
  const u32* q;
  u64 nw, idx;
  u64 cur;
  u32 cnt, nxt, lastMask;
  __device__ __forceinline__ u32 fetch() {
    if (idx >= nw) return 0u;
    u32 v = __ldg(q + idx);
    if (idx == nw - 1) v &= lastMask;
    return v;
  }
  __device__ __forceinline__ void refill() {
    if (cnt <= 32) { cur |= (u64)nxt << cnt; cnt += 32; idx++; nxt = fetch(); }
  }
  __device__ void init(const u8* p, u64 nbytes) {
    const u32 a = (u32)((u64)p & 3);
    q = reinterpret_cast<const u32*>(p - a);
    nw = nbytes ? (a + nbytes + 3) / 4 : 0;
    const u32 vb = (u32)((a + nbytes - 1) & 3) + 1;             // valid bytes in the last word
    lastMask = vb == 4 ? 0xFFFFFFFFu : ((1u << (8 * vb)) - 1u);
    idx = 0; cur = 0; cnt = 0;
    nxt = fetch();
    refill();
    cur >>= 8 * a; cnt -= 8 * a;
    refill();
  }
  __device__ __forceinline__ u32 peek(u32 nb) const { return (u32)cur & ((1u << nb) - 1u); }
  __device__ __forceinline__ void skip(u32 nb) { cur >>= nb; cnt -= nb; refill(); }
};

#define DECODE_SMEM_MODEL 4096   // markov models up to order 5 are staged in shared memory

__global__ void __launch_bounds__(32) k_decode_slices(Geom g, const u8* __restrict__ stream, const u64* __restrict__ codeOff, int order,
                                                       const u8* __restrict__ model, u32* EVall, u32* EHall, u32* stackAll,
                                                       const u64* __restrict__ stackOff, ull* scal) {
  __shared__ u8 smodel[DECODE_SMEM_MODEL];
  const u32 z = blockIdx.x;
  const u32 mbytes = order > 0 ? (4u << (2 * order)) : 0u;
  const bool msm = order > 0 && mbytes <= DECODE_SMEM_MODEL;
  if (msm) {
    for (u32 i = threadIdx.x; i < mbytes; i += blockDim.x) smodel[i] = model[i];
    __syncwarp();
  }
  if (threadIdx.x != 0 || z >= g.sz) return;
  const u8* code = stream + codeOff[z];
  const u64 clen = codeOff[z + 1] - codeOff[z];
  const int xw = ckl_byte_width((u64)g.sx + 1), yw = ckl_byte_width((u64)g.sy + 1);
  if (clen < (u64)(4 + yw)) return;                    // no index: nothing to paint
  const u64 isz = 4 + ld_le(code, 4);
  if (isz > clen) { atomicExch(&scal[SC_ERROR], 10ull); return; }
  const u64 nw = (u64)g.sy * g.W;
  DecodeState s;
  s.g = &g; s.EV = EVall + (u64)z * nw; s.EH = EHall + (u64)z * nw;
  s.x = s.y = 0; s.stack = stackAll + stackOff[z]; s.sp = 0; s.cap = (u32)(stackOff[z + 1] - stackOff[z]); s.bad = false;
  BocIter it;
  it.p = code; it.idx = 4; it.end = isz; it.xw = xw; it.yw = yw; it.sxe = g.sx + 1;
  it.ny = (u32)ld_le(code + 4, yw); it.idx += yw; it.yi = 0; it.nx = 0; it.y = 0; it.x = 0;
  Fsm f; f.last_move = 255; f.pend = -1; f.open = 0;
  const u8* body = code + isz;
  const u64 blen = clen - isz;
  BitReader br;
  br.init(body, blen);
  if (order == 0) {
    u32 last = 0;
    const u64 nfields = blen * 4;
    for (u64 i = 0; i < nfields; i++) {
      last = (last + br.peek(2)) & 3u;
      br.skip(2);
      if (!fsm_feed(f, s, it, last)) break;
    }
  } else if (blen) {
    const u8* mdl = msm ? smodel : model;
    const u32 top = 2 * (order - 1);
    u32 mv = br.peek(2);
    br.skip(2);
    u32 ctx = mv << top;
    bool go = fsm_feed(f, s, it, mv);
    u64 pos = 2;
    const u64 nbit = blen * 8;
    while (go && pos < nbit) {
      const u32 v = br.peek(3);
      u32 rank, len;
      if (!(v & 1)) { rank = 0; len = 1; } else if (!(v & 2)) { rank = 1; len = 2; } else if (!(v & 4)) { rank = 2; len = 3; } else { rank = 3; len = 3; }
      const u32 d = mdl[(u64)ctx * 4 + rank];
      br.skip(len);
      pos += len;
      mv = (mv + d) & 3u;
      ctx = (ctx >> 2) + (d << top);
      go = fsm_feed(f, s, it, mv);
    }
  }
  if (s.bad) atomicExch(&scal[SC_ERROR], 11ull);
}

void launch_decode_slices(const Geom& g, const u8* stream, const u64* codeOff, int permissible, int order, const u8* model,
                          u32* EV, u32* EH, u32* stack, const u64* stackOff, ull* scal, cudaStream_t st) {
  (void)permissible;
  CUDA_CHECK(cudaMemsetAsync(EV, 0, g.words() * 4, st));
  CUDA_CHECK(cudaMemsetAsync(EH, 0, g.words() * 4, st));
  k_decode_slices<<<g.sz, 32, 0, st>>>(g, stream, codeOff, order, model, EV, EH, stack, stackOff, scal);
  LAUNCH_CHECK();
}

// crack planes -> "differ" planes.  IMPERMISSIBLE: identical.  PERMISSIBLE: cracks mark connected neighbours.
__global__ void __launch_bounds__(256) k_planes_from_cracks(Geom g, u32* EV, u32* EH) {
  const u64 nwords = g.words(), stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
    const u64 row = i / g.W;
    const u32 w = (u32)(i - row * g.W);
    const u32 y = (u32)(row % g.sy);
    const u32 valid = (w == g.W - 1 && (g.sx & 31)) ? ((1u << (g.sx & 31)) - 1u) : 0xFFFFFFFFu;
    EV[i] = ~EV[i] & valid & ~(w == 0 ? 1u : 0u);
    EH[i] = y >= 1 ? (~EH[i] & valid) : 0u;
  }
}
void launch_planes_from_cracks(const Geom& g, int permissible, u32* EV, u32* EH, cudaStream_t st) {
  if (!permissible) return;
  k_planes_from_cracks<<<grid1(g.words(), 256, 148 * 16), 256, 0, st>>>(g, EV, EH);
  LAUNCH_CHECK();
}

// label of every run: component rank -> key -> unique label (decode_flat)
__global__ void __launch_bounds__(256) k_run_labels(Geom g, u64 total_runs, const u64* __restrict__ runBase, const u32* __restrict__ runComp,
                                                     const u8* __restrict__ uniq, const u8* __restrict__ keys, u64 n_uniq, u64 n_keys,
                                                     int sw, int kw, const u64* __restrict__ keyBase, u64* __restrict__ runLabel) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < total_runs; r += stride) {
    u32 lo = 0, hi = g.sz;
    while (hi - lo > 1) { const u32 m = (lo + hi) >> 1; if (runBase[m] <= r) lo = m; else hi = m; }
    const u64 ki = keyBase[lo] + runComp[r];
    u64 label = 0;
    if (ki < n_keys) {
      const u64 k = ld_le(keys + ki * (u64)kw, kw);
      if (k < n_uniq) label = ld_le(uniq + k * (u64)sw, sw);
    }
    runLabel[r] = label;
  }
}
void launch_run_labels(const Geom& g, const CclBufs& B, const u8* stream, u64 uniq_off, u64 keys_off, u64 n_uniq, u64 n_keys_total,
                       int stored_width, int key_width, const u64* keyBase, u64* runLabel, cudaStream_t st) {
  // total runs lives in runBase[sz]; the caller passes it through B.runBase on the host side via scal; we re-read here
  u64 total_runs = 0;
  CUDA_CHECK(cudaMemcpyAsync(&total_runs, B.runBase.as<u64>() + g.sz, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (!total_runs) return;
  k_run_labels<<<grid1(total_runs, 256, 148 * 16), 256, 0, st>>>(g, total_runs, B.runBase.as<u64>(), B.runComp.as<u32>(),
                                                                 stream + uniq_off, stream + keys_off, n_uniq, n_keys_total,
                                                                 stored_width, key_width, keyBase, runLabel);
  LAUNCH_CHECK();
}

// paint: one warp per 32-pixel word, lane <-> pixel; the only full-width write of decompress.
template <typename OUT, bool MASK, bool FORTRAN>
__global__ void __launch_bounds__(256) k_paint(Geom g, const u32* __restrict__ DV, const u32* __restrict__ wordPrefix,
                                                const u32* __restrict__ rowBase, const u64* __restrict__ runBase,
                                                const u64* __restrict__ runLabel, u64 label, OUT* __restrict__ out) {
  const u32 lane = threadIdx.x & 31;
  const u64 nwords = g.words();
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  for (u64 i = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nwords; i += nwarps) {
    const u64 row = i / g.W;
    const u32 w = (u32)(i - row * g.W);
    const u32 z = (u32)(row / g.sy), y = (u32)(row - (u64)z * g.sy);
    const u32 x = w * 32 + lane;
    if (x >= g.sx) continue;
    const u32 dv = DV[i];
    const u32 rid = rowBase[row] + wordPrefix[i] + __popc(dv & ((2u << lane) - 1u));
    const u64 v = runLabel[runBase[z] + rid];
    const u64 o = FORTRAN ? ((u64)z * g.sxy + (u64)y * g.sx + x) : ((u64)z + (u64)g.sz * ((u64)y + (u64)g.sy * x));
    out[o] = MASK ? (OUT)(v == label) : (OUT)v;
  }
}

void launch_paint(const Geom& g, const u32* DV, const CclBufs& B, const u64* runLabel, int out_width, int has_label, u64 label,
                  int fortran_order, void* out, cudaStream_t st) {
  const u32 grid = grid1(g.words(), 8, 148 * 8);
  const u32* wp = B.wordPrefix.as<u32>();
  const u32* rb = B.rowBase.as<u32>();
  const u64* rB = B.runBase.as<u64>();
#define PAINT(T, M, F) k_paint<T, M, F><<<grid, 256, 0, st>>>(g, DV, wp, rb, rB, runLabel, label, (T*)out)
  if (has_label) {
    if (fortran_order) PAINT(u8, true, true); else PAINT(u8, true, false);
  } else if (fortran_order) {
    switch (out_width) { case 1: PAINT(u8, false, true); break; case 2: PAINT(u16, false, true); break;
                         case 4: PAINT(u32, false, true); break; default: PAINT(u64, false, true); break; }
  } else {
    switch (out_width) { case 1: PAINT(u8, false, false); break; case 2: PAINT(u16, false, false); break;
                         case 4: PAINT(u32, false, false); break; default: PAINT(u64, false, false); break; }
  }
#undef PAINT
  LAUNCH_CHECK();
}
