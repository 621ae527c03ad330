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
#include <cstdlib>

#include <algorithm>

#include "ckl_internal.cuh"

static u32 grid1(u64 n, u32 bs, u32 cap_blocks) {
  u64 b = (n + bs - 1) / bs;
  if (b < 1) b = 1;
  if (b > (u64)cap_blocks * (u64)g_ckl_grid_mult) b = (u64)cap_blocks * (u64)g_ckl_grid_mult;
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
  const u64 mbytes = order > 0 ? (4ull << (2 * order)) : 0ull;      // order <= 12 (checked by the caller)
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

void launch_exscan_u32_u64(const u32* in, u32 n, u32 stride, u64* out, ull* total_out, u64 add_each, cudaStream_t st);

void launch_decode_slices(const Geom& g, const u8* stream, const u64* codeOff, int permissible, int order, const u8* model,
                          u32* EV, u32* EH, u32* stack, const u64* stackOff, ull* scal, cudaStream_t st) {
  (void)permissible;
  CUDA_CHECK(cudaMemsetAsync(EV, 0, g.words() * 4, st));
  CUDA_CHECK(cudaMemsetAsync(EH, 0, g.words() * 4, st));
  k_decode_slices<<<g.sz, 32, 0, st>>>(g, stream, codeOff, order, model, EV, EH, stack, stackOff, scal);
  LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// Scan-parallel decoder (validated against the oracle by tests/bringup/proto_decode.py).
//   k_mk_scan / k_mk_decode   order > 0 only: parallel bit decoder -> packed 2-bit difference fields
//   k_dec_classify per 16-field word: absolute moves (prefix sum mod 4) and the second-of-escape-pair mask S
//                  (S[i] = opp[i] & ~S[i-1], solved per 32-field window with an add-carry trick); counts events
//   k_dec_compact  event list: codepoint index | move << 30
//                  and Q = running displacement of all non-event codepoints (block scan)
//   k_dec_chain    serial over EVENTS only (~4 % of the codepoints): BOC chain starts, revisit stack, positions;
//                  the displacement of a segment (plain moves between two events; the codepoint before an event is
//                  the dropped first-of-pair) is a difference of two Q values
//   k_dec_mark     one thread per 16-codepoint word: position from Q and the segment start, set the crack bits
#define L5 0x55555555u
__device__ __forceinline__ u32 add4(u32 a, u32 b) { return (a ^ b) ^ ((a & b & L5) << 1); }     // per 2-bit field, mod 4
__device__ __forceinline__ u32 word_prefix4(u32 x) {
  x = add4(x, x << 2); x = add4(x, x << 4); x = add4(x, x << 8); x = add4(x, x << 16);
  return x;
}
// fields whose move is the opposite of the previous move (flag in the low bit of the field)
__device__ __forceinline__ u32 opp_flags(u32 M, u32 prev_move) {
  const u32 X = M ^ ((M << 2) | prev_move);
  return (X >> 1) & ~X & L5;
}

// 32-bit word `wi` of a slice's field stream (zero past the end)
__device__ __forceinline__ u32 dec_word(const DecSlice& d, const u8* __restrict__ stream, const u32* __restrict__ fields, int order, u64 wi) {
  if (order > 0) return fields[d.wordOff + wi];
  const u8* body = stream + d.body;
  const u32 a = (u32)((u64)body & 3);
  const u32* base = reinterpret_cast<const u32*>(body - a);
  const u64 nb = (u64)d.blen;
  if (wi * 4 >= nb) return 0u;
  u32 lo = __ldg(base + wi), hi = 0;
  if (a && (wi + 1) * 4 < a + nb) hi = __ldg(base + wi + 1);
  u32 w = a ? __funnelshift_r(lo, hi, 8 * a) : lo;
  const u64 rem = nb - wi * 4;
  if (rem < 4) w &= (1u << (8 * (u32)rem)) - 1u;
  return w;
}

// ---- order > 0, parallel form -------------------------------------------------------------------------------
// The code 0 / 10 / 110 / 111 (ranks 0..3, markov.hpp:444-458) is self-synchronising: a 0 bit always ends a code, and ones
// after a boundary are consumed three at a time, so the parser state at any bit position is (length of the run of ones
// right before it, not reaching below bit 2) mod 3 -- a local computation.  Pass 1 (k_mk_scan): per 32-bit word of a
// slice's bitstream, entry state and number of codes that START in the word; a block scan numbers the symbols.  Pass 2
// (k_mk_decode): 256 threads per slice decode disjoint word ranges.  The context chain (symbol -> context -> model row ->
// next symbol) is the one truly serial thing: every thread first guesses its entry context by decoding a warm-up stretch
// before its range, then the guesses are checked against the exit context of the thread before and wrong ones redone until
// all agree -- the result is exactly the serial decode (markov.hpp:268-323), whatever the guesses were.
struct MkCursor {
  u64 wi;            // next word to fetch
  u32 cur, nxt, bp;  // window = (nxt : cur) >> bp
};
__device__ __forceinline__ void mk_seek(MkCursor& c, const DecSlice& d, const u8* __restrict__ stream, u32 pos) {
  const u64 w = pos >> 5;
  c.cur = dec_word(d, stream, nullptr, 0, w); c.nxt = dec_word(d, stream, nullptr, 0, w + 1); c.wi = w + 2; c.bp = pos & 31u;
}
// length and rank of the code at the cursor: 0 -> (1, 0); 10 -> (2, 1); 110 -> (3, 2); 111 -> (3, 3)
__device__ __forceinline__ void mk_peek(const MkCursor& c, u32& len, u32& rank) {
  const u32 v2 = (__funnelshift_r(c.cur, c.nxt, c.bp) & 7u) * 2u;
  len = (0xD9D9u >> v2) & 3u; rank = (0xC484u >> v2) & 3u;
}
__device__ __forceinline__ void mk_skip(MkCursor& c, const DecSlice& d, const u8* __restrict__ stream, u32 len) {
  c.bp += len;
  if (c.bp >= 32) { c.bp -= 32; c.cur = c.nxt; c.nxt = dec_word(d, stream, nullptr, 0, c.wi++); }
}
// parser state at the start of word w (w >= 1): run of ones ending at bit 32 w - 1, not reaching below bit 2
__device__ __forceinline__ u32 mk_entry_state(const DecSlice& d, const u8* __restrict__ stream, u64 w) {
  u32 run = 0;
  for (u64 q = w; q > 0;) {
    q--;
    u32 x = dec_word(d, stream, nullptr, 0, q);
    if (q == 0) x &= ~3u;                                       // the two raw bits of the first symbol are a boundary
    const u32 ones = __clz(~x);                                 // leading (most significant = latest) ones of the word
    run += ones;
    if (ones < 32) break;
  }
  return run % 3u;
}
#define MK_WORD_STATE_SHIFT 30
__global__ void __launch_bounds__(256) k_mk_scan(const DecSlice* __restrict__ ds, u32 sz, const u8* __restrict__ stream,
                                                  u32* __restrict__ wordSym, u32* __restrict__ ncpOut) {
  __shared__ u32 sm[33];
  for (u32 z = blockIdx.x; z < sz; z += gridDim.x) {
    const DecSlice d = ds[z];
    const u32 nbits = d.blen * 8u;
    const u32 nW = (d.blen + 3u) / 4u;
    u32* ws = wordSym + d.wordOff;
    u32 carry = 0;
    for (u32 w0 = 0; w0 < nW; w0 += blockDim.x) {
      const u32 w = w0 + threadIdx.x;
      u32 cnt = 0, st = 0;
      if (w < nW) {
        u32 pos = 2;
        if (w > 0) { st = mk_entry_state(d, stream, w); pos = 32u * w - st; }
        MkCursor c;
        mk_seek(c, d, stream, pos);
        u32 len, rank;
        if (w > 0 && st > 0) { mk_peek(c, len, rank); mk_skip(c, d, stream, len); pos += len; }      // the code straddling in from the word before
        const u32 end = min(32u * (w + 1u), nbits);
        while (pos < end) { mk_peek(c, len, rank); mk_skip(c, d, stream, len); pos += len; cnt++; }
      }
      u32 tot;
      const u32 ex = carry + block_excl_scan(cnt, sm, tot);
      if (w < nW) ws[w] = ex | (st << MK_WORD_STATE_SHIFT);
      carry += tot;
    }
    if (threadIdx.x == 0) ncpOut[z] = d.blen ? carry + 1u : 0u;      // + the first symbol (two raw bits)
    __syncthreads();
  }
}

// codes that start in words [a, b) of the slice, decoded from entry context ctx4 (pre-multiplied by 4); returns the exit context
template <bool WRITE>
__device__ __forceinline__ u32 mk_run(const DecSlice& d, const u8* __restrict__ stream, const u32* __restrict__ ws, u32 a, u32 b, u32 nbits,
                                      u32 ctx4, const u8* mdl, u32 top2, u32* __restrict__ out) {
  if (a >= b) return ctx4;
  u32 pos = 2, st = 0;
  if (a > 0) { st = ws[a] >> MK_WORD_STATE_SHIFT; pos = 32u * a - st; }
  MkCursor c;
  mk_seek(c, d, stream, pos);
  u32 len, rank;
  if (a > 0 && st > 0) { mk_peek(c, len, rank); mk_skip(c, d, stream, len); pos += len; }
  const u32 end = min(32u * b, nbits);
  u32 idx = 1u + (ws[a] & ((1u << MK_WORD_STATE_SHIFT) - 1u));
  u32 acc = 0, accw = idx >> 4;
  while (pos < end) {
    mk_peek(c, len, rank);
    const u32 dsym = mdl[ctx4 + rank];
    ctx4 = ((ctx4 >> 2) & ~3u) + (dsym << top2);
    if (WRITE) {
      if ((idx >> 4) != accw) { if (acc) atomicOr(out + accw, acc); acc = 0; accw = idx >> 4; }
      acc |= dsym << (2u * (idx & 15u));
    }
    idx++;
    mk_skip(c, d, stream, len);
    pos += len;
  }
  if (WRITE && acc) atomicOr(out + accw, acc);
  return ctx4;
}
#define MK_WARM 3u            // words of warm-up before a thread's range (about 60 symbols)
__global__ void __launch_bounds__(256) k_mk_decode(const DecSlice* __restrict__ ds, u32 sz, const u8* __restrict__ stream, int order,
                                                    const u8* __restrict__ model, const u32* __restrict__ wordSym, u32* __restrict__ fields) {
  __shared__ u8 smodel[DECODE_SMEM_MODEL];
  __shared__ u32 s_end[256];
  const u64 mbytes = 4ull << (2 * order);      // order <= 12 (checked by the caller)
  const bool msm = mbytes <= DECODE_SMEM_MODEL;
  if (msm) for (u32 i = threadIdx.x; i < mbytes; i += blockDim.x) smodel[i] = model[i];
  __syncthreads();
  const u8* mdl = msm ? smodel : model;
  const u32 top2 = 2 * (order - 1) + 2;
  const u32 t = threadIdx.x;
  for (u32 z = blockIdx.x; z < sz; z += gridDim.x) {
    const DecSlice d = ds[z];
    const u32 nbits = d.blen * 8u;
    const u32 nW = (d.blen + 3u) / 4u;
    const u32* ws = wordSym + d.wordOff;
    u32* out = fields + d.wordOff;
    if (nW) {
      const u32 first = dec_word(d, stream, nullptr, 0, 0) & 3u;      // first symbol: two raw bits; it seeds the context
      const u32 ctx0 = first << top2;
      if (t == 0) atomicOr(out, first);
      const u32 wpt = (nW + blockDim.x - 1) / blockDim.x;
      const u32 a = min(nW, t * wpt), b = min(nW, a + wpt);
      // entry context: exact for the thread that starts the stream, a warm-up guess for the others
      u32 start = ctx0;
      if (a > 0) {
        const u32 aw = a > MK_WARM ? a - MK_WARM : 0u;
        start = mk_run<false>(d, stream, ws, aw, a, nbits, aw == 0 ? ctx0 : 0u, mdl, top2, nullptr);
      }
      u32 endc = mk_run<false>(d, stream, ws, a, b, nbits, start, mdl, top2, nullptr);
      for (;;) {                                                      // settle: entry context == exit context of the thread before
        s_end[t] = endc;
        __syncthreads();
        const u32 want = t == 0 ? ctx0 : s_end[t - 1];
        const bool redo = want != start;
        if (redo) { start = want; endc = mk_run<false>(d, stream, ws, a, b, nbits, start, mdl, top2, nullptr); }
        if (!__syncthreads_or(redo ? 1 : 0)) break;
      }
      mk_run<true>(d, stream, ws, a, b, nbits, start, mdl, top2, out);
    }
    __syncthreads();
  }
}

// displacement of the non-event fields [0, nf) of a word (moves: 0 up, 1 right, 2 down, 3 left), packed dy * 65536 + dx
__device__ __forceinline__ int word_disp(u32 M, u32 S, u32 nf) {
  u32 rm = L5 & ~S;
  if (nf < 16) rm &= (1u << (2 * nf)) - 1u;
  const u32 b0 = M & L5, b1 = (M >> 1) & L5;
  const int dx = __popc(b0 & ~b1 & rm) - __popc(b0 & b1 & rm);
  const int dy = __popc(b1 & ~b0 & rm) - __popc(~b0 & ~b1 & rm);
  return dy * 65536 + dx;
}
__device__ __forceinline__ int2 unpack_disp(int v) {
  const int dx = (int)(short)(v & 0xFFFF);
  return make_int2(dx, (v - dx) >> 16);
}

__global__ void __launch_bounds__(256) k_dec_classify(const DecSlice* __restrict__ ds, u32 sz, const u8* __restrict__ stream,
                                                       const u32* __restrict__ fields, int order, const u32* __restrict__ ncpIn,
                                                       u32* __restrict__ Mw, u32* __restrict__ Sw, int2* __restrict__ Qw,
                                                       u32* __restrict__ nevOut, ull* scal) {
  __shared__ u32 sm[33];
  for (u32 z = blockIdx.x; z < sz; z += gridDim.x) {
    const DecSlice d = ds[z];
    const u64 ncp = order > 0 ? (u64)ncpIn[z] : (u64)d.blen * 4;
    const u64 nwords = (ncp + 15) / 16;
    u32 carry = 0, count = 0;
    int qx = 0, qy = 0;                                       // displacement of all non-event codepoints before the chunk
    for (u64 w0 = 0; w0 < nwords; w0 += blockDim.x) {
      const u64 wi = w0 + threadIdx.x;
      const bool in = wi < nwords;
      const u32 w = in ? dec_word(d, stream, fields, order, wi) : 0u;
      const u32 incl = word_prefix4(w);
      u32 tot;
      const u32 ex = block_excl_scan(incl >> 30, sm, tot);
      const u32 e = (carry + ex) & 3u;                        // absolute move of the codepoint before this word
      carry = (carry + tot) & 3u;
      u32 M = 0, S = 0;
      int disp = 0;
      if (in) {
        M = add4(incl, e * L5);
        u32 Oc = opp_flags(M, e);
        if (wi == 0) Oc &= ~1u;                               // the first codepoint has no predecessor
        const u64 rem = ncp - wi * 16;
        if (rem < 16) Oc &= (1u << (2 * (u32)rem)) - 1u;
        u32 Op = 0;
        if (wi > 0) {
          const u32 inclp = word_prefix4(dec_word(d, stream, fields, order, wi - 1));
          const u32 ep = (e - (inclp >> 30)) & 3u;
          Op = opp_flags(add4(inclp, ep * L5), ep);
          if (wi == 1) Op &= ~1u;
        }
        if (Op == L5 && (Oc & 1u)) atomicExch(&scal[SC_FIRST], 1ull);    // an opposite-run longer than a word: serial fallback
        // second-of-pair mask: positions at even distance from the start of their run of `opp` flags
        u64 comb = (u64)Op | ((u64)Oc << 32);
        comb |= comb << 1;
        const u64 st = comb & ~(comb << 2) & 0x5555555555555555ull;
        const u64 t = comb + (st & 0x1111111111111111ull);
        const u64 evr = comb & ~t, odr = comb & ~evr;
        S = (u32)(((evr & 0x3333333333333333ull) | (odr & 0xCCCCCCCCCCCCCCCCull)) >> 32) & L5;
        Mw[d.wordOff + wi] = M;
        Sw[d.wordOff + wi] = S;
        count += __popc(S);
        disp = word_disp(M, S, rem < 16 ? (u32)rem : 16u);
      }
      u32 dtot;
      const int2 dex = unpack_disp((int)block_excl_scan((u32)disp, sm, dtot));
      if (in) Qw[d.wordOff + wi] = make_int2(qx + dex.x, qy + dex.y);
      const int2 dt = unpack_disp((int)dtot);
      qx += dt.x; qy += dt.y;
    }
    u32 total;
    block_excl_scan(count, sm, total);
    if (threadIdx.x == 0) nevOut[z] = total;
  }
}

// event list (codepoint index | move << 30) and, per word, the number of events before it
__global__ void __launch_bounds__(256) k_dec_compact(const DecSlice* __restrict__ ds, u32 sz, int order, const u32* __restrict__ ncpIn,
                                                      const u32* __restrict__ Mw, const u32* __restrict__ Sw,
                                                      const u64* __restrict__ evOff, u32* __restrict__ evIdx, u32* __restrict__ evBaseW) {
  __shared__ u32 sm[33];
  for (u32 z = blockIdx.x; z < sz; z += gridDim.x) {
    const DecSlice d = ds[z];
    const u64 ncp = order > 0 ? (u64)ncpIn[z] : (u64)d.blen * 4;
    const u64 nwords = (ncp + 15) / 16;
    u32* out = evIdx + evOff[z];
    u32 carry = 0;
    for (u64 w0 = 0; w0 < nwords; w0 += blockDim.x) {
      const u64 wi = w0 + threadIdx.x;
      u32 S = wi < nwords ? Sw[d.wordOff + wi] : 0u;
      u32 tot;
      u32 o = carry + block_excl_scan(__popc(S), sm, tot);
      carry += tot;
      if (wi < nwords) evBaseW[d.wordOff + wi] = o;
      if (S) {
        const u32 M = Mw[d.wordOff + wi];
        while (S) {
          const u32 b = __ffs(S) - 1;                         // even bit position = 2 * field
          S &= S - 1;
          out[o++] = (u32)(wi * 16 + (b >> 1)) | (((M >> b) & 3u) << 30);
        }
      }
    }
  }
}

// Q(i): displacement of all non-event codepoints before codepoint i of a slice
__device__ __forceinline__ int2 disp_before(const u32* __restrict__ Mz, const u32* __restrict__ Sz, const int2* __restrict__ Qz, u32 i) {
  const u32 w = i >> 4, k = i & 15;
  int2 q = Qz[w];
  if (k) { const int2 p = unpack_disp(word_disp(Mz[w], Sz[w], k)); q.x += p.x; q.y += p.y; }
  return q;
}

// serial over the events of a slice: lanes load 32 events and their segment displacements, lane 0 runs the chain /
// revisit-stack logic.  Outputs per event: start position of its segment and Q at the segment's first codepoint.
// per event (fully parallel): displacement of the segment that ends at the event (written to segStart, which the chain
// pass below reads and then overwrites with the segment's start position) and Q at the segment's first codepoint
__global__ void __launch_bounds__(256) k_dec_evsum(const DecSlice* __restrict__ ds, u32 sz, const u32* __restrict__ Mw, const u32* __restrict__ Sw,
                                                    const int2* __restrict__ Qw, const u64* __restrict__ evOff, const u32* __restrict__ evIdx,
                                                    int2* __restrict__ segStart, int2* __restrict__ segQ, int2* __restrict__ segSum) {
  for (u32 z = blockIdx.y; z < sz; z += gridDim.y) {
    const DecSlice d = ds[z];
    const u64 e0 = evOff[z];
    const u32 nev = (u32)(evOff[z + 1] - e0);
    const u32* Mz = Mw + d.wordOff;
    const u32* Sz = Sw + d.wordOff;
    const int2* Qz = Qw + d.wordOff;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < nev; j += gridDim.x * blockDim.x) {
      const u32 ei = evIdx[e0 + j] & 0x3FFFFFFFu;
      const u32 first = j ? (evIdx[e0 + j - 1] & 0x3FFFFFFFu) + 1 : 0u;
      const int2 qf = disp_before(Mz, Sz, Qz, first);
      const int2 ql = disp_before(Mz, Sz, Qz, ei - 1);       // the codepoint before an event is the dropped first-of-pair
      const int2 sum = make_int2(ql.x - qf.x, ql.y - qf.y);
      segSum[e0 + j] = sum;
      segQ[e0 + j] = qf;
    }
  }
}

#define DEC_STACK 512
__device__ __forceinline__ u32 dec_smem_addr(const void* p) {
  u32 a = (u32)__cvta_generic_to_shared(p);
  asm volatile("" : "+r"(a));          // opaque: keep the address in a register instead of rematerialising it
  return a;
}
__device__ __forceinline__ int2 dec_lds64(u32 a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void dec_sts64(u32 a, int2 v) { asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
// Serial form of the chain pass: the fallback for slices the warp-parallel kernel below hands back (`redo[z]`; the segment
// sums it needs are kept in segSum).
__global__ void __launch_bounds__(32) k_dec_chain_serial(Geom g, const DecSlice* __restrict__ ds, const u8* __restrict__ stream,
                                                   const u32* __restrict__ Mw, const u32* __restrict__ Sw, const int2* __restrict__ Qw,
                                                   const u64* __restrict__ evOff, const u32* __restrict__ evIdx,
                                                   int2* __restrict__ segStart, int2* __restrict__ segQ, int2* __restrict__ gstack,
                                                   u32* __restrict__ nevUsed, const u32* __restrict__ redo, const int2* __restrict__ segSum,
                                                   ull* scal) {
  __shared__ int2 sstack[DEC_STACK];
  __shared__ int2 s_sum[32], s_start[32];
  __shared__ u32 s_used;
  const u32 z = blockIdx.x, lane = threadIdx.x;
  if (!redo[z]) return;
  const DecSlice d = ds[z];
  const u64 e0 = evOff[z];
  const u32 nev = (u32)(evOff[z + 1] - e0);
  const u32* Mz = Mw + d.wordOff;
  const u32* Sz = Sw + d.wordOff;
  const int2* Qz = Qw + d.wordOff;
  const int xw = ckl_byte_width((u64)g.sx + 1), yw = ckl_byte_width((u64)g.sy + 1);
  // lane 0 state
  BocIter it;
  int x = 0, y = 0;
  u32 open = 0, sp = 0, used = nev;
  bool stop = false;
  if (lane == 0) {
    const u8* code = stream + d.code;
    it.p = code; it.idx = 4; it.end = d.isz; it.xw = xw; it.yw = yw; it.sxe = g.sx + 1;
    it.ny = d.isz >= (u32)(4 + yw) ? (u32)ld_le(code + 4, yw) : 0u; it.idx += yw; it.yi = 0; it.nx = 0; it.y = 0; it.x = 0;
    s_used = nev;
  }
  int2* gst = gstack + e0;
  // shared memory through explicit 32-bit shared-space addresses (the serial loop below otherwise re-derives the shared
  // window base on every access); event types travel as a ballot mask in a register
  const u32 a_sum = dec_smem_addr(s_sum), a_start = dec_smem_addr(s_start), a_stack = dec_smem_addr(sstack);
  // the next batch's inputs are loaded before the serial pass over the current one (one coalesced load each)
  int2 nsum = make_int2(0, 0);
  u32 nev_w = 0;
  if (lane < nev) { nsum = segSum[e0 + lane]; nev_w = evIdx[e0 + lane]; }
  for (u32 base = 0; base < nev; base += 32) {
    const u32 j = base + lane;
    const int2 csum = nsum;
    const u32 m = nev_w >> 30;
    const bool is_t = j < nev && (m == 0 || m == 3);
    if (j + 32 < nev) { nsum = segSum[e0 + j + 32]; nev_w = evIdx[e0 + j + 32]; }
    s_sum[lane] = csum;
    const u32 tmask = __ballot_sync(FULL_MASK, is_t);
    __syncwarp();
    if (lane == 0 && !stop) {
      const u32 n = min(32u, nev - base);
#pragma unroll 4
      for (u32 k = 0; k < n; k++) {
        if (__builtin_expect(open == 0, 0)) {            // next chain: start vertex from the BOC index
          u32 vx, vy;
          if (!it.next(vx, vy)) { used = base + k; stop = true; break; }
          if (vx > g.sx || vy > g.sy) { atomicExch(&scal[SC_ERROR], 11ull); used = base + k; stop = true; break; }
          x = (int)vx; y = (int)vy; sp = 0; open = 1;
        }
        dec_sts64(a_start + k * 8u, make_int2(x, y));
        const int2 sm = dec_lds64(a_sum + k * 8u);
        x += sm.x; y += sm.y;
        if ((tmask >> k) & 1u) {                          // 't'
          open--;
          if (sp > 0) {
            --sp;
            const int2 q = sp < DEC_STACK ? dec_lds64(a_stack + sp * 8u) : gst[sp - DEC_STACK];
            x = q.x; y = q.y;
          }
        } else {                                          // 'b': reference quirk -- pushed as x + sx*y (crackcodes.hpp:772,850)
          open++;
          const int2 q = x == (int)g.sx ? make_int2(0, y + 1) : make_int2(x, y);
          if (sp < DEC_STACK) dec_sts64(a_stack + sp * 8u, q); else gst[sp - DEC_STACK] = q;
          sp++;
        }
      }
      if (stop) s_used = used;
    }
    __syncwarp();
    if (j < nev && j < s_used) segStart[e0 + j] = s_start[lane];
    __syncwarp();
  }
  if (lane == 0) nevUsed[z] = s_used;
}

// Warp-parallel chain pass.  The serial dependency of the pass is the revisit stack: after a 't' the position is the one
// its matching 'b' pushed.  32 events at a time: bracket matching inside the batch (the match of a 't' is the event after the
// nearest earlier event whose depth is not above the depth after the 't'; deeper pops come from the stack the batch was
// entered with), which makes every event's position "the position after some earlier event (+ its own segment)" -- a
// forest of depth <= 32, resolved with five rounds of pointer jumping through shuffles.  'b's left open at the end of the
// batch go onto the stack.  A chain ends at the 't' that finds the stack empty; the batch is split there and the next start
// vertex read from the beginning-of-chain index.  The reference's push quirk (a position with x == sx is stored as (0, y+1),
// crackcodes.hpp:772,850) is not linear: a slice where it occurs is handed to the serial kernel (`redo`).
#define DCH_SMEM 3072           // stack entries held in shared memory (deeper entries live in global memory)
#define DCH_ROOT 32u
#define DCH_ABS 33u
__global__ void __launch_bounds__(32) k_dec_chain(Geom g, const DecSlice* __restrict__ ds, const u8* __restrict__ stream,
                                                   const u64* __restrict__ evOff, const u32* __restrict__ evIdx,
                                                   int2* __restrict__ segStart, const int2* __restrict__ segSum, int2* __restrict__ gstack,
                                                   u32* __restrict__ nevUsed, u32* __restrict__ redo, ull* scal, const bool force_serial) {
  __shared__ int2 sstack[DCH_SMEM];
  const u32 z = blockIdx.x, lane = threadIdx.x;
  const DecSlice d = ds[z];
  const u64 e0 = evOff[z];
  const u32 nev = (u32)(evOff[z + 1] - e0);
  const int xw = ckl_byte_width((u64)g.sx + 1), yw = ckl_byte_width((u64)g.sy + 1);
  BocIter it;
  {
    const u8* code = stream + d.code;
    it.p = code; it.idx = 4; it.end = d.isz; it.xw = xw; it.yw = yw; it.sxe = g.sx + 1;
    it.ny = d.isz >= (u32)(4 + yw) ? (u32)ld_le(code + 4, yw) : 0u; it.idx += yw; it.yi = 0; it.nx = 0; it.y = 0; it.x = 0;
  }
  int2* gst = gstack + e0;
  int curx = 0, cury = 0;
  u32 sp = 0, used = nev;
  bool open = false, quirk = false;
  for (u32 base = 0; base < nev && used == nev; base += 32) {
    const u32 j = base + lane;
    const bool valid = j < nev;
    const u32 m = valid ? evIdx[e0 + j] >> 30 : 1u;
    const bool isT = valid && (m == 0 || m == 3), isB = valid && !isT;
    const int2 sum = valid ? segSum[e0 + j] : make_int2(0, 0);
    u32 lo = 0;
    while (lo < 32 && base + lo < nev) {
      if (!open) {                                         // next chain: start vertex from the BOC index (lane 0 reads it)
        u32 vx = 0, vy = 0, ok = 0;
        if (lane == 0) ok = it.next(vx, vy) ? 1u : 0u;
        ok = __shfl_sync(FULL_MASK, ok, 0); vx = __shfl_sync(FULL_MASK, vx, 0); vy = __shfl_sync(FULL_MASK, vy, 0);
        if (!ok || vx > g.sx || vy > g.sy) {
          if (ok && lane == 0) atomicExch(&scal[SC_ERROR], 11ull);
          used = base + lo;
          break;
        }
        curx = (int)vx; cury = (int)vy; sp = 0; open = true;
      }
      const bool act = valid && lane >= lo;
      const u32 bm = __ballot_sync(FULL_MASK, isB && act), tm = __ballot_sync(FULL_MASK, isT && act);
      const u32 upto = ((2u << lane) - 1u) & ~((1u << lo) - 1u);
      const int v = (int)__popc(bm & upto) - (int)__popc(tm & upto);          // depth after this event, relative to the batch entry
      const u32 endm = __ballot_sync(FULL_MASK, act && isT && (int)sp + v < 0); // a 't' that finds the stack empty ends the chain
      const u32 hi = endm ? (u32)__ffs(endm) - 1u : 31u;
      const bool on = act && lane <= hi;
      // nearest earlier active event whose depth is <= this one's (brute force over the 31 distances)
      int found = -1;
#pragma unroll
      for (u32 step = 1; step < 32; step++) {
        const int c = __shfl_up_sync(FULL_MASK, v, step);
        if (found < 0 && lane >= lo + step && c <= v) found = (int)(lane - step);
      }
      u32 par = DCH_ABS;
      int offx = 0, offy = 0;
      u32 mlane = 32;                                      // the in-batch 'b' a 't' returns to
      if (on && isB) { par = lane == lo ? DCH_ROOT : lane - 1; offx = sum.x; offy = sum.y; }
      else if (on && isT) {
        if (found >= 0) mlane = (u32)found + 1u;
        else if (v >= 0) mlane = lo;
        if (mlane < 32) par = mlane;
        else if ((int)sp + v >= 0) {                       // from the stack the batch was entered with
          const u32 si = sp + (u32)v;
          const int2 q = si < DCH_SMEM ? sstack[si] : gst[si - DCH_SMEM];
          offx = q.x; offy = q.y;
        }
      }
      if (on && isT && endm && lane == hi) { par = DCH_ABS; mlane = 32; }       // the chain-ending 't': its position is not used
      const u32 matched = __reduce_or_sync(FULL_MASK, mlane < 32 ? 1u << mlane : 0u);
#pragma unroll
      for (int r = 0; r < 5; r++) {
        const u32 src = par & 31u;
        const int px = __shfl_sync(FULL_MASK, offx, src), py = __shfl_sync(FULL_MASK, offy, src);
        const u32 pp = __shfl_sync(FULL_MASK, par, src);
        if (par < 32) { offx += px; offy += py; par = pp; }
      }
      const int ax = (par == DCH_ROOT ? curx : 0) + offx, ay = (par == DCH_ROOT ? cury : 0) + offy;    // position after the event
      // the push quirk: a pushed position with x == sx would come back as (0, y + 1)
      if (__any_sync(FULL_MASK, on && isB && ax == (int)g.sx)) { quirk = true; break; }
      int bx = __shfl_up_sync(FULL_MASK, ax, 1), by = __shfl_up_sync(FULL_MASK, ay, 1);
      if (lane == lo) { bx = curx; by = cury; }
      if (on) segStart[e0 + j] = make_int2(bx, by);
      __syncwarp();                                        // stack reads above, stack writes below
      if (on && isB && !((matched >> lane) & 1u)) {        // still open at the end of the batch
        const u32 si = sp + (u32)v - 1u;
        if (si < DCH_SMEM) sstack[si] = make_int2(ax, ay); else gst[si - DCH_SMEM] = make_int2(ax, ay);
      }
      __syncwarp();
      if (endm) { open = false; sp = 0; lo = hi + 1; }
      else {
        const u32 last = min(31u, nev - 1u - base);        // last valid lane of the batch
        sp = (u32)((int)sp + __shfl_sync(FULL_MASK, v, last));
        curx = __shfl_sync(FULL_MASK, ax, last); cury = __shfl_sync(FULL_MASK, ay, last);
        lo = 32;
      }
    }
    if (quirk) break;
  }
  if (lane == 0) { nevUsed[z] = used; redo[z] = (quirk || force_serial) ? 1u : 0u; }
}

// marking: one thread per 16-codepoint word (uniform work).  Position at the word start = start of the segment the
// word begins in + (Q(word start) - Q(segment start)); an event inside the word switches to the next segment's start.
__global__ void __launch_bounds__(256) k_dec_mark(Geom g, const DecSlice* __restrict__ ds, u32 sz, int order, const u32* __restrict__ ncpIn,
                                                   const u32* __restrict__ Mw, const u32* __restrict__ Sw, const int2* __restrict__ Qw,
                                                   const u32* __restrict__ evBaseW, const u64* __restrict__ evOff,
                                                   const int2* __restrict__ segStart, const int2* __restrict__ segQ,
                                                   const u32* __restrict__ nevUsed, u32* __restrict__ EVall, u32* __restrict__ EHall,
                                                   ull* scal) {
  const u64 nw = (u64)g.sy * g.W;
  const int sx = (int)g.sx, sy = (int)g.sy;
  for (u32 z = blockIdx.y; z < sz; z += gridDim.y) {
    const DecSlice d = ds[z];
    const u64 ncp = order > 0 ? (u64)ncpIn[z] : (u64)d.blen * 4;
    const u32 nwords = (u32)((ncp + 15) / 16);
    const u32 used = nevUsed[z];
    const u64 e0 = evOff[z];
    u32* EV = EVall + (u64)z * nw;
    u32* EH = EHall + (u64)z * nw;
    for (u32 w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
      u32 s = evBaseW[d.wordOff + w];
      if (s >= used) continue;                              // codepoints after the last chain (padding)
      const u32 M = Mw[d.wordOff + w], S = Sw[d.wordOff + w];
      const u32 Snext = w + 1 < nwords ? Sw[d.wordOff + w + 1] : 0u;
      const u32 drop = (S >> 2) | ((Snext & 1u) << 30);     // first-of-pair fields: the field before an event
      const u32 nf = (u32)min((u64)16, ncp - (u64)w * 16);
      const int2 ss = segStart[e0 + s], sq = segQ[e0 + s], q = Qw[d.wordOff + w];
      int x = ss.x + q.x - sq.x, y = ss.y + q.y - sq.y;
      bool bad = false;
      u32* pend = nullptr;
      u32 pmask = 0;
      for (u32 k = 0; k < nf; k++) {
        const u32 fb = 1u << (2 * k);
        if (S & fb) {                                       // event: the next codepoint starts the next segment
          if (++s >= used) break;
          const int2 p = segStart[e0 + s];
          x = p.x; y = p.y;
          continue;
        }
        if (drop & fb) continue;
        // one branch-free body for the four moves (0 up, 1 right, 2 down, 3 left), so the lanes of a warp stay together:
        // vertical moves mark EV at column x (row y-1 going up, y going down), horizontal moves mark EH at row y
        // (column x-1 going left, x going right); cracks on the image border are not stored
        const u32 m = (M >> (2 * k)) & 3u;
        const u32 isH = m & 1u, neg = (0x9u >> m) & 1u;
        const int step = neg ? -1 : 1;
        const int nx = x + (isH ? step : 0), ny = y + (isH ? 0 : step);
        if ((u32)nx > (u32)sx || (u32)ny > (u32)sy || (u32)x > (u32)sx || (u32)y > (u32)sy) { bad = true; break; }
        const int col = x - (int)(isH & neg), row = y - (int)((isH ^ 1u) & neg);
        const bool interior = isH ? (y > 0 && y < sy) : (x > 0 && x < sx);
        if (interior) {
          u32* a = (isH ? EH : EV) + (u64)row * g.W + (col >> 5);
          if (a != pend) { if (pend) atomicOr(pend, pmask); pend = a; pmask = 0; }
          pmask |= 1u << (col & 31);
        }
        x = nx; y = ny;
      }
      if (pend) atomicOr(pend, pmask);
      if (bad) atomicExch(&scal[SC_ERROR], 11ull);
    }
  }
}

static u32 dec_grid(u64 n, u32 bs, u32 per_sm) { return ckl_grid(n, bs, per_sm, false); }

// per-slice descriptors from the code offsets: index size read from the stream (crackcodes.hpp:283-316)
__global__ void k_dec_slices_init(Geom g, const u8* __restrict__ stream, const u64* __restrict__ codeOff, const u64* __restrict__ wordOff,
                                  DecSlice* __restrict__ out, ull* scal) {
  const u32 z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= g.sz) return;
  const u64 yw = (u64)ckl_byte_width((u64)g.sy + 1);
  DecSlice d;
  d.code = codeOff[z];
  d.wordOff = wordOff[z];
  const u64 clen = codeOff[z + 1] - codeOff[z];
  d.isz = 0; d.body = d.code; d.blen = 0;
  if (clen >= 4 + yw) {
    const u64 isz = 4 + ld_le(stream + d.code, 4);
    if (isz > clen) atomicExch(&scal[SC_ERROR], 10ull);
    else { d.isz = (u32)isz; d.body = d.code + isz; d.blen = (u32)(clen - isz); }
  }
  out[z] = d;
}
void launch_decode_slices_init(const Geom& g, const u8* stream, const u64* codeOff, const u64* wordOff, DecSlice* out, ull* scal,
                               cudaStream_t st) {
  k_dec_slices_init<<<(g.sz + 127) / 128, 128, 0, st>>>(g, stream, codeOff, wordOff, out, scal);
  LAUNCH_CHECK();
}

// phase 1: moves + event masks + displacement prefixes + per-slice event counts (exclusive scan into D.evOff, total
// in scal[SC_LAST])
void launch_decode_classify(const Geom& g, const u8* stream, int order, const u8* model, DecodeBufs& D, u64 total_words, ull* scal,
                            cudaStream_t st) {
  D.Mw.ensure(total_words * 4 + 16);
  D.Sw.ensure(total_words * 4 + 16);
  D.Qw.ensure(total_words * 8 + 16);
  D.evBaseW.ensure(total_words * 4 + 16);
  D.nev.ensure((u64)g.sz * 4);
  D.ncp.ensure((u64)g.sz * 4);
  D.evOff.ensure(((u64)g.sz + 1) * 8);
  D.nevUsed.ensure((u64)g.sz * 4);
  const DecSlice* ds = D.slices.as<DecSlice>();
  if (order > 0) {
    D.fields.ensure(total_words * 4 + 16);
    // evBaseW is free until k_dec_compact: it holds the per-word symbol numbering of the bitstreams meanwhile
    CUDA_CHECK(cudaMemsetAsync(D.fields.p, 0, total_words * 4, st));
    k_mk_scan<<<dec_grid(g.sz, 1, 8), 256, 0, st>>>(ds, g.sz, stream, D.evBaseW.as<u32>(), D.ncp.as<u32>());
    LAUNCH_CHECK();
    k_mk_decode<<<dec_grid(g.sz, 1, 8), 256, 0, st>>>(ds, g.sz, stream, order, model, D.evBaseW.as<u32>(), D.fields.as<u32>());
    LAUNCH_CHECK();
  }
  k_dec_classify<<<dec_grid(g.sz, 1, 8), 256, 0, st>>>(ds, g.sz, stream, D.fields.as<u32>(), order, D.ncp.as<u32>(), D.Mw.as<u32>(),
                                                       D.Sw.as<u32>(), D.Qw.as<int2>(), D.nev.as<u32>(), scal);
  LAUNCH_CHECK();
  launch_exscan_u32_u64(D.nev.as<u32>(), g.sz, 1, D.evOff.as<u64>(), &scal[SC_LAST], 0, st);
}

// phase 2 (total_events known on the host): event list, chain pass, marking
void launch_decode_mark(const Geom& g, const u8* stream, int order, DecodeBufs& D, u64 total_events, u64 total_words, u32* EV, u32* EH,
                        ull* scal, cudaStream_t st) {
  CUDA_CHECK(cudaMemsetAsync(EV, 0, g.words() * 4, st));
  CUDA_CHECK(cudaMemsetAsync(EH, 0, g.words() * 4, st));
  if (!total_events) return;
  D.evIdx.ensure(total_events * 4 + 16);
  D.segQ.ensure(total_events * 8 + 16);
  D.segStart.ensure(total_events * 8 + 16);
  D.gstack.ensure(total_events * 8 + 16);
  D.segSum.ensure(total_events * 8 + 16);
  D.redo.ensure((u64)g.sz * 4 + 16);
  const DecSlice* ds = D.slices.as<DecSlice>();
  k_dec_compact<<<dec_grid(g.sz, 1, 8), 256, 0, st>>>(ds, g.sz, order, D.ncp.as<u32>(), D.Mw.as<u32>(), D.Sw.as<u32>(), D.evOff.as<u64>(),
                                                      D.evIdx.as<u32>(), D.evBaseW.as<u32>());
  LAUNCH_CHECK();
  {
    const u64 per_slice = (total_events + g.sz - 1) / g.sz;
    u32 gx = (u32)((per_slice + 255) / 256);
    if (gx > 32) gx = 32;
    if (gx < 1) gx = 1;
    k_dec_evsum<<<dim3(gx, g.sz < 65535u ? g.sz : 65535u), 256, 0, st>>>(ds, g.sz, D.Mw.as<u32>(), D.Sw.as<u32>(), D.Qw.as<int2>(),
                                                                          D.evOff.as<u64>(), D.evIdx.as<u32>(), D.segStart.as<int2>(),
                                                                          D.segQ.as<int2>(), D.segSum.as<int2>());
    LAUNCH_CHECK();
  }
  // k_dec_evsum left the segment sums in segSum; the chain pass turns them into segment start positions
  static int force_env = -1;      // test hook: every slice takes the serial fallback (tests/test_gpu_parity.py)
  if (force_env < 0) { const char* e = getenv("CKL_TEST_SERIAL_CHAIN"); force_env = (e && atoi(e) > 0) ? 1 : 0; }
  const bool force_serial = force_env == 1;
  k_dec_chain<<<g.sz, 32, 0, st>>>(g, ds, stream, D.evOff.as<u64>(), D.evIdx.as<u32>(), D.segStart.as<int2>(), D.segSum.as<int2>(),
                                   D.gstack.as<int2>(), D.nevUsed.as<u32>(), D.redo.as<u32>(), scal, force_serial);
  LAUNCH_CHECK();
  k_dec_chain_serial<<<g.sz, 32, 0, st>>>(g, ds, stream, D.Mw.as<u32>(), D.Sw.as<u32>(), D.Qw.as<int2>(), D.evOff.as<u64>(), D.evIdx.as<u32>(),
                                          D.segStart.as<int2>(), D.segQ.as<int2>(), D.gstack.as<int2>(), D.nevUsed.as<u32>(), D.redo.as<u32>(),
                                          D.segSum.as<int2>(), scal);
  LAUNCH_CHECK();
  const u64 per_slice = (total_words + g.sz - 1) / g.sz;
  u32 gx = (u32)((per_slice + 255) / 256);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  k_dec_mark<<<dim3(gx, g.sz < 65535u ? g.sz : 65535u), 256, 0, st>>>(g, ds, g.sz, order, D.ncp.as<u32>(), D.Mw.as<u32>(), D.Sw.as<u32>(),
                                                                       D.Qw.as<int2>(), D.evBaseW.as<u32>(), D.evOff.as<u64>(),
                                                                       D.segStart.as<int2>(), D.segQ.as<int2>(), D.nevUsed.as<u32>(),
                                                                       EV, EH, scal);
  LAUNCH_CHECK();
}

// crack planes -> "differ" planes.  IMPERMISSIBLE: identical.  PERMISSIBLE: cracks mark connected neighbours.
__global__ void __launch_bounds__(256) k_planes_from_cracks(Geom g, u32* EV, u32* EH) {
  const u64 nwords = g.words(), stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
    const u64 row = fdiv(i, g.W);
    const u32 w = (u32)(i - row * g.W);
    const u32 y = (u32)(row - fdiv(row, g.sy) * g.sy);
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

// paint: one warp per 32-pixel word, lane <-> pixel; the only full-width write of decompress.
template <typename OUT, bool MASK, bool FORTRAN>
__global__ void __launch_bounds__(256) k_paint(Geom g, const u32* __restrict__ DV, const u32* __restrict__ wordPrefix,
                                                const u32* __restrict__ rowBase, const u64* __restrict__ runBase,
                                                const u64* __restrict__ runLabel, u64 label, OUT* __restrict__ out) {
  const u32 lane = threadIdx.x & 31;
  const u64 nwords = g.words();
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  for (u64 i = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nwords; i += nwarps) {
    const u64 row = fdiv(i, g.W);
    const u32 w = (u32)(i - row * g.W);
    const u32 z = (u32)fdiv(row, g.sy), y = (u32)(row - (u64)z * g.sy);
    const u32 x = w * 32 + lane;
    if (x >= g.sx) continue;
    const u32 dv = DV[i];
    const u32 rid = rowBase[row] + wordPrefix[i] + __popc(dv & ((2u << lane) - 1u));
    const u64 v = runLabel[runBase[z] + rid];
    const u64 o = FORTRAN ? ((u64)z * g.sxy + (u64)y * g.sx + x) : ((u64)z + (u64)g.sz * ((u64)y + (u64)g.sy * x));
    out[o] = MASK ? (OUT)(v == label) : (OUT)v;
  }
}

// Fortran-order paint: one warp per row walks the row 32 pixels at a time; the run index inside the row is a
// running popcount of the DV words (no per-word prefix array), plane words are loaded 32 at a time and broadcast.
template <typename OUT, bool MASK>
__global__ void __launch_bounds__(256) k_paint_rows(Geom g, const u32* __restrict__ DV, const u32* __restrict__ rowBase,
                                                     const u64* __restrict__ runBase, const u64* __restrict__ runLabel, u64 label,
                                                     OUT* __restrict__ out, const bool STREAM) {
  const u32 lane = threadIdx.x & 31;
  const u32 lmask = (2u << lane) - 1u;
  const u64 rows = g.rows();
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  for (u64 row = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += nwarps) {
    const u32 z = (u32)fdiv(row, g.sy);
    const u64* rl = runLabel + runBase[z] + rowBase[row];
    OUT* o = out + row * g.sx;
    const u32* dvrow = DV + row * g.W;
    u32 carry = 0;
    for (u32 w0 = 0; w0 < g.W; w0 += 32) {
      const u32 dvl = w0 + lane < g.W ? dvrow[w0 + lane] : 0u;
      const u32 nk = min(32u, g.W - w0);
      for (u32 k = 0; k < nk; k++) {
        const u32 dv = __shfl_sync(FULL_MASK, dvl, k);
        const u32 x = (w0 + k) * 32 + lane;
        if (x < g.sx) {
          const u64 v = rl[carry + __popc(dv & lmask)];
          const OUT r = MASK ? (OUT)(v == label) : (OUT)v;
          if (STREAM) __stcs(o + x, r); else o[x] = r;    // streaming store: the output is written once and not re-read here
        }
        carry += __popc(dv);
      }
    }
  }
}

// TMA-staged Fortran-order paint (sx a multiple of 32 * WORDS): a warp owns a band of 32 * WORDS pixels and walks down a
// range of rows.  The run index of a pixel is rowBase + wordPrefix (both from the run numbering of the CCL) + a popcount of
// its word's DV bits, so the WORDS label gathers of a row are independent; labels go lane <-> pixel into a shared-memory
// stage (conflict-free st.shared) and every finished row of the band leaves as ONE bulk store (cp.async.bulk shared ->
// global, up to 2 KB), a few rows in flight per warp -- no lane computes a global store address, and the writes reach
// memory as whole contiguous rows.
#define PT_WARPS 4
#define PT_ROWS 128
template <typename OUT, bool MASK, int WORDS, int PT_STAGES>
__global__ void __launch_bounds__(32 * PT_WARPS) k_paint_tma(Geom g, const u32* __restrict__ DV, const u32* __restrict__ wordPrefix,
                                                              const u32* __restrict__ rowBase, const u64* __restrict__ runBase,
                                                              const u64* __restrict__ runLabel, u64 label, OUT* __restrict__ out) {
  constexpr u32 ROWB = 32 * WORDS * sizeof(OUT);
  extern __shared__ __align__(128) u8 pt_smem[];
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const u32 lmask = (2u << lane) - 1u;
  const u32 ring = (u32)__cvta_generic_to_shared(pt_smem) + wib * (PT_STAGES * ROWB);
  const u64 nwarps = (u64)gridDim.x * PT_WARPS;
  const u32 nband = g.sx / (32 * WORDS), nrange = (g.sy + PT_ROWS - 1) / PT_ROWS;
  const u64 items = (u64)g.sz * nrange * nband;
  u32 n = 0;
  for (u64 it = (u64)blockIdx.x * PT_WARPS + wib; it < items; it += nwarps) {
    const u32 band = (u32)(it % nband);
    const u64 t2 = it / nband;
    const u32 range = (u32)(t2 % nrange), z = (u32)(t2 / nrange);
    const u32 y0 = range * PT_ROWS, y1 = min(g.sy, y0 + PT_ROWS);
    const u32 w0 = band * WORDS;
    const u64* rlz = runLabel + runBase[z];
    u64 row = (u64)z * g.sy + y0;
    u32 dvl = lane < WORDS ? DV[row * g.W + w0 + lane] : 0u;
    u32 wpl = lane < WORDS ? wordPrefix[row * g.W + w0 + lane] : 0u;
    u32 rb = rowBase[row];
    for (u32 y = y0; y < y1; y++, n++, row++) {
      const u32 cdv = dvl, cwp = wpl, crb = rb;
      if (y + 1 < y1) {                                    // the next row's plane words are in flight while this row is painted
        if (lane < WORDS) { dvl = DV[(row + 1) * g.W + w0 + lane]; wpl = wordPrefix[(row + 1) * g.W + w0 + lane]; }
        rb = rowBase[row + 1];
      }
      const u32 stage = ring + (n % PT_STAGES) * ROWB;
      // the bulk store that read this stage PT_STAGES rows ago must have finished reading it
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PT_STAGES - 1) : "memory");
      __syncwarp();
      u64 v[WORDS];
#pragma unroll
      for (int j = 0; j < WORDS; j++) {
        const u32 dv = __shfl_sync(FULL_MASK, cdv, j);
        const u32 pre = __shfl_sync(FULL_MASK, cwp, j);
        v[j] = __ldg(rlz + crb + pre + __popc(dv & lmask));
      }
#pragma unroll
      for (int j = 0; j < WORDS; j++) {
        const u32 a = stage + (j * 32 + lane) * (u32)sizeof(OUT);
        if constexpr (MASK) { const u32 b = v[j] == label ? 1u : 0u; asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
        else if constexpr (sizeof(OUT) == 8) asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"((ull)v[j]) : "memory");
        else if constexpr (sizeof(OUT) == 4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"((u32)v[j]) : "memory");
        else if constexpr (sizeof(OUT) == 2) asm volatile("st.shared.u16 [%0], %1;" ::"h"((u16)v[j]), "r"(a) : "memory");
        else { const u32 b = (u32)v[j] & 0xFFu; asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk-copy engine
      __syncwarp();
      if (lane == 0) {
        OUT* dst = out + row * g.sx + (u64)w0 * 32;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(stage), "r"(ROWB) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // shared memory must outlive the last stores
}
template <typename OUT, bool MASK, int WORDS, int PT_STAGES>
static void paint_tma(const Geom& g, const u32* DV, const CclBufs& B, const u64* runLabel, u64 label, void* out, cudaStream_t st) {
  const size_t smem = (size_t)PT_WARPS * PT_STAGES * 32 * WORDS * sizeof(OUT);
  const u64 items = (u64)g.sz * ((g.sy + PT_ROWS - 1) / PT_ROWS) * (g.sx / (32 * WORDS));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const u64 cap = (u64)sms * std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 1024))) * (u64)g_ckl_grid_mult;
  const u64 need = (items + PT_WARPS - 1) / PT_WARPS;
  if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_paint_tma<OUT, MASK, WORDS, PT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_paint_tma<OUT, MASK, WORDS, PT_STAGES><<<(u32)std::max<u64>(1, std::min(need, cap)), 32 * PT_WARPS, smem, st>>>(
      g, DV, B.wordPrefix.as<u32>(), B.rowBase.as<u32>(), B.runBase.as<u64>(), runLabel, label, (OUT*)out);
  LAUNCH_CHECK();
}

// C-order paint (out[z + sz * (y + sy * x)], crackle.hpp:649-654) as a tiled transpose: a block takes a tile of ZT slices x 32
// pixels of one row y.  Reading is lane <-> x (one warp per slice: the run-label gather is as coalesced as in the Fortran
// paint), writing is lane <-> z through a padded shared-memory tile, so every store instruction writes 32 consecutive z of
// one (x, y) -- 32 * sizeof(OUT) contiguous bytes instead of one element per lane every sz * sy * sizeof(OUT) bytes.
template <typename OUT, bool MASK, int ZT>
__global__ void __launch_bounds__(256) k_paint_c_tiled(Geom g, const u32* __restrict__ DV, const u32* __restrict__ wordPrefix,
                                                        const u32* __restrict__ rowBase, const u64* __restrict__ runBase,
                                                        const u64* __restrict__ runLabel, u64 label, OUT* __restrict__ out) {
  __shared__ OUT tile[ZT][33];
  const u32 tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 8 warps
  const u32 ztiles = (g.sz + ZT - 1) / ZT;
  const u64 ntiles = (u64)g.W * g.sy * ztiles;
  for (u64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const u32 zt = (u32)(t % ztiles);
    const u64 t2 = t / ztiles;
    const u32 y = (u32)(t2 % g.sy), w = (u32)(t2 / g.sy);
    const u32 z0 = zt * ZT, x = w * 32 + tx;
#pragma unroll 4
    for (u32 zi = ty; zi < ZT; zi += 8) {
      const u32 z = z0 + zi;
      OUT r = 0;
      if (z < g.sz && x < g.sx) {
        const u64 row = (u64)z * g.sy + y;
        const u64 i = row * g.W + w;
        const u32 dv = DV[i];
        const u64 v = runLabel[runBase[z] + rowBase[row] + wordPrefix[i] + __popc(dv & ((2u << tx) - 1u))];
        r = MASK ? (OUT)(v == label) : (OUT)v;
      }
      tile[zi][tx] = r;
    }
    __syncthreads();
    for (u32 xi = ty; xi < 32; xi += 8) {
      const u32 xx = w * 32 + xi;
      if (xx >= g.sx) continue;
      OUT* o = out + (u64)g.sz * ((u64)y + (u64)g.sy * xx) + z0;
#pragma unroll
      for (u32 zz = 0; zz < ZT; zz += 32)
        if (z0 + zz + tx < g.sz) __stcs(o + zz + tx, tile[zz + tx][xi]);
    }
    __syncthreads();
  }
}

// Band form with direct stores (sx a multiple of 256): the same decomposition as above without the staging -- the eight label
// gathers of a row are issued together (no running popcount along the row: wordPrefix gives every word its first run), the
// next row's plane words are already in flight, and each word is one coalesced 32-lane store.
template <typename OUT, bool MASK>
__global__ void __launch_bounds__(256) k_paint_band(Geom g, const u32* __restrict__ DV, const u32* __restrict__ wordPrefix,
                                                     const u32* __restrict__ rowBase, const u64* __restrict__ runBase,
                                                     const u64* __restrict__ runLabel, u64 label, OUT* __restrict__ out) {
  constexpr int WORDS = 8;
  const u32 lane = threadIdx.x & 31;
  const u32 lmask = (2u << lane) - 1u;
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  const u32 nband = g.sx / (32 * WORDS), nrange = (g.sy + PT_ROWS - 1) / PT_ROWS;
  const u64 items = (u64)g.sz * nrange * nband;
  for (u64 it = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < items; it += nwarps) {
    const u32 band = (u32)(it % nband);
    const u64 t2 = it / nband;
    const u32 range = (u32)(t2 % nrange), z = (u32)(t2 / nrange);
    const u32 y0 = range * PT_ROWS, y1 = min(g.sy, y0 + PT_ROWS);
    const u32 w0 = band * WORDS;
    const u64* rlz = runLabel + runBase[z];
    u64 row = (u64)z * g.sy + y0;
    u32 dvl = lane < WORDS ? DV[row * g.W + w0 + lane] : 0u;
    u32 wpl = lane < WORDS ? wordPrefix[row * g.W + w0 + lane] : 0u;
    u32 rb = rowBase[row];
    for (u32 y = y0; y < y1; y++, row++) {
      const u32 cdv = dvl, cwp = wpl, crb = rb;
      if (y + 1 < y1) {
        if (lane < WORDS) { dvl = DV[(row + 1) * g.W + w0 + lane]; wpl = wordPrefix[(row + 1) * g.W + w0 + lane]; }
        rb = rowBase[row + 1];
      }
      u64 v[WORDS];
#pragma unroll
      for (int j = 0; j < WORDS; j++) {
        const u32 dv = __shfl_sync(FULL_MASK, cdv, j);
        const u32 pre = __shfl_sync(FULL_MASK, cwp, j);
        v[j] = __ldg(rlz + crb + pre + __popc(dv & lmask));
      }
      OUT* o = out + row * g.sx + (u64)w0 * 32 + lane;
#pragma unroll
      for (int j = 0; j < WORDS; j++) __stcs(o + 32 * j, MASK ? (OUT)(v[j] == label) : (OUT)v[j]);
    }
  }
}

void launch_paint(const Geom& g, const u32* DV, const CclBufs& B, const u64* runLabel, int out_width, int has_label, u64 label,
                  int fortran_order, void* out, cudaStream_t st) {
  // Measured on B200: uint64 1024^3 1.65 ms (5.3 TB/s, 81 % of the measured peak) vs 1.71 ms for the warp-per-row form; the
  // 1-byte mask of a 2048x2048x256 volume 7.52 vs 7.99 ms per call.
  if (fortran_order && g.sx % 256 == 0 && (has_label || out_width == 8)) {
    const u64 items = (u64)g.sz * ((g.sy + PT_ROWS - 1) / PT_ROWS) * (g.sx / 256);
    const u32 gridb = grid1(items, 8, 148 * 8);
    if (has_label) k_paint_band<u8, true><<<gridb, 256, 0, st>>>(g, DV, B.wordPrefix.as<u32>(), B.rowBase.as<u32>(), B.runBase.as<u64>(), runLabel, label, (u8*)out);
    else k_paint_band<u64, false><<<gridb, 256, 0, st>>>(g, DV, B.wordPrefix.as<u32>(), B.rowBase.as<u32>(), B.runBase.as<u64>(), runLabel, label, (u64*)out);
    LAUNCH_CHECK();
    return;
  }
  // Measured on B200 (1024^3 uint64 / 2048x2048x256 uint32 + mask): the bulk-store form wins for uint32 (7.63 vs 7.97 ms per
  // decompress); for uint64 (2.04 vs 1.71 ms) and the 1-byte mask (8.17 vs 7.98 ms) it loses to the direct-store forms --
  // their 48-64 resident warps per SM hide the label gathers better than the 12-24 warps the staging buffers leave room for.
  if (fortran_order && ((u64)out & 15) == 0 && !has_label && out_width == 4) {
    if (g.sx % 512 == 0) { paint_tma<u32, false, 16, 4>(g, DV, B, runLabel, label, out, st); return; }
    if (g.sx % 256 == 0) { paint_tma<u32, false, 8, 4>(g, DV, B, runLabel, label, out, st); return; }
  }
  if (!fortran_order) {      // measured (1024x1024x512 uint64): 3.2 ms, against 24.5 ms for one scattered element per lane
    constexpr int ZT = 32;                                  // slices per tile: 128 * sizeof(OUT) contiguous bytes per (x, y)
    const u64 ntiles = (u64)g.W * g.sy * ((g.sz + ZT - 1) / ZT);
    const u32 gridc = grid1(ntiles, 1, 148 * 6);
    const u32* wpc = B.wordPrefix.as<u32>(); const u32* rbc = B.rowBase.as<u32>(); const u64* rBc = B.runBase.as<u64>();
#define PAINTC(T, M) k_paint_c_tiled<T, M, ZT><<<gridc, 256, 0, st>>>(g, DV, wpc, rbc, rBc, runLabel, label, (T*)out)
    if (has_label) PAINTC(u8, true);
    else switch (out_width) { case 1: PAINTC(u8, false); break; case 2: PAINTC(u16, false); break;
                              case 4: PAINTC(u32, false); break; default: PAINTC(u64, false); break; }
#undef PAINTC
    LAUNCH_CHECK();
    return;
  }
  const u32 grid = grid1(g.words(), 8, 148 * 8);
  const u32* wp = B.wordPrefix.as<u32>();
  const u32* rb = B.rowBase.as<u32>();
  const u64* rB = B.runBase.as<u64>();
#define PAINT(T, M, F) k_paint<T, M, F><<<grid, 256, 0, st>>>(g, DV, wp, rb, rB, runLabel, label, (T*)out)
#define PAINTR(T, M) k_paint_rows<T, M><<<grid1(g.rows(), 8, 148 * 8), 256, 0, st>>>(g, DV, rb, rB, runLabel, label, (T*)out, true)
  if (has_label) {
    if (fortran_order) PAINTR(u8, true); else PAINT(u8, true, false);
  } else if (fortran_order) {
    switch (out_width) { case 1: PAINTR(u8, false); break; case 2: PAINTR(u16, false); break;
                         case 4: PAINTR(u32, false); break; default: PAINTR(u64, false); break; }
  } else {
    switch (out_width) { case 1: PAINT(u8, false, false); break; case 2: PAINT(u16, false, false); break;
                         case 4: PAINT(u32, false, false); break; default: PAINT(u64, false, false); break; }
  }
#undef PAINT
#undef PAINTR
  LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// Re-encoding with another markov order (crackle::reencode_with_markov_order, src/crackle.hpp:860-984): the codepoints
// of every slice are recovered with the scan-parallel front end above (moves Mw, event list) and handed to the
// encoder's packers; nothing is decoded to voxels.
//
// Exact codepoint count of a slice: codes end with padding (up to 3 fields at order 0, up to 7 bits at order > 0).  The
// symbols end when the last chain of the beginning-of-chain index closes; every chain opens with one branch, a 'b' adds
// one and a 't' removes one, so the end is the first event at which (#t - #b) reaches the number of chains.
__global__ void __launch_bounds__(256) k_reenc_ncp(Geom g, const DecSlice* __restrict__ ds, const u8* __restrict__ stream, int order,
                                                    const u32* __restrict__ ncpIn, const u64* __restrict__ evOff,
                                                    const u32* __restrict__ evIdx, u32* __restrict__ sliceInfo) {
  __shared__ u32 sm[33];
  __shared__ u32 s_nch, s_best;
  const int xw = ckl_byte_width((u64)g.sx + 1), yw = ckl_byte_width((u64)g.sy + 1);
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    const DecSlice d = ds[z];
    if (threadIdx.x == 0) {
      u32 nch = 0;
      if (d.isz >= (u32)(4 + yw)) {
        const u8* p = stream + d.code;
        u64 idx = 4;
        const u32 ny = (u32)ld_le(p + idx, yw); idx += yw;
        for (u32 r = 0; r < ny && idx + yw + xw <= d.isz; r++) {
          idx += yw;
          const u32 nx = (u32)ld_le(p + idx, xw);
          idx += xw + (u64)nx * xw;
          nch += nx;
        }
      }
      s_nch = nch; s_best = 0xFFFFFFFFu;
    }
    __syncthreads();
    const u32 nch = s_nch;
    const u64 e0 = evOff[z];
    const u32 nev = (u32)(evOff[z + 1] - e0);
    const u32 cap = order > 0 ? ncpIn[z] : d.blen * 4u;
    int carry = 0;
    for (u32 j0 = 0; j0 < nev && nch; j0 += blockDim.x) {
      const u32 j = j0 + threadIdx.x;
      int v = 0;
      u32 e = 0;
      if (j < nev) { e = evIdx[e0 + j]; const u32 m = e >> 30; v = (m == 0 || m == 3) ? 1 : -1; }      // 't' : 'b'
      u32 tot;
      const int incl = carry + (int)block_excl_scan((u32)v, sm, tot) + v;
      if (j < nev && v == 1 && incl == (int)nch) atomicMin(&s_best, j);
      carry += (int)tot;
      __syncthreads();
      if (s_best != 0xFFFFFFFFu) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 ncp = 0;
      if (nch) ncp = s_best != 0xFFFFFFFFu ? (evIdx[e0 + s_best] & 0x3FFFFFFFu) + 1u : cap;
      sliceInfo[(u64)z * 4 + 0] = ncp;
      sliceInfo[(u64)z * 4 + 1] = nch;
      sliceInfo[(u64)z * 4 + 2] = d.isz;                          // the index is carried over verbatim
      sliceInfo[(u64)z * 4 + 3] = d.isz + (ncp + 3) / 4;          // order-0 size; the markov encoder overwrites it
    }
    __syncthreads();
  }
}
// absolute moves, 16 per word -> one byte per codepoint at the encoder's codepoint offsets
__global__ void __launch_bounds__(256) k_reenc_unpack(Geom g, const DecSlice* __restrict__ ds, const u32* __restrict__ Mw,
                                                       const u32* __restrict__ sliceInfo, const u64* __restrict__ cpOff, u8* __restrict__ cp) {
  for (u32 z = blockIdx.y; z < g.sz; z += gridDim.y) {
    const u32 ncp = sliceInfo[(u64)z * 4 + 0];
    const u32 nwords = (ncp + 15) / 16;
    const u32* M = Mw + ds[z].wordOff;
    u8* out = cp + cpOff[z];
    for (u32 w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
      const u32 m = M[w];
      const u32 n = min(16u, ncp - w * 16);
      for (u32 k = 0; k < n; k++) out[w * 16 + k] = (u8)((m >> (2 * k)) & 3u);
    }
  }
}
// beginning-of-chain index of every slice copied from the source stream; order 0 also packs the difference fields
__global__ void __launch_bounds__(256) k_reenc_emit(Geom g, const DecSlice* __restrict__ ds, const u8* __restrict__ stream,
                                                     const u32* __restrict__ sliceInfo, const u64* __restrict__ cpOff,
                                                     const u8* __restrict__ cp, const u64* __restrict__ codeOff, int pack0,
                                                     u8* __restrict__ dst) {
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    const DecSlice d = ds[z];
    u8* out = dst + codeOff[z];
    for (u32 b = threadIdx.x; b < d.isz; b += blockDim.x) out[b] = stream[d.code + b];
    if (!pack0) continue;
    const u32 ncp = sliceInfo[(u64)z * 4 + 0];
    const u8* c = cp + cpOff[z];
    for (u32 b = threadIdx.x; b < (ncp + 3) / 4; b += blockDim.x) {
      u32 last = b ? c[4 * b - 1] : 0, acc = 0;                    // differences carry across chains, initial 0
      for (u32 k = 0; k < 4; k++) {
        const u32 i = 4 * b + k;
        if (i < ncp) { const u32 v = c[i]; acc |= ((v - last) & 3u) << (2 * k); last = v; }
      }
      out[d.isz + b] = (u8)acc;
    }
  }
}

// front end: events of every slice (after launch_decode_classify + the host knows total_events)
void launch_reencode_codepoints(const Geom& g, const u8* stream, int order, DecodeBufs& D, u64 total_events, u32* sliceInfo,
                                cudaStream_t st) {
  const DecSlice* ds = D.slices.as<DecSlice>();
  D.evIdx.ensure(total_events * 4 + 16);
  if (total_events) {
    k_dec_compact<<<dec_grid(g.sz, 1, 8), 256, 0, st>>>(ds, g.sz, order, D.ncp.as<u32>(), D.Mw.as<u32>(), D.Sw.as<u32>(), D.evOff.as<u64>(),
                                                        D.evIdx.as<u32>(), D.evBaseW.as<u32>());
    LAUNCH_CHECK();
  }
  k_reenc_ncp<<<dec_grid(g.sz, 1, 8), 256, 0, st>>>(g, ds, stream, order, D.ncp.as<u32>(), D.evOff.as<u64>(), D.evIdx.as<u32>(), sliceInfo);
  LAUNCH_CHECK();
}
void launch_reencode_unpack(const Geom& g, DecodeBufs& D, const u32* sliceInfo, const u64* cpOff, u8* cp, u64 total_words, cudaStream_t st) {
  const u64 per_slice = (total_words + g.sz - 1) / g.sz;
  u32 gx = (u32)((per_slice + 255) / 256);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  k_reenc_unpack<<<dim3(gx, g.sz < 65535u ? g.sz : 65535u), 256, 0, st>>>(g, D.slices.as<DecSlice>(), D.Mw.as<u32>(), sliceInfo, cpOff, cp);
  LAUNCH_CHECK();
}
void launch_reencode_emit(const Geom& g, const u8* stream, DecodeBufs& D, const u32* sliceInfo, const u64* cpOff, const u8* cp,
                          const u64* codeOff, int pack0, u8* dst, cudaStream_t st) {
  k_reenc_emit<<<dec_grid(g.sz, 1, 8), 256, 0, st>>>(g, D.slices.as<DecSlice>(), stream, sliceInfo, cpOff, cp, codeOff, pack0, dst);
  LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// Voxel connectivity graph (operations.hpp:667-826; bit layout of crackcodes.hpp:706-862): one byte per voxel, a set bit =
// the neighbour in that direction is reachable, 00 -z +z -y +y -x +x.  The 2-D bits are the crack planes read the other way
// round: an interior edge is passable where the differ-plane bit is clear; the edges on the image border were never
// touched by a crack, so they keep the reference's fill value -- passable (1) for the IMPERMISSIBLE format (planes start at
// 0b1111 and cracks clear bits), closed (0) for PERMISSIBLE (planes start at 0 and cracks set bits).  Connectivity 6 adds
// +z / -z where the labels of vertically adjacent voxels agree (compared through their unique-table keys, painted as a
// uint32 volume) and opens the outer faces of the first and last decoded slice.  One thread per 4 pixels of a row.
template <bool Z>
__global__ void __launch_bounds__(256) k_vcg(Geom g, const u32* __restrict__ DV, const u32* __restrict__ DH, u32 border,
                                              const u32* __restrict__ key, u8* __restrict__ out) {
  const u32 q = (g.sx + 3) / 4;
  const u64 n = g.rows() * q, stride = (u64)gridDim.x * blockDim.x;
  const bool aligned = (g.sx & 3u) == 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 row = fdiv(i, q);
    const u32 x0 = (u32)(i - row * q) * 4u;
    const u32 z = (u32)fdiv(row, g.sy), y = (u32)(row - (u64)z * g.sy);
    const u64 wi = row * g.W + (x0 >> 5);
    const u32 sh = x0 & 31u;
    u32 v5 = (DV[wi] >> sh) & 0x1Fu;                              // differ bits of pixels x0 .. x0+4
    if (sh == 28u && x0 + 4 < g.sx) v5 |= (DV[wi + 1] & 1u) << 4;
    const u32 dh = (DH[wi] >> sh) & 0xFu;
    const u32 dhb = (y + 1 < g.sy) ? (DH[wi + g.W] >> sh) & 0xFu : 0u;
    const u64 o = (u64)z * g.sxy + (u64)y * g.sx + x0;
    u32 word = 0;
#pragma unroll
    for (u32 j = 0; j < 4; j++) {
      const u32 x = x0 + j;
      if (x >= g.sx) break;
      u32 b = (x + 1 < g.sx) ? (~(v5 >> (j + 1)) & 1u) : border;
      b |= ((x > 0) ? (~(v5 >> j) & 1u) : border) << 1;
      b |= ((y + 1 < g.sy) ? (~(dhb >> j) & 1u) : border) << 2;
      b |= ((y > 0) ? (~(dh >> j) & 1u) : border) << 3;
      if (Z) {
        const u32 k = key[o + j];
        b |= ((z + 1 < g.sz) ? (u32)(key[o + j + g.sxy] == k) : 1u) << 4;
        b |= ((z > 0) ? (u32)(key[o + j - g.sxy] == k) : 1u) << 5;
      }
      if (aligned) word |= b << (8u * j);
      else out[o + j] = (u8)b;
    }
    if (aligned) *reinterpret_cast<u32*>(out + o) = word;
  }
}
void launch_vcg(const Geom& g, const u32* DV, const u32* DH, int permissible, const u32* key, u8* out, cudaStream_t st) {
  const u64 n = g.rows() * (u64)((g.sx + 3) / 4);
  const u32 grid = grid1(n, 256, 148 * 16);
  if (key) k_vcg<true><<<grid, 256, 0, st>>>(g, DV, DH, permissible ? 0u : 1u, key, out);
  else k_vcg<false><<<grid, 256, 0, st>>>(g, DV, DH, permissible ? 0u : 1u, nullptr, out);
  LAUNCH_CHECK();
}
