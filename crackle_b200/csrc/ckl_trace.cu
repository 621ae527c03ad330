// ckl_trace.cu -- crack-code chain tracing on the edge bit-planes, exact reproduction of the reference's
// start-point and branch order, followed by the parallel post-passes (initial-branch reversal, spurious branch
// removal, escape coding, chain ordering) and the order-0 packer.
//
// Reference behaviour reproduced:
//   create_crack_codes walk / next_cluster / erase_edge     src/crackcodes.hpp:390-450, :41-64
//   remove_initial_branch                                   src/crackcodes.hpp:185-242
//   remove_spurious_branches                                src/crackcodes.hpp:250-281  (done online in the walker)
//   symbols_to_codepoints                                   src/crackcodes.hpp:128-183
//   write_boc_index / pack_codepoints                       src/crackcodes.hpp:318-372, :455-496
//
// Crack graph on bit-planes (pixel-indexed, W words per row): right edge of vertex (vx,vy) = EH bit vx of row vy
// (horizontal crack above pixel (vx,vy), vy >= 1); down edge of vertex (vx,vy) = EV bit vx of row vy (vertical
// crack left of pixel (vx,vy), vx >= 1).  Walking an edge clears its bit.
#include "ckl_internal.cuh"

#define NONE32 0xFFFFFFFFu

static int g_sms_t = 0;
static int sms() {
  if (!g_sms_t) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_t, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_t <= 0) g_sms_t = 148;
  }
  return g_sms_t;
}
static u32 grid_cap(u64 items, u32 per_block, u32 blocks_per_sm) {
  u64 need = (items + per_block - 1) / per_block;
  u64 cap = (u64)sms() * blocks_per_sm;
  if (need < 1) need = 1;
  return (u32)(need < cap ? need : cap);
}
void launch_exscan_u32_u64(const u32* in, u32 n, u32 stride, u64* out, ull* total_out, u64 add_each, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 valid_x(const Geom& g, u32 w) {
  return (w == g.W - 1 && (g.sx & 31)) ? ((1u << (g.sx & 31)) - 1u) : 0xFFFFFFFFu;
}
// effective crack words (IMPERMISSIBLE: cracks where labels differ; PERMISSIBLE: where they are equal)
__device__ __forceinline__ u32 crack_h(const Geom& g, const u32* DH, u64 idx, u32 y, u32 w, int perm) {
  const u32 d = DH[idx];
  return perm ? (y >= 1 ? (~d & valid_x(g, w)) : 0u) : d;
}
__device__ __forceinline__ u32 crack_v(const Geom& g, const u32* DV, u64 idx, u32 w, int perm) {
  const u32 d = DV[idx];
  return perm ? (~d & valid_x(g, w) & ~(w == 0 ? 1u : 0u)) : d;
}

// Builds the mutable crack planes and exact per-slice capacity bounds:
//   E = edges, B >= number of 'b' symbols (pushes), C >= number of chains.
__global__ void __launch_bounds__(256) k_trace_prepare(Geom g, const u32* __restrict__ DV, const u32* __restrict__ DH, int perm,
                                                        u32* __restrict__ EV, u32* __restrict__ EH, u32* bounds) {
  const u64 nwords = g.words(), stride = (u64)gridDim.x * blockDim.x;
  const u64 nloop = (nwords + stride - 1) / stride;
  for (u64 k = 0; k < nloop; k++) {
    const u64 i = k * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 z = NONE32, e = 0, b = 0, c = 0;
    if (i < nwords) {
      const u64 row = i / g.W;
      const u32 w = (u32)(i - row * g.W);
      z = (u32)(row / g.sy);
      const u32 y = (u32)(row - (u64)z * g.sy);
      const u32 r = crack_h(g, DH, i, y, w, perm);
      const u32 d = crack_v(g, DV, i, w, perm);
      EH[i] = r;
      EV[i] = d;
      const u32 rprev = w > 0 ? crack_h(g, DH, i - 1, y, w - 1, perm) : 0u;
      const u32 l = (r << 1) | (rprev >> 31);
      const u32 u = y > 0 ? crack_v(g, DV, i - g.W, w, perm) : 0u;
      // bit-sliced degree of the 32 vertices of this word
      const u32 s1 = r ^ l, c1 = r & l, s2 = d ^ u, c2 = d & u;
      const u32 sum0 = s1 ^ s2, carry = s1 & s2;
      const u32 two = c1 ^ c2 ^ carry, four = c1 & c2;
      const u32 deg3 = sum0 & two, deg2 = ~sum0 & two;
      const u32 corner = ~u & ~l & (r | d);                 // vertices that can start a chain
      e = __popc(r) + __popc(d);
      b = 2 * __popc(deg3) + 3 * __popc(four) + __popc(deg2 & corner);
      c = __popc(corner);
    }
    const u32 z0 = __shfl_sync(FULL_MASK, z, 0);
    if (__all_sync(FULL_MASK, z == z0)) {
      e = __reduce_add_sync(FULL_MASK, e);
      b = __reduce_add_sync(FULL_MASK, b);
      c = __reduce_add_sync(FULL_MASK, c);
      if ((threadIdx.x & 31) == 0 && z0 != NONE32) {
        if (e) atomicAdd(bounds + (u64)z0 * 4 + 0, e);
        if (b) atomicAdd(bounds + (u64)z0 * 4 + 1, b);
        if (c) atomicAdd(bounds + (u64)z0 * 4 + 2, c);
      }
    } else if (z != NONE32) {
      if (e) atomicAdd(bounds + (u64)z * 4 + 0, e);
      if (b) atomicAdd(bounds + (u64)z * 4 + 1, b);
      if (c) atomicAdd(bounds + (u64)z * 4 + 2, c);
    }
  }
}

// caps[z] = {symCap, stackCap, chainCap, cpCap}
__global__ void k_trace_caps(u32 sz, const u32* __restrict__ bounds, u32* __restrict__ caps) {
  const u32 z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= sz) return;
  const u32 E = bounds[(u64)z * 4], B = bounds[(u64)z * 4 + 1], C = bounds[(u64)z * 4 + 2];
  caps[(u64)z * 4 + 0] = E + 2 * B + C + 8;        // symbols: moves + b + t (t = b + chains)
  caps[(u64)z * 4 + 1] = B + 4;                    // revisit stack depth
  caps[(u64)z * 4 + 2] = C + 2;                    // chains
  caps[(u64)z * 4 + 3] = E + 4 * B + 2 * C + 16;   // codepoints: moves + 2b + 2t
}

void launch_trace_prepare(const Geom& g, const u32* DV, const u32* DH, int permissible, TraceBufs& T, ull* scal, cudaStream_t st) {
  const u64 nwords = g.words();
  T.EV.ensure(nwords * 4);
  T.EH.ensure(nwords * 4);
  T.bounds.ensure((u64)g.sz * 4 * 4 * 2);          // bounds (4 x u32) + caps (4 x u32) per slice
  T.offs.ensure(((u64)g.sz + 1) * 8 * 4);
  T.sliceInfo.ensure((u64)g.sz * 4 * 4);
  T.codeOff.ensure(((u64)g.sz + 1) * 8);
  u32* bounds = T.bounds.as<u32>();
  u32* caps = bounds + (u64)g.sz * 4;
  CUDA_CHECK(cudaMemsetAsync(bounds, 0, (u64)g.sz * 4 * 4, st));
  k_trace_prepare<<<grid_cap(nwords, 256, 16), 256, 0, st>>>(g, DV, DH, permissible, T.EV.as<u32>(), T.EH.as<u32>(), bounds);
  LAUNCH_CHECK();
  k_trace_caps<<<(g.sz + 255) / 256, 256, 0, st>>>(g.sz, bounds, caps);
  LAUNCH_CHECK();
  u64* offs = T.offs.as<u64>();
  const u64 n1 = (u64)g.sz + 1;
  launch_exscan_u32_u64(caps + 0, g.sz, 4, offs + 0 * n1, &scal[SC_SYMCAP], 0, st);
  launch_exscan_u32_u64(caps + 1, g.sz, 4, offs + 1 * n1, &scal[SC_STACKCAP], 0, st);
  launch_exscan_u32_u64(caps + 2, g.sz, 4, offs + 2 * n1, &scal[SC_CHAINCAP], 0, st);
  launch_exscan_u32_u64(caps + 3, g.sz, 4, offs + 3 * n1, &scal[SC_CPCAP], 0, st);
}

// ---------------------------------------------------------------------------------------------------------
// The walk.  One block per slice; the block finds the next vertex with edges cooperatively (next_cluster),
// thread 0 walks the chain.
struct TraceParams {
  Geom g;
  u32 *EV, *EH;
  const u64* offs;      // 4 arrays of (sz+1): sym, stack, chain, cp
  const u32* caps;      // per slice 4 x u32
  u8* sym;
  uint2* stack;
  ChainRec* chain;
  u8* cp;
  u32* cpPrefix;
  u32* sliceInfo;       // per slice: nsym|ncp, nchains, bocBytes, codeBytes
  ull* scal;
};

__device__ __forceinline__ u32 block_min_u32(u32 v, u32* sm) {
  v = __reduce_min_sync(FULL_MASK, v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  u32 r = NONE32;
  for (u32 i = 0; i < (blockDim.x >> 5); i++) r = min(r, sm[i]);
  __syncthreads();
  return r;
}

__device__ bool walk_chain(const Geom& g, u32* __restrict__ EV, u32* __restrict__ EH, u32 vx, u32 vy, u8* sym, u32& nsym, u32 symCap,
                           uint2* stack, u32 stackCap, ChainRec& rec, ull* scal) {
  bool ok = true;
  const u32 sxe = g.sx + 1, W = g.W;
  u32 x = vx, y = vy, sp = 0, nB = 0;
  const u32 begin = nsym;
  bool firstT = true, t2 = false, justPopped = false;
  u32 t2f = 0, poppedB = 0, adjStart = vx + sxe * vy;
  for (;;) {
    // adjacency of vertex (x,y): bit0 right, bit1 left, bit2 down, bit3 up
    u32 adj = 0;
    if (y < g.sy) {
      const u32* hrow = EH + (u64)y * W;
      if (x < g.sx) {
        adj |= (hrow[x >> 5] >> (x & 31)) & 1u;
        adj |= ((EV[(u64)y * W + (x >> 5)] >> (x & 31)) & 1u) << 2;
      }
      if (x > 0) adj |= ((hrow[(x - 1) >> 5] >> ((x - 1) & 31)) & 1u) << 1;
    }
    if (y > 0 && x < g.sx) adj |= ((EV[(u64)(y - 1) * W + (x >> 5)] >> (x & 31)) & 1u) << 3;

    if (adj == 0) {
      // a 't': dead end after a move, or -- directly after a pop -- a spurious branch (remove_spurious_branches)
      if (firstT) {
        firstT = false;
        if (nB == 1 && sym[begin] == 'b') { t2 = true; t2f = nsym - begin; adjStart = x + sxe * y; }   // remove_initial_branch applies
      }
      if (nsym >= symCap) { atomicExch(&scal[SC_ERROR], 1ull); ok = false; break; }
      if (justPopped && !(t2 && poppedB == begin)) { sym[poppedB] = 's'; sym[nsym++] = 's'; }
      else sym[nsym++] = 't';
      if (sp == 0) break;
      const uint2 e = stack[--sp];
      y = e.x / sxe;
      x = e.x - y * sxe;
      poppedB = e.y;
      justPopped = true;
      continue;
    }
    justPopped = false;
    if (nsym + 2 > symCap) { atomicExch(&scal[SC_ERROR], 1ull); ok = false; break; }
    if (adj & (adj - 1)) {                                  // popcount > 1: branch point
      if (sp >= stackCap) { atomicExch(&scal[SC_ERROR], 2ull); ok = false; break; }
      stack[sp++] = make_uint2(x + sxe * y, nsym);
      sym[nsym++] = 'b';
      nB++;
    }
    const u32 k = __ffs(adj) - 1;                            // priority: right, left, down, up
    if (k == 0) { EH[(u64)y * W + (x >> 5)] &= ~(1u << (x & 31)); sym[nsym++] = 'r'; x++; }
    else if (k == 1) { EH[(u64)y * W + ((x - 1) >> 5)] &= ~(1u << ((x - 1) & 31)); sym[nsym++] = 'l'; x--; }
    else if (k == 2) { EV[(u64)y * W + (x >> 5)] &= ~(1u << (x & 31)); sym[nsym++] = 'd'; y++; }
    else { EV[(u64)(y - 1) * W + (x >> 5)] &= ~(1u << (x & 31)); sym[nsym++] = 'u'; y--; }
  }
  rec.adjStart = adjStart;
  rec.symBegin = begin;
  rec.symEnd = nsym;
  rec.t2f = t2 ? t2f : 0;
  return ok;
}

__global__ void __launch_bounds__(128) k_trace_walk(TraceParams P) {
  __shared__ u32 sm[8];
  __shared__ u32 s_scan, s_nsym, s_nch;
  const Geom g = P.g;
  const u64 n1 = (u64)g.sz + 1;
  const u32 nw = g.sy * g.W;
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    u32* EV = P.EV + (u64)z * nw;
    u32* EH = P.EH + (u64)z * nw;
    u8* sym = P.sym + P.offs[0 * n1 + z];
    uint2* stack = P.stack + P.offs[1 * n1 + z];
    ChainRec* chains = P.chain + P.offs[2 * n1 + z];
    const u32 symCap = P.caps[(u64)z * 4 + 0], stackCap = P.caps[(u64)z * 4 + 1], chainCap = P.caps[(u64)z * 4 + 2];
    if (threadIdx.x == 0) { s_scan = 0; s_nsym = 0; s_nch = 0; }
    __syncthreads();
    for (;;) {
      // next_cluster: first word at or after s_scan with any edge bit (all earlier vertices are exhausted)
      u32 found = NONE32;
      const u32 start = s_scan;
      for (u32 base = start; base < nw && found == NONE32; base += blockDim.x * 8) {
        u32 cand = NONE32;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const u32 i = base + k * blockDim.x + threadIdx.x;
          if (cand == NONE32 && i < nw && (EH[i] | EV[i])) cand = i;
        }
        found = block_min_u32(cand, sm);
      }
      if (found == NONE32) break;
      if (threadIdx.x == 0) {
        const u32 vy = found / g.W, w = found - vy * g.W;
        const u32 m = EH[found] | EV[found];
        const u32 vx = w * 32 + (__ffs(m) - 1);
        u32 nsym = s_nsym, nch = s_nch;
        if (nch >= chainCap) { atomicExch(&P.scal[SC_ERROR], 3ull); s_scan = nw; }
        else {
          const bool ok = walk_chain(g, EV, EH, vx, vy, sym, nsym, symCap, stack, stackCap, chains[nch], P.scal);
          s_nsym = nsym;
          s_nch = nch + 1;
          s_scan = ok ? found : nw;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) { P.sliceInfo[(u64)z * 4 + 0] = s_nsym; P.sliceInfo[(u64)z * 4 + 1] = s_nch; }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// post-passes
__device__ __forceinline__ u8 flip_sym(u8 c) { return c == 'u' ? 'd' : c == 'd' ? 'u' : c == 'l' ? 'r' : c == 'r' ? 'l' : c; }
// symbol i of the chain after remove_initial_branch (positions 1..f-1 reversed and flipped, 0 and f skipped)
__device__ __forceinline__ u8 eff_sym(const u8* sym, const ChainRec& c, u32 i) {
  if (c.t2f) {
    const u32 j = i - c.symBegin;
    if (j == 0 || j == c.t2f) return 's';
    if (j < c.t2f) return flip_sym(sym[c.symBegin + c.t2f - j]);
  }
  return sym[i];
}
__device__ __forceinline__ u32 chain_of(const ChainRec* chains, u32 nch, u32 i) {
  u32 lo = 0, hi = nch;
  while (hi - lo > 1) { const u32 m = (lo + hi) >> 1; if (chains[m].symBegin <= i) lo = m; else hi = m; }
  return lo;
}
__device__ __forceinline__ u8 move_code(u8 s) { return s == 'u' ? 0 : s == 'r' ? 1 : s == 'd' ? 2 : 3; }   // crackcodes.hpp:20-26

__global__ void __launch_bounds__(256) k_trace_post(TraceParams P) {
  __shared__ u32 sm[33];
  const Geom g = P.g;
  const u64 n1 = (u64)g.sz + 1;
  const u32 sxe = g.sx + 1;
  const int xw = ckl_byte_width(g.sx + 1), yw = ckl_byte_width(g.sy + 1);
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    const u8* sym = P.sym + P.offs[0 * n1 + z];
    ChainRec* chains = P.chain + P.offs[2 * n1 + z];
    u8* cp = P.cp + P.offs[3 * n1 + z];
    u32* pre = P.cpPrefix + P.offs[0 * n1 + z];
    const u32 nsym = P.sliceInfo[(u64)z * 4 + 0], nch = P.sliceInfo[(u64)z * 4 + 1];
    // (a) codepoints per symbol -> exclusive prefix
    u32 carry = 0;
    for (u32 i0 = 0; i0 < nsym; i0 += blockDim.x) {
      const u32 i = i0 + threadIdx.x;
      u32 cnt = 0;
      if (i < nsym) {
        const u8 e = eff_sym(sym, chains[chain_of(chains, nch, i)], i);
        cnt = e == 's' ? 0 : (e == 'b' || e == 't') ? 2 : 1;
      }
      u32 tot;
      const u32 ex = block_excl_scan(cnt, sm, tot);
      if (i < nsym) pre[i] = carry + ex;
      carry += tot;
    }
    if (threadIdx.x == 0) pre[nsym] = carry;
    __syncthreads();
    const u32 ncp = carry;
    // (b) per-chain codepoint counts, (c) rank by adjusted start vertex (chains of a slice are vertex-disjoint)
    for (u32 c = threadIdx.x; c < nch; c += blockDim.x) chains[c].ncp = pre[chains[c].symEnd] - pre[chains[c].symBegin];
    for (u32 c = threadIdx.x; c < nch; c += blockDim.x) {
      const u32 a = chains[c].adjStart;
      u32 rank = 0;
      for (u32 o = 0; o < nch; o++) rank += chains[o].adjStart < a ? 1u : 0u;
      chains[rank].sortedIdx = c;
      chains[rank].sortedStart = a;
    }
    __syncthreads();
    // (d) output base of every chain in sorted order; count distinct start rows for the BOC index
    carry = 0;
    u32 nrows_part = 0;
    for (u32 r0 = 0; r0 < nch; r0 += blockDim.x) {
      const u32 r = r0 + threadIdx.x;
      const u32 v = r < nch ? chains[chains[r].sortedIdx].ncp : 0;
      u32 tot;
      const u32 ex = block_excl_scan(v, sm, tot);
      if (r < nch) {
        chains[chains[r].sortedIdx].outBase = carry + ex;
        if (r == 0 || chains[r].sortedStart / sxe != chains[r - 1].sortedStart / sxe) nrows_part++;
      }
      carry += tot;
    }
    u32 nrows;
    block_excl_scan(nrows_part, sm, nrows);
    // (e) absolute codepoints (symbols_to_codepoints), written at their final position
    for (u32 i = threadIdx.x; i < nsym; i += blockDim.x) {
      const ChainRec& c = chains[chain_of(chains, nch, i)];
      const u8 e = eff_sym(sym, c, i);
      if (e == 's') continue;
      u8* o = cp + c.outBase + (pre[i] - pre[c.symBegin]);
      if (e != 'b' && e != 't') { o[0] = move_code(e); continue; }
      // previous kept symbol(s)
      u32 j = i;
      u8 p = 0;
      u32 tcount = 0;            // kept 't' symbols between the previous move and this symbol
      while (j > c.symBegin) {
        j--;
        const u8 q = eff_sym(sym, c, j);
        if (q == 's') continue;
        if (q == 't') { tcount++; continue; }
        p = q;
        break;
      }
      if (e == 'b') {
        // (UP,DOWN) unless first symbol of the chain or the previous codepoint is DOWN -> (LEFT,RIGHT)
        const bool alt = (i == c.symBegin) || (tcount == 0 && p == 'd');
        o[0] = alt ? 3 : 0;
        o[1] = alt ? 1 : 2;
      } else {
        // (DOWN,UP) unless the previous codepoint is UP -> (RIGHT,LEFT); consecutive 't's alternate
        const bool alt = ((p == 'u') ? 1u : 0u) ^ (tcount & 1u);
        o[0] = alt ? 1 : 2;
        o[1] = alt ? 3 : 0;
      }
    }
    if (threadIdx.x == 0) {
      const u32 boc = 4 + yw + nrows * (yw + xw) + nch * xw;
      P.sliceInfo[(u64)z * 4 + 0] = ncp;
      P.sliceInfo[(u64)z * 4 + 2] = boc;
      P.sliceInfo[(u64)z * 4 + 3] = boc + (ncp + 3) / 4;
    }
    __syncthreads();
  }
}

static TraceParams make_params(const Geom& g, TraceBufs& T, ull* scal);
void launch_trace_walk(const Geom& g, TraceBufs& T, ull* scal, cudaStream_t st) {
  k_trace_walk<<<grid_cap(g.sz, 1, 16), 128, 0, st>>>(make_params(g, T, scal));
  LAUNCH_CHECK();
}
void launch_trace_post(const Geom& g, TraceBufs& T, ull* scal, cudaStream_t st) {
  k_trace_post<<<grid_cap(g.sz, 1, 8), 256, 0, st>>>(make_params(g, T, scal));
  LAUNCH_CHECK();
  // total codepoints (for the "all slices empty" rule, crackle.hpp:107-118)
  launch_exscan_u32_u64(T.sliceInfo.as<u32>(), g.sz, 4, T.codeOff.as<u64>(), &scal[SC_CODEPOINTS], 0, st);
}
void launch_trace(const Geom& g, TraceBufs& T, ull* scal, cudaStream_t st) {
  launch_trace_walk(g, T, scal, st);
  launch_trace_post(g, T, scal, st);
}

// ---------------------------------------------------------------------------------------------------------
// beginning-of-chain index (write_boc_index): serial per slice, chains are few
__device__ void put_le(u8* p, u64 v, int w) { for (int i = 0; i < w; i++) p[i] = (u8)(v >> (8 * i)); }
__device__ u32 write_boc(u8* dst, const ChainRec* chains, u32 nch, u32 boc_bytes, u32 sxe, int xw, int yw) {
  u32 o = 0;
  put_le(dst, boc_bytes - 4, 4); o += 4;
  // number of distinct rows
  u32 ny = 0;
  for (u32 r = 0; r < nch; r++) if (r == 0 || chains[r].sortedStart / sxe != chains[r - 1].sortedStart / sxe) ny++;
  put_le(dst + o, ny, yw); o += yw;
  u32 prev_y = 0;
  for (u32 r = 0; r < nch;) {
    const u32 y = chains[r].sortedStart / sxe;
    u32 e = r;
    while (e < nch && chains[e].sortedStart / sxe == y) e++;
    put_le(dst + o, y - prev_y, yw); o += yw; prev_y = y;
    put_le(dst + o, e - r, xw); o += xw;
    u32 prev_x = 0;
    for (u32 q = r; q < e; q++) { const u32 x = chains[q].sortedStart - y * sxe; put_le(dst + o, x - prev_x, xw); o += xw; prev_x = x; }
    r = e;
  }
  return o;
}

__global__ void __launch_bounds__(256) k_pack_order0(TraceParams P, const u64* __restrict__ codeOff, u8* __restrict__ dst) {
  const Geom g = P.g;
  const u64 n1 = (u64)g.sz + 1;
  const int xw = ckl_byte_width(g.sx + 1), yw = ckl_byte_width(g.sy + 1);
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    const ChainRec* chains = P.chain + P.offs[2 * n1 + z];
    const u8* cp = P.cp + P.offs[3 * n1 + z];
    const u32 ncp = P.sliceInfo[(u64)z * 4 + 0], nch = P.sliceInfo[(u64)z * 4 + 1], boc = P.sliceInfo[(u64)z * 4 + 2];
    u8* out = dst + codeOff[z];
    if (threadIdx.x == 0) write_boc(out, chains, nch, boc, g.sx + 1, xw, yw);
    const u32 nbytes = (ncp + 3) / 4;
    for (u32 b = threadIdx.x; b < nbytes; b += blockDim.x) {
      u32 last = b ? cp[4 * b - 1] : 0;                 // differences carry across chains, initial 0
      u32 acc = 0;
      for (u32 k = 0; k < 4; k++) {
        const u32 i = 4 * b + k;
        if (i < ncp) { const u32 c = cp[i]; acc |= ((c - last) & 3u) << (2 * k); last = c; }
      }
      out[boc + b] = (u8)acc;
    }
  }
}

void launch_code_sizes_order0(const Geom& g, TraceBufs& T, ull* scal, cudaStream_t st) {
  launch_exscan_u32_u64(T.sliceInfo.as<u32>() + 3, g.sz, 4, T.codeOff.as<u64>(), &scal[SC_CODE_BYTES], 0, st);
}

static TraceParams make_params(const Geom& g, TraceBufs& T, ull* scal) {
  TraceParams P;
  P.g = g;
  P.EV = T.EV.as<u32>(); P.EH = T.EH.as<u32>();
  P.offs = T.offs.as<u64>();
  P.caps = T.bounds.as<u32>() + (u64)g.sz * 4;
  P.sym = T.sym.as<u8>(); P.stack = T.stack.as<uint2>(); P.chain = T.chain.as<ChainRec>();
  P.cp = T.cp.as<u8>(); P.cpPrefix = T.cpPrefix.as<u32>(); P.sliceInfo = T.sliceInfo.as<u32>();
  P.scal = scal;
  return P;
}

void launch_pack_order0(const Geom& g, TraceBufs& T, u8* dst, cudaStream_t st) {
  TraceParams P = make_params(g, T, nullptr);
  k_pack_order0<<<grid_cap(g.sz, 1, 8), 256, 0, st>>>(P, T.codeOff.as<u64>(), dst);
  LAUNCH_CHECK();
}

// exported for ckl_markov.cu
__global__ void __launch_bounds__(64) k_write_boc_only(TraceParams P, const u64* __restrict__ codeOff, u8* __restrict__ dst) {
  const Geom g = P.g;
  const u64 n1 = (u64)g.sz + 1;
  const int xw = ckl_byte_width(g.sx + 1), yw = ckl_byte_width(g.sy + 1);
  const u32 z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= g.sz) return;
  const ChainRec* chains = P.chain + P.offs[2 * n1 + z];
  write_boc(dst + codeOff[z], chains, P.sliceInfo[(u64)z * 4 + 1], P.sliceInfo[(u64)z * 4 + 2], g.sx + 1, xw, yw);
}
void launch_write_boc_only(const Geom& g, TraceBufs& T, u8* dst, cudaStream_t st) {
  TraceParams P = make_params(g, T, nullptr);
  k_write_boc_only<<<(g.sz + 63) / 64, 64, 0, st>>>(P, T.codeOff.as<u64>(), dst);
  LAUNCH_CHECK();
}
