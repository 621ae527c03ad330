// ckl_planes.cu -- label volume -> "differ" bit-planes (the only full-width read of compress), run-based
// per-slice 4-connected component labelling on the bit-planes, raster-order component ranks, and the
// CRC-32C of each slice's (virtual) uint32 component image.
//
// Reference behaviour reproduced (results, not code):
//   lib::max_label / lib::pixel_pairs                  src/lib.hpp:224-256
//   Graph::init edge predicate                         src/crackcodes.hpp:66-125
//   cc3d::connected_components2d_4 / color_connectivity_graph + relabel   src/cc3d.hpp:114-369
//       final ids = rank of each component's first pixel in x-fastest raster order
//   crc32c(cc_labels)                                  src/labels.hpp:81, src/crackle.hpp:599-611
#include <algorithm>

#include "ckl_internal.cuh"

// ---------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T shfl_up1(T v) {
  if constexpr (sizeof(T) == 8) return (T)__shfl_up_sync(FULL_MASK, (ull)v, 1);
  else return (T)__shfl_up_sync(FULL_MASK, (u32)v, 1);
}

// One warp owns a strip of RS rows across the whole row width and walks it 32 pixels at a time: lane <-> x, the
// up-neighbour of a row is the previous row's register, the left neighbour comes from the lane below through one
// rotate-shuffle (lane 31 injects its value of the previous column, so lane 0 sees the pixel left of the word).
// Plane words are collected lane-per-word and stored 32 at a time (coalesced).
// DV bit x of row y: label(x,y) != label(x-1,y) (x>0).  DH bit: label(x,y) != label(x,y-1) (y>0).
// scal[SC_MAX] receives the bitwise OR of all labels: it has the same byte width as lib::max_label, which is all the
// caller derives from it (crackle.hpp:233-235).  pixel_pairs counts equal flat-order neighbours (lib.hpp:249-256).
template <typename T> __device__ __forceinline__ T shfl_idx(T v, u32 src) {
  if constexpr (sizeof(T) == 8) return (T)__shfl_sync(FULL_MASK, (ull)v, src);
  else return (T)__shfl_sync(FULL_MASK, (u32)v, src);
}

// one 32-pixel column step of a strip: FULL = all RS rows present (no per-row bounds tests)
template <typename T, int RS, bool FULL>
__device__ __forceinline__ void edges_column(const T* __restrict__ rowp, u64 sx, u32 nr, bool inx, bool has_left, bool top0, bool first_col,
                                             u32 lane, u32 wsel, T& up, u64& orall, u32& neq, u32 (&dvacc)[RS], u32 (&dhacc)[RS]) {
  T cur[RS], lf[RS];
#pragma unroll
  for (int r = 0; r < RS; r++) {
    cur[r] = 0; lf[r] = 0;
    if ((FULL || r < (int)nr) && inx) {
      cur[r] = rowp[(u64)r * sx];
      if (lane == 0 && (has_left || r > 0)) lf[r] = rowp[(u64)r * sx - 1];     // flat predecessor (previous word / previous row)
    }
  }
#pragma unroll
  for (int r = 0; r < RS; r++) {
    if (FULL || r < (int)nr) {
      const T v = cur[r];
      T l = shfl_up1<T>(v);
      if (lane == 0) l = lf[r];
      u32 vb = __ballot_sync(FULL_MASK, inx && v != l);
      u32 hb = __ballot_sync(FULL_MASK, inx && v != up);
      if (r == 0 && top0) hb = 0;
      if (lane == 0 && r == 0 && !has_left) vb |= 1u;      // flat index 0 has no predecessor: not an equal pair
      neq += __popc(vb);
      if (first_col) vb &= ~1u;
      if (lane == wsel) { dvacc[r] = vb; dhacc[r] = hb; }
      up = v;
      orall |= (u64)v;
    }
  }
}

template <typename T, int RS>
__global__ void __launch_bounds__(256) k_edges(const T* __restrict__ L, Geom g, u32* __restrict__ DV,
                                                u32* __restrict__ DH, ull* scal) {
  const u32 lane = threadIdx.x & 31;
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  const u32 ntile = (g.sy + RS - 1) / RS;
  const u64 items = (u64)g.sz * ntile;
  u64 orall = 0, pairs = 0;
  for (u64 it = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < items; it += nwarps) {
    const u32 z = (u32)(it / ntile), tile = (u32)(it - (u64)z * ntile);
    const u32 y0 = tile * RS;
    const u32 nr = min((u32)RS, g.sy - y0);
    const T* base = L + (u64)z * g.sxy + (u64)y0 * g.sx;
    const bool top0 = y0 == 0, has_prev = (z | y0) != 0;    // has_prev: row y0 has a flat predecessor at x = 0
    u32 dvacc[RS], dhacc[RS];
#pragma unroll
    for (int r = 0; r < RS; r++) dvacc[r] = dhacc[r] = 0;
    u32 neq = 0;
    for (u32 w = 0; w < g.W; w++) {
      const u32 x = w * 32 + lane;
      const bool inx = x < g.sx;
      T up = 0;
      if (!top0 && inx) up = *(base - g.sx + x);
      const bool has_left = w > 0 || has_prev;
      if (nr == RS) edges_column<T, RS, true>(base + x, g.sx, nr, inx, has_left, top0, w == 0, lane, w & 31, up, orall, neq, dvacc, dhacc);
      else edges_column<T, RS, false>(base + x, g.sx, nr, inx, has_left, top0, w == 0, lane, w & 31, up, orall, neq, dvacc, dhacc);
      if ((w & 31) == 31 || w == g.W - 1) {
        const u32 ws = (w & ~31u) + lane;
        if (ws <= w) {
#pragma unroll
          for (int r = 0; r < RS; r++) {
            if (r < (int)nr) {
              const u64 o = ((u64)z * g.sy + y0 + r) * g.W + ws;
              DV[o] = dvacc[r];
              DH[o] = dhacc[r];
            }
          }
        }
      }
    }
    // pixel_pairs of the strip: voxels minus unequal flat neighbours (neq counts lane-0's ballot copy: warp-uniform)
    pairs += (u64)g.sx * nr - neq;
  }
  // block reduction -> two atomics per block (pairs is warp-uniform: lane 0 carries it)
  __shared__ u64 s_or[8], s_pr[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) orall |= (u64)__shfl_xor_sync(FULL_MASK, (ull)orall, o);
  if (lane == 0) { s_or[threadIdx.x >> 5] = orall; s_pr[threadIdx.x >> 5] = pairs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (u32 i = 1; i < (blockDim.x >> 5); i++) { orall |= s_or[i]; pairs += s_pr[i]; }
    atomicOr(&scal[SC_MAX], (ull)orall);
    atomicAdd(&scal[SC_PAIRS], (ull)pairs);
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged form of the edge kernel (sx a multiple of 256, uint32 / uint64): a warp owns a band of 256 pixels and walks
// down a range of rows.  Rows arrive through the bulk-copy engine -- one cp.async.bulk of 2 KB (+ 16 bytes before the band,
// which holds the left neighbour of its first pixel) per row into a per-warp ring of EDT_STAGES shared-memory stages, each
// with its own mbarrier -- so the bytes in flight (stages x 2 KB per warp, ~190 KB per SM) do not occupy registers and no
// lane computes a load address.  The consumer side is lane <-> pixel: conflict-free ld.shared of the pixel and of its left
// neighbour (no shuffles), ballots against the previous row held in registers, and the band's 8 + 8 plane words of a row
// leave as two 32-byte stores.  The warp is its own producer: after a row has been compared its stage is re-armed with the
// row EDT_STAGES further down.
#define EDT_WORDS 8
#define EDT_STAGES 4
#define EDT_ROWS 256
#define EDT_WARPS 4
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
template <typename T> __device__ __forceinline__ T lds_t(u32 a) {
  if constexpr (sizeof(T) == 8) { ull v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return (T)v; }
  else { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return (T)v; }
}
template <typename T>
__global__ void __launch_bounds__(32 * EDT_WARPS) k_edges_tma(const T* __restrict__ L, Geom g, u32* __restrict__ DV, u32* __restrict__ DH, ull* scal) {
  constexpr u32 ROWB = 32 * EDT_WORDS * sizeof(T), STAGEB = ROWB + 16;
  extern __shared__ __align__(128) u8 edt_smem[];
  __shared__ __align__(8) u64 bars[EDT_WARPS][EDT_STAGES];
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const u32 ring = (u32)__cvta_generic_to_shared(edt_smem) + wib * (EDT_STAGES * STAGEB);
  const u32 bar0 = (u32)__cvta_generic_to_shared(&bars[wib][0]);
  if (lane == 0)
    for (u32 s0 = 0; s0 < EDT_STAGES; s0++) mbar_init(bar0 + 8 * s0, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const u64 nwarps = (u64)gridDim.x * EDT_WARPS;
  const u32 nband = g.sx / (32 * EDT_WORDS), nrange = (g.sy + EDT_ROWS - 1) / EDT_ROWS;
  const u64 items = (u64)g.sz * nrange * nband;
  u64 orall = 0, pairs = 0;
  u32 n = 0;                                               // rows consumed by this warp so far: stage = n % STAGES, parity = (n / STAGES) & 1
  for (u64 it = (u64)blockIdx.x * EDT_WARPS + wib; it < items; it += nwarps) {
    const u32 band = (u32)(it % nband);
    const u64 t2 = it / nband;
    const u32 range = (u32)(t2 % nrange), z = (u32)(t2 / nrange);
    const u32 y0 = range * EDT_ROWS, y1 = min(g.sy, y0 + EDT_ROWS);
    const u32 w0 = band * EDT_WORDS;
    const T* col = L + (u64)z * g.sxy + (u64)w0 * 32;
    // producer: row y of the band into stage (n + (y - y0)) % STAGES; the 16 bytes before the band come along (not for flat index 0)
    auto issue = [&](u32 y, u32 slot) {
      const u32 st = slot % EDT_STAGES;
      const bool first = (w0 | y | z) == 0;
      const T* src = col + (u64)y * g.sx;
      const u32 bytes = first ? ROWB : STAGEB;
      mbar_expect_tx(bar0 + 8 * st, bytes);
      bulk_g2s(ring + st * STAGEB + (first ? 16u : 0u), first ? (const void*)src : (const void*)((const u8*)src - 16), bytes, bar0 + 8 * st);
    };
    if (lane == 0)
      for (u32 k = 0; k < EDT_STAGES && y0 + k < y1; k++) issue(y0 + k, n + k);
    T up[EDT_WORDS];
#pragma unroll
    for (int j = 0; j < EDT_WORDS; j++) up[j] = y0 > 0 ? col[(u64)(y0 - 1) * g.sx + 32 * j + lane] : (T)0;
    u32 neq = 0;
    for (u32 y = y0; y < y1; y++, n++) {
      const u32 st = n % EDT_STAGES;
      mbar_wait(bar0 + 8 * st, (n / EDT_STAGES) & 1u);
      const u32 base = ring + st * STAGEB + 16 + lane * (u32)sizeof(T);
      const bool has_prev = (w0 | y | z) != 0;
      u32 mydv = 0, mydh = 0;
#pragma unroll
      for (int j = 0; j < EDT_WORDS; j++) {
        const T v = lds_t<T>(base + j * 32 * (u32)sizeof(T));
        const T l = lds_t<T>(base + j * 32 * (u32)sizeof(T) - (u32)sizeof(T));
        u32 vb = __ballot_sync(FULL_MASK, v != l);
        u32 hb = __ballot_sync(FULL_MASK, v != up[j]);
        if (j == 0 && !has_prev) vb |= 1u;                  // flat index 0 has no predecessor: not an equal pair (its pad is stale)
        neq += __popc(vb);
        if (j == 0 && w0 == 0) vb &= ~1u;                   // column 0 has no left neighbour inside its row
        if (y == 0) hb = 0;
        if (lane == (u32)j) mydv = vb;
        if (lane == (u32)j + EDT_WORDS) mydh = hb;
        orall |= (u64)v;
        up[j] = v;
      }
      const u64 o = ((u64)z * g.sy + y) * g.W + w0;
      if (lane < EDT_WORDS) DV[o + lane] = mydv;
      else if (lane < 2 * EDT_WORDS) DH[o + lane - EDT_WORDS] = mydh;
      __syncwarp();                                         // every lane has consumed the stage (the values above are in use)
      if (lane == 0 && y + EDT_STAGES < y1) issue(y + EDT_STAGES, n + EDT_STAGES);
    }
    pairs += (u64)(32 * EDT_WORDS) * (y1 - y0) - neq;       // warp-uniform
  }
  __shared__ u64 s_or[EDT_WARPS], s_pr[EDT_WARPS];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) orall |= (u64)__shfl_xor_sync(FULL_MASK, (ull)orall, o);
  if (lane == 0) { s_or[wib] = orall; s_pr[wib] = pairs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (u32 i = 1; i < EDT_WARPS; i++) { orall |= s_or[i]; pairs += s_pr[i]; }
    atomicOr(&scal[SC_MAX], (ull)orall);
    atomicAdd(&scal[SC_PAIRS], (ull)pairs);
  }
}

static u32 grid_for(u64 items, u32 per_block, u32 blocks_per_sm) { return ckl_grid(items, per_block, blocks_per_sm); }

void launch_edges(const void* labels, int width, const Geom& g, u32* DV, u32* DH, ull* scal, cudaStream_t st) {
  // TMA-staged form for uint32 / uint64 rows that are a multiple of 256 pixels
  if (g.sx % (32 * EDT_WORDS) == 0 && width >= 4 && ((u64)labels & 15) == 0) {
    const u64 items = (u64)g.sz * ((g.sy + EDT_ROWS - 1) / EDT_ROWS) * (g.sx / (32 * EDT_WORDS));
    const size_t smem = (size_t)EDT_WARPS * EDT_STAGES * (32 * EDT_WORDS * width + 16);
    const u32 per_sm = (u32)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 1024)));
    const u32 grid = grid_for(items, EDT_WARPS, per_sm);
    if (width == 4) {
      if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_edges_tma<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_edges_tma<u32><<<grid, 32 * EDT_WARPS, smem, st>>>((const u32*)labels, g, DV, DH, scal);
    } else {
      if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_edges_tma<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_edges_tma<u64><<<grid, 32 * EDT_WARPS, smem, st>>>((const u64*)labels, g, DV, DH, scal);
    }
    LAUNCH_CHECK();
    return;
  }
  // register-strip form (any shape, any width): one warp = 32 px x 16 rows, 16 loads in flight per lane.  Measured on B200
  // (1024^3 uint64): 8 rows 2106 us, 16 rows 1774 us, 32 rows 2072 us (250 registers)
  const u32 threads = 256u;
  const u64 items = (u64)g.sz * ((g.sy + 15) / 16);
  const u32 grid = grid_for(items, threads / 32, 4);
#define EDGES(T, R) k_edges<T, R><<<grid, threads, 0, st>>>((const T*)labels, g, DV, DH, scal)
  switch (width) { case 1: EDGES(u8, 16); break; case 2: EDGES(u16, 16); break; case 4: EDGES(u32, 16); break; default: EDGES(u64, 16); break; }
#undef EDGES
  LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// runs: a run is a maximal horizontal segment without a vertical crack.  Run index inside a row = number of DV
// bits at positions <= x; slice-local run id = rowBase[row] + that.
__global__ void __launch_bounds__(256) k_row_prefix(Geom g, const u32* __restrict__ DV, u32* __restrict__ wordPrefix,
                                                     u32* __restrict__ rowRuns) {
  const u32 lane = threadIdx.x & 31;
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  const u64 rows = g.rows();
  for (u64 row = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += nwarps) {
    u32 carry = 0;
    for (u32 w0 = 0; w0 < g.W; w0 += 32) {
      const u32 w = w0 + lane;
      const u32 v = w < g.W ? DV[row * g.W + w] : 0;
      const u32 c = __popc(v);
      const u32 inc = warp_incl_scan(c);
      if (w < g.W) wordPrefix[row * g.W + w] = carry + inc - c;
      carry += __shfl_sync(FULL_MASK, inc, 31);
    }
    if (lane == 0) rowRuns[row] = carry + 1;
  }
}

__global__ void __launch_bounds__(256) k_slice_scan(Geom g, const u32* __restrict__ rowRuns, u32* __restrict__ rowBase,
                                                     u32* __restrict__ sliceRuns) {
  __shared__ u32 sm[33];
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    u32 carry = 0;
    for (u32 y0 = 0; y0 < g.sy; y0 += blockDim.x) {
      const u32 y = y0 + threadIdx.x;
      const u32 v = y < g.sy ? rowRuns[(u64)z * g.sy + y] : 0;
      u32 tot;
      const u32 ex = block_excl_scan(v, sm, tot);
      if (y < g.sy) rowBase[(u64)z * g.sy + y] = carry + ex;
      carry += tot;
    }
    if (threadIdx.x == 0) sliceRuns[z] = carry;
  }
}

// single-block exclusive scan of n strided u32 values into u64 out[0..n] (out[n] = total), total also to *total_out
__global__ void __launch_bounds__(1024) k_exscan_u32_u64(const u32* __restrict__ in, u32 n, u32 stride, u64* __restrict__ out,
                                                          ull* total_out, u64 add_each) {
  __shared__ u64 swarp[33];
  const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  u64 carry = 0;
  for (u32 i0 = 0; i0 < n; i0 += blockDim.x) {
    const u32 i = i0 + threadIdx.x;
    const u64 v = i < n ? (u64)in[(u64)i * stride] + add_each : 0;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u64 t = __shfl_up_sync(FULL_MASK, (ull)inc, o);
      if (lane >= (u32)o) inc += t;
    }
    if (lane == 31) swarp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      u64 s = swarp[lane];
      u64 si = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(FULL_MASK, (ull)si, o);
        if (lane >= (u32)o) si += t;
      }
      swarp[lane] = si - s;
      if (lane == 31) swarp[32] = si;
    }
    __syncthreads();
    if (i < n) out[i] = carry + swarp[wid] + inc - v;
    carry += swarp[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[n] = carry;
    if (total_out) *total_out = carry;
  }
}
void launch_exscan_u32_u64(const u32* in, u32 n, u32 stride, u64* out, ull* total_out, u64 add_each, cudaStream_t st) {
  k_exscan_u32_u64<<<1, 1024, 0, st>>>(in, n, stride, out, total_out, add_each);
  LAUNCH_CHECK();
}

void launch_ccl_count(const Geom& g, const u32* DV, CclBufs& B, ull* scal, cudaStream_t st) {
  B.wordPrefix.ensure(g.words() * 4);
  B.rowRuns.ensure(g.rows() * 4);
  B.rowBase.ensure(g.rows() * 4);
  B.sliceRuns.ensure((u64)g.sz * 4);
  B.runBase.ensure(((u64)g.sz + 1) * 8);
  k_row_prefix<<<grid_for(g.rows(), 8, 8), 256, 0, st>>>(g, DV, B.wordPrefix.as<u32>(), B.rowRuns.as<u32>());
  LAUNCH_CHECK();
  k_slice_scan<<<grid_for(g.sz, 1, 8), 256, 0, st>>>(g, B.rowRuns.as<u32>(), B.rowBase.as<u32>(), B.sliceRuns.as<u32>());
  LAUNCH_CHECK();
  launch_exscan_u32_u64(B.sliceRuns.as<u32>(), g.sz, 1, B.runBase.as<u64>(), &scal[SC_RUNS], 0, st);
}

__device__ __forceinline__ u32 mask_le(u32 b) { return (2u << b) - 1u; }   // bits 0..b (b = 31 -> all ones)

__global__ void __launch_bounds__(256) k_run_init(Geom g, const u32* __restrict__ DV, const u32* __restrict__ wordPrefix,
                                                   const u32* __restrict__ rowBase, const u64* __restrict__ runBase,
                                                   u32* __restrict__ runStart) {
  const u64 nwords = g.words(), stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
    const u64 row = fdiv(i, g.W);
    const u32 w = (u32)(i - row * g.W);
    const u32 z = (u32)fdiv(row, g.sy), y = (u32)(row - (u64)z * g.sy);
    const u32 dv = DV[i];
    u32 starts = dv | (w == 0 ? 1u : 0u);
    const u32 rb = rowBase[row] + wordPrefix[i];
    const u64 gb = runBase[z];
    while (starts) {
      const u32 b = __ffs(starts) - 1;
      starts &= starts - 1;
      const u32 rid = rb + __popc(dv & mask_le(b));
      runStart[gb + rid] = y * g.sx + w * 32 + b;
    }
  }
}

// union-find with "smaller id wins" on a parent array that may live in shared memory; finds compress paths
// (atomicMin keeps parents monotonically decreasing, so concurrent compression never loses a link)
__device__ __forceinline__ u32 ufc_find(volatile u32* par, u32 a) {
  u32 p = par[a];
  while (p != a) {
    const u32 gp = par[p];
    if (gp != p) atomicMin((u32*)par + a, gp);
    a = p; p = gp;
  }
  return a;
}
__device__ __forceinline__ void ufc_unite(volatile u32* par, u32 a, u32 b) {
  for (;;) {
    a = ufc_find(par, a);
    b = ufc_find(par, b);
    if (a == b) return;
    if (a < b) { const u32 t = a; a = b; b = t; }   // a > b: hang a under b
    const u32 old = atomicMin((u32*)par + a, b);
    if (old == a) return;
    a = old;
  }
}

// unions between row `y` (y >= 1) and the row above, for the 32-pixel word `i` = (row, w); run ids relative to `base`
__device__ __forceinline__ void unite_word(const Geom& g, const u32* __restrict__ DV, const u32* __restrict__ DH,
                                           const u32* __restrict__ wordPrefix, const u32* __restrict__ rowBase, u64 row, u32 w,
                                           volatile u32* par, u32 base) {
  const u64 i = row * g.W + w;
  const u32 valid = (w == g.W - 1 && (g.sx & 31)) ? ((1u << (g.sx & 31)) - 1u) : 0xFFFFFFFFu;
  const u32 dv = DV[i], dvu = DV[i - g.W];
  const u32 conn = ~DH[i] & valid;                       // pixel connected to the pixel above
  u32 cand = conn & (dv | dvu | ~(conn << 1));           // first pixel of every (run, upper run, group) overlap
  if (!cand) return;
  const u32 rb = rowBase[row] + wordPrefix[i] - base, rbu = rowBase[row - 1] + wordPrefix[i - g.W] - base;
  while (cand) {
    const u32 b = __ffs(cand) - 1;
    cand &= cand - 1;
    ufc_unite(par, rb + __popc(dv & mask_le(b)), rbu + __popc(dvu & mask_le(b)));
  }
}

// Band-local linking, two sweeps over the words of rows y0+1 .. y1-1.  A candidate is the first pixel of a
// (run, upper run) overlap group.  PHASE 1: the FIRST group of every run links the run straight to the upper run
// (plain store, one writer per run; upper ids are smaller, so the forest is rooted at component minima-to-be).
// PHASE 2: every further group is a real merge -> union-find.  Most runs of a segmentation touch exactly one upper
// run, so almost all of the work is the atomic-free phase 1.
// PHASE 3 = phase 1 that also queues the candidates of phase 2 as (run << 16 | upper run) pairs of band-local ids, so the
// second sweep over the words is replaced by a walk over the (short) queue; *qn counts every queued pair, stored or not.
template <int PHASE>
__device__ __forceinline__ void link_word(const Geom& g, const u32* __restrict__ DV, const u32* __restrict__ DH,
                                          const u32* __restrict__ wordPrefix, const u32* __restrict__ rowBase, u64 row, u32 w,
                                          volatile u32* par, u32 base, u32* queue = nullptr, u32* qn = nullptr, u32 qcap = 0) {
  const u64 i = row * g.W + w;
  const u32 valid = (w == g.W - 1 && (g.sx & 31)) ? ((1u << (g.sx & 31)) - 1u) : 0xFFFFFFFFu;
  const u32 dv = DV[i], dvu = DV[i - g.W];
  const u32 conn = ~DH[i] & valid;                       // pixel connected to the pixel above
  if (!conn) return;
  const u32 connPrev = w ? ~DH[i - 1] : 0u;              // previous word of the row (all 32 pixels valid)
  const u32 contIn = (connPrev >> 31) & ~(dv | dvu) & 1u; // bit 0 continues the group of pixel 32w-1
  u32 cand = conn & (dv | dvu | ~((conn << 1) | contIn));
  if (!cand) return;
  const u32 rb = rowBase[row] + wordPrefix[i] - base, rbu = rowBase[row - 1] + wordPrefix[i - g.W] - base;
  // does the run that enters this word from the left already have a connected pixel?
  bool hadIn = false;
  if (w && !(dv & 1u)) {
    u32 j = w - 1, cj = connPrev;
    for (;;) {
      const u32 dvj = DV[row * g.W + j];
      if (dvj) { hadIn = (cj & ~((1u << (31 - __clz(dvj))) - 1u)) != 0; break; }
      if (cj) { hadIn = true; break; }
      if (j == 0) break;
      j--;
      cj = ~DH[row * g.W + j];
    }
  }
  while (cand) {
    const u32 b = __ffs(cand) - 1;
    cand &= cand - 1;
    const u32 below = dv & mask_le(b);                   // run starts at or before b
    const u32 s = below ? 31 - __clz(below) : 0;         // first pixel of the run inside this word
    const u32 seg = ((1u << b) - 1u) & ~((1u << s) - 1u);
    const bool first = !(conn & seg) && !(below == 0 && hadIn);
    const u32 r = rb + __popc(below), u = rbu + __popc(dvu & mask_le(b));
    if (PHASE == 1) { if (first) par[r] = u; }
    else if (PHASE == 3) {
      if (first) par[r] = u;
      else { const u32 q = atomicAdd(qn, 1u); if (q < qcap) queue[q] = (r << 16) | u; }
    }
    else if (!first) ufc_unite(par, r, u);
  }
}

// Band pass: one block owns CCL_BAND rows of one slice and solves them in shared memory (global memory when the
// band has too many runs); the band's trees are flattened and written out with slice-local run ids.
#define CCL_BAND 64
#define CCL_SMEM_RUNS 12288     // 48 KB of parents; with the 4 KB queue beside them the launch opts in to more than 48 KB
#define CCL_QUEUE 1024       // queued phase-2 unions per band (shared memory); a band with more takes the second sweep
__global__ void __launch_bounds__(256) k_band_ccl(Geom g, u32 nbands, const u32* __restrict__ DV, const u32* __restrict__ DH,
                                                   const u32* __restrict__ wordPrefix, const u32* __restrict__ rowBase,
                                                   const u32* __restrict__ sliceRuns, const u64* __restrict__ runBase, u32* parent,
                                                   const u32 smem_runs) {
  extern __shared__ u32 spar[];                          // smem_runs entries
  __shared__ u32 queue[CCL_QUEUE];
  __shared__ u32 qn;
  const u64 nitems = (u64)g.sz * nbands;
  for (u64 item = blockIdx.x; item < nitems; item += gridDim.x) {
    const u32 z = (u32)(item / nbands), band = (u32)(item - (u64)z * nbands);
    const u32 y0 = band * CCL_BAND, y1 = min(g.sy, y0 + CCL_BAND);
    const u64 row0 = (u64)z * g.sy + y0;
    const u32 base = rowBase[row0];
    const u32 end = y1 < g.sy ? rowBase[row0 + (y1 - y0)] : sliceRuns[z];
    const u32 n = end - base;
    u32* gpar = parent + runBase[z] + base;
    const bool sm = n <= smem_runs;
    volatile u32* par = sm ? spar : gpar;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) par[i] = i;
    if (threadIdx.x == 0) qn = 0;
    __syncthreads();
    const u32 nw = (y1 - y0 - 1) * g.W;                  // words of rows y0+1 .. y1-1
    const bool queued = n <= 65535u;                     // band-local ids fit the 16-bit halves of a queue entry
    for (u32 k = threadIdx.x; k < nw; k += blockDim.x) {
      const u32 r = k / g.W, w = k - r * g.W;
      if (queued) link_word<3>(g, DV, DH, wordPrefix, rowBase, row0 + 1 + r, w, par, base, queue, &qn, CCL_QUEUE);
      else link_word<1>(g, DV, DH, wordPrefix, rowBase, row0 + 1 + r, w, par, base);
    }
    __syncthreads();
    // pointer jumping: any ancestor is a valid parent, so the rounds need no barriers in between
    for (int round = 0; round < 3; round++)
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) { const u32 p = par[i]; const u32 pp = par[p]; if (pp != p) par[i] = pp; }
    __syncthreads();
    const u32 nq = qn;
    if (queued && nq <= CCL_QUEUE) {
      for (u32 k = threadIdx.x; k < nq; k += blockDim.x) { const u32 e = queue[k]; ufc_unite(par, e >> 16, e & 0xFFFFu); }
    } else {
      for (u32 k = threadIdx.x; k < nw; k += blockDim.x) {
        const u32 r = k / g.W, w = k - r * g.W;
        link_word<2>(g, DV, DH, wordPrefix, rowBase, row0 + 1 + r, w, par, base);
      }
    }
    __syncthreads();
    if (sm) {
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) gpar[i] = base + ufc_find(par, i);
    } else {
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) { const u32 r = ufc_find(par, i); par[i] = r; }
      __syncthreads();
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) par[i] += base;
    }
    __syncthreads();
  }
}

// Border pass: unions across band boundaries, in global memory on the flattened band trees
__global__ void __launch_bounds__(256) k_border_union(Geom g, u32 nbands, const u32* __restrict__ DV, const u32* __restrict__ DH,
                                                       const u32* __restrict__ wordPrefix, const u32* __restrict__ rowBase,
                                                       const u64* __restrict__ runBase, u32* parent) {
  const u64 nitems = (u64)g.sz * (nbands - 1) * g.W, stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nitems; i += stride) {
    const u32 w = (u32)(i % g.W);
    const u64 t = i / g.W;
    const u32 band = (u32)(t % (nbands - 1)) + 1, z = (u32)(t / (nbands - 1));
    unite_word(g, DV, DH, wordPrefix, rowBase, (u64)z * g.sy + (u64)band * CCL_BAND, w, parent + runBase[z], 0);
  }
}

// per slice: rank of every root run among the slice's roots (= raster rank of the component's first pixel)
__global__ void __launch_bounds__(256) k_root_rank(Geom g, const u32* __restrict__ parent, const u32* __restrict__ sliceRuns,
                                                    const u64* __restrict__ runBase, u32* __restrict__ compRank,
                                                    u32* __restrict__ nz) {
  __shared__ u32 sm[33];
  constexpr u32 PER = 8;                                 // consecutive runs per thread: eight independent loads per block scan
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    const u32 n = sliceRuns[z];
    const u64 gb = runBase[z];
    u32 carry = 0;
    for (u32 i0 = 0; i0 < n; i0 += blockDim.x * PER) {
      const u32 i = i0 + threadIdx.x * PER;
      u32 flags = 0;
#pragma unroll
      for (u32 j = 0; j < PER; j++)
        if (i + j < n && __ldcg(parent + gb + i + j) == i + j) flags |= 1u << j;
      u32 tot;
      u32 r = carry + block_excl_scan(__popc(flags), sm, tot);
#pragma unroll
      for (u32 j = 0; j < PER; j++)
        if ((flags >> j) & 1u) compRank[gb + i + j] = r++;
      carry += tot;
    }
    if (threadIdx.x == 0) nz[z] = carry;
  }
}

// CRC-32C of the virtual uint32 image cc[pixel] = runComp[run(pixel)].  The image is constant along runs and runs
// tile the slice in raster order, so with d_r = comp[r] ^ comp[r-1] placed at the run start and extending to the
// end of the image (GF(2)-linearity of the raw CRC):  raw = XOR_r  d_r(x) * H[sxy - start_r],
// H[m] = sum_{j=1..m} x^(32 j) mod P  (table built once per context).  One short GF(2) multiply per RUN.
// H table: level tables by one thread (H[0..B], H[k*B]), then a parallel fill  H[kB + j] = H[kB] * x^(32 j) ^ H[j]
#define CRCH_B 1024u
__global__ void k_crcH_levels(const CrcTables* __restrict__ tabs, u32 n, u32* __restrict__ Hl, u32* __restrict__ Gk) {
  if (threadIdx.x || blockIdx.x) return;
  u32 h = 0;
  Hl[0] = 0;
  for (u32 j = 1; j <= CRCH_B; j++) { h = crc_word(tabs->t, h ^ 0x80000000u, 0u); Hl[j] = h; }   // H[j] = x^32 (1 + H[j-1])
  const u32 xB = gf_xpow32(tabs->pw, CRCH_B);
  const u32 K = n / CRCH_B + 1;
  u32 gk = 0;
  Gk[0] = 0;
  for (u32 k = 1; k <= K; k++) { gk = gf_mul(gk, xB) ^ h; Gk[k] = gk; }                           // H[(k)B] = H[(k-1)B] x^(32B) ^ H[B]
}
__global__ void __launch_bounds__(256) k_crcH_fill(const CrcTables* __restrict__ tabs, u32 n, const u32* __restrict__ Hl,
                                                    const u32* __restrict__ Gk, u32* __restrict__ H) {
  __shared__ u32 pw[4][256];
  for (u32 i = threadIdx.x; i < 1024; i += blockDim.x) (&pw[0][0])[i] = (&tabs->pw[0][0])[i];
  __syncthreads();
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x; m <= n; m += stride) {
    const u32 k = (u32)(m / CRCH_B), j = (u32)(m - (u64)k * CRCH_B);
    H[m] = k ? (gf_mul(Gk[k], gf_xpow32(pw, j)) ^ Hl[j]) : Hl[j];
  }
}
static void ensure_crcH(CclBufs& B, u32 sxy, const CrcTables* d_tables, cudaStream_t st) {
  if (B.crcHn >= (u64)sxy + 1 && B.crcH.p) return;
  B.crcH.ensure(((u64)sxy + 1) * 4);
  B.crcHl.ensure(((u64)CRCH_B + 1) * 4 + ((u64)sxy / CRCH_B + 2) * 4);
  u32* Hl = B.crcHl.as<u32>();
  u32* Gk = Hl + CRCH_B + 1;
  k_crcH_levels<<<1, 1, 0, st>>>(d_tables, sxy, Hl, Gk);
  LAUNCH_CHECK();
  k_crcH_fill<<<grid_for((u64)sxy + 1, 256, 8), 256, 0, st>>>(d_tables, sxy, Hl, Gk, B.crcH.as<u32>());
  LAUNCH_CHECK();
  B.crcHn = (u64)sxy + 1;
}

// Fused per-run finish: component rank of the run (find + rank of the root), the run's CRC contribution (the rank
// of the previous run comes from the neighbouring lane), and
//   MODE 0 (compress):   first pixel of every component (-> label gather)
//   MODE 1 (decompress): label of the run = uniq[key[keyBase[z] + comp]]   (labels::decode_flat, labels.hpp:453-506)
// grid = (chunks, slices): no per-run slice search.
struct RunLabelSrc { const u8* uniq; const u8* keys; u64 n_uniq, n_keys; int sw, kw; const u64* keyBase; u64* runLabel;
                     const u64* uniq64; const u64* keys64; };   // aligned copies of the two stream tables (may be null)
// little-endian field of a compile-time width at an arbitrary (unaligned) stream offset
template <int W>
__device__ __forceinline__ u64 ld_le_w(const u8* __restrict__ p) {
  u64 v = 0;
#pragma unroll
  for (int i = 0; i < W; i++) v |= (u64)p[i] << (8 * i);
  return v;
}
__device__ __forceinline__ u64 ld_le_dev(const u8* __restrict__ p, int w) {
  switch (w) {                                             // warp-uniform
    case 1: return p[0];
    case 2: return ld_le_w<2>(p);
    case 4: return ld_le_w<4>(p);
    default: return ld_le_w<8>(p);
  }
}
// d(x) * h(x) for the 16-bit differences of component ids: the eight multiples h x^16 .. h x^23 once, then the two bytes of
// d select among them (Horner over bytes: the high byte's sum is multiplied by x^8 through the byte table).
__device__ __forceinline__ u32 gf_mul_d16(u32 d, u32 h, const u32* t0) {
  h = t0[h & 0xFF] ^ (h >> 8);                            // * x^8
  h = t0[h & 0xFF] ^ (h >> 8);                            // * x^16: the 16 low bits of d are x^16 .. x^31
  u32 a = 0, b = 0;                                       // a: bits 15..8 of d (x^0 .. x^7), b: bits 7..0 (x^8 .. x^15)
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a ^= h & (0u - ((d >> (15 - j)) & 1u));
    b ^= h & (0u - ((d >> (7 - j)) & 1u));
    h = (h >> 1) ^ (CKL_CRC_POLY & (0u - (h & 1u)));
  }
  return a ^ t0[b & 0xFF] ^ (b >> 8);
}

// One warp owns a contiguous range of a slice's runs and walks it 32 runs at a time, so the component of the run before
// lane 0's comes from the previous iteration (one lone find per range instead of one per 32 runs).
template <int MODE>
__global__ void __launch_bounds__(256) k_run_finish(Geom g, const u32* __restrict__ parent, const u32* __restrict__ sliceRuns,
                                                     const u64* __restrict__ runBase, const u32* __restrict__ compRank,
                                                     const u32* __restrict__ runStart, const u64* __restrict__ compBase,
                                                     const u32* __restrict__ H, const CrcTables* __restrict__ tabs,
                                                     u32* __restrict__ compPix, RunLabelSrc src, u32* sliceCrc) {
  __shared__ u32 t0[256];
  for (u32 i = threadIdx.x; i < 256; i += blockDim.x) t0[i] = tabs->t[0][i];
  __syncthreads();
  const u32 lane = threadIdx.x & 31;
  const u32 wps = gridDim.x * (blockDim.x >> 5);                  // warps per slice
  const u32 wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (u32 z = blockIdx.y; z < g.sz; z += gridDim.y) {
    const u32 n = sliceRuns[z];
    const u64 gb = runBase[z];
    const u32* par = parent + gb;
    const u32* rank = compRank + gb;
    const u32 L = (((n + wps - 1) / wps) + 31u) & ~31u;           // runs per warp, a multiple of 32
    const u32 lo = wid * L;
    if (lo >= n) continue;
    const u32 hi = min(n, lo + L);
    const u64 cb = compBase[z];
    const u64 kb = MODE != 0 ? src.keyBase[z] : 0;
    u32 carry = lo ? rank[uf_find(par, lo - 1)] : 0u;             // component of the run before the range (warp-uniform)
    u32 x = 0;
    for (u32 i0 = lo; i0 < hi; i0 += 32) {
      const u32 i = i0 + lane;
      u32 c = 0;
      if (i < hi) {
        const u32 root = uf_find(par, i);
        c = rank[root];
        if (MODE == 0) { if (root == i) compPix[cb + c] = runStart[gb + i]; }
        else if (MODE == 2) {                                   // compressed-domain statistics: the run's index into the unique table
          const u64 ki = kb + c;
          u64 key = ~0ull;
          if (ki < src.n_keys) key = src.keys64 ? src.keys64[ki] : ld_le_dev(src.keys + ki * (u64)src.kw, src.kw);
          src.runLabel[gb + i] = key;
        } else {
          const u64 ki = kb + c;
          u64 label = 0;
          if (ki < src.n_keys) {
            if (src.keys64) {                                   // aligned tables: one load each instead of kw + sw byte loads
              const u64 key = src.keys64[ki];
              if (key < src.n_uniq) label = src.uniq64[key];
            } else {
              const u64 key = ld_le_dev(src.keys + ki * (u64)src.kw, src.kw);
              if (key < src.n_uniq) label = ld_le_dev(src.uniq + key * (u64)src.sw, src.sw);
            }
          }
          src.runLabel[gb + i] = label;
        }
      }
      u32 cprev = __shfl_up_sync(FULL_MASK, c, 1);
      if (lane == 0) cprev = carry;
      carry = __shfl_sync(FULL_MASK, c, 31);
      if (i < hi) {
        const u32 d = c ^ cprev;
        if (d) {
          const u32 h = H[(u32)g.sxy - runStart[gb + i]];
          x ^= d < 65536u ? gf_mul_d16(d, h, t0) : gf_mul(d, h);
        }
      }
    }
    x = __reduce_xor_sync(FULL_MASK, x);
    if (lane == 0 && x) atomicXor(sliceCrc + z, x);
  }
}

__global__ void k_crc_finalize(u32* crc, u32 n, u32 init_term) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) crc[i] = ~(crc[i] ^ init_term);
}
void launch_crc_finalize_slices(u32* sliceCrc, u32 sz, u32 init_term, cudaStream_t st) {
  k_crc_finalize<<<(sz + 255) / 256, 256, 0, st>>>(sliceCrc, sz, init_term);
  LAUNCH_CHECK();
}

void launch_ccl_solve(const Geom& g, const u32* DV, const u32* DH, CclBufs& B, const CrcTables* d_tables, ull* scal,
                      u64 total_runs, cudaStream_t st) {
  (void)d_tables;
  const u64 nwords = g.words();
  B.nz.ensure((u64)g.sz * 4);
  B.compBase.ensure(((u64)g.sz + 1) * 8);
  B.sliceCrc.ensure((u64)g.sz * 4);
  u32* parent = B.parent.as<u32>();
  k_run_init<<<grid_for(nwords, 256, 16), 256, 0, st>>>(g, DV, B.wordPrefix.as<u32>(), B.rowBase.as<u32>(), B.runBase.as<u64>(),
                                                         B.runStart.as<u32>());
  LAUNCH_CHECK();
  const u32 nbands = (g.sy + CCL_BAND - 1) / CCL_BAND;
  {
    // shared-memory capacity per band: twice the average number of runs per band (bands above it solve in global memory),
    // between 3072 and CCL_SMEM_RUNS entries -- a smaller footprint keeps more blocks resident on this latency-bound kernel
    const u64 avg = total_runs / ((u64)g.sz * nbands) + 1;
    u32 cap = (u32)std::min<u64>(CCL_SMEM_RUNS, std::max<u64>(3072, 2 * avg));
    cap = (cap + 1023u) & ~1023u;
    const u32 per_sm = std::min<u32>(8u, (u32)((200u * 1024u) / (cap * 4u + CCL_QUEUE * 4u + 1024u)));
    if ((size_t)cap * 4 + CCL_QUEUE * 4 + 64 > 48 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(k_band_ccl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CCL_SMEM_RUNS * 4)));
    k_band_ccl<<<grid_for((u64)g.sz * nbands, 1, per_sm), 256, (size_t)cap * 4, st>>>(g, nbands, DV, DH, B.wordPrefix.as<u32>(),
                                                                                       B.rowBase.as<u32>(), B.sliceRuns.as<u32>(),
                                                                                       B.runBase.as<u64>(), parent, cap);
  }
  LAUNCH_CHECK();
  if (nbands > 1) {
    k_border_union<<<grid_for((u64)g.sz * (nbands - 1) * g.W, 256, 8), 256, 0, st>>>(g, nbands, DV, DH, B.wordPrefix.as<u32>(),
                                                                                       B.rowBase.as<u32>(), B.runBase.as<u64>(), parent);
    LAUNCH_CHECK();
  }
  k_root_rank<<<grid_for(g.sz, 1, 8), 256, 0, st>>>(g, parent, B.sliceRuns.as<u32>(), B.runBase.as<u64>(), B.compRank.as<u32>(),
                                                     B.nz.as<u32>());
  LAUNCH_CHECK();
  launch_exscan_u32_u64(B.nz.as<u32>(), g.sz, 1, B.compBase.as<u64>(), &scal[SC_COMPONENTS], 0, st);
}

// little-endian fields of `width` bytes at an arbitrary stream offset -> aligned uint64 array (the decoder's unique-label and
// key tables: every run looks both up, so the unaligned byte loads are paid once per entry instead of once per run)
__global__ void __launch_bounds__(256) k_unpack_le(const u8* __restrict__ src, int width, u64 n, u64* __restrict__ dst) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = ld_le_dev(src + i * (u64)width, width);
}
void launch_unpack_le(const u8* src, int width, u64 n, u64* dst, cudaStream_t st) {
  if (!n) return;
  k_unpack_le<<<grid_for(n, 256, 8), 256, 0, st>>>(src, width, n, dst);
  LAUNCH_CHECK();
}

// second half of the solve: ranks -> per-run results + per-slice CRCs.  `decode` selects MODE 1 (run labels).
void launch_ccl_finish(const Geom& g, CclBufs& B, u64 total_runs, const CrcTables* d_tables, u32 crc_init_term,
                       const CclDecodeSrc* decode, cudaStream_t st) {
  ensure_crcH(B, (u32)g.sxy, d_tables, st);
  CUDA_CHECK(cudaMemsetAsync(B.sliceCrc.p, 0, (u64)g.sz * 4, st));
  if (total_runs) {
    const u64 per_slice = (total_runs + g.sz - 1) / g.sz;
    u32 gx = (u32)((per_slice + 255) / 256);
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    const dim3 grid(gx, g.sz < 65535u ? g.sz : 65535u);
    RunLabelSrc src{};
    if (decode && decode->keys_only) {
      src.uniq = decode->uniq; src.keys = decode->keys; src.n_uniq = decode->n_uniq; src.n_keys = decode->n_keys;
      src.sw = decode->sw; src.kw = decode->kw; src.keyBase = decode->keyBase; src.runLabel = decode->runLabel;
      src.uniq64 = decode->uniq64; src.keys64 = decode->keys64;
      k_run_finish<2><<<grid, 256, 0, st>>>(g, B.parent.as<u32>(), B.sliceRuns.as<u32>(), B.runBase.as<u64>(), B.compRank.as<u32>(),
                                           B.runStart.as<u32>(), B.compBase.as<u64>(), B.crcH.as<u32>(), d_tables, B.compPix.as<u32>(),
                                           src, B.sliceCrc.as<u32>());
    } else if (decode) {
      src.uniq = decode->uniq; src.keys = decode->keys; src.n_uniq = decode->n_uniq; src.n_keys = decode->n_keys;
      src.sw = decode->sw; src.kw = decode->kw; src.keyBase = decode->keyBase; src.runLabel = decode->runLabel;
      src.uniq64 = decode->uniq64; src.keys64 = decode->keys64;
      k_run_finish<1><<<grid, 256, 0, st>>>(g, B.parent.as<u32>(), B.sliceRuns.as<u32>(), B.runBase.as<u64>(), B.compRank.as<u32>(),
                                           B.runStart.as<u32>(), B.compBase.as<u64>(), B.crcH.as<u32>(), d_tables, B.compPix.as<u32>(),
                                           src, B.sliceCrc.as<u32>());
    } else {
      k_run_finish<0><<<grid, 256, 0, st>>>(g, B.parent.as<u32>(), B.sliceRuns.as<u32>(), B.runBase.as<u64>(), B.compRank.as<u32>(),
                                           B.runStart.as<u32>(), B.compBase.as<u64>(), B.crcH.as<u32>(), d_tables, B.compPix.as<u32>(),
                                           src, B.sliceCrc.as<u32>());
    }
    LAUNCH_CHECK();
  }
  launch_crc_finalize_slices(B.sliceCrc.as<u32>(), g.sz, crc_init_term, st);
}

// ---------------------------------------------------------------------------------------------------------
// generic CRC-32C of a device byte buffer
__global__ void __launch_bounds__(256) k_crc_bytes(const u8* __restrict__ d, u64 n, const CrcTables* __restrict__ tabs, u32* acc) {
  __shared__ u32 t[4][256];
  __shared__ u32 pw[4][256];
  for (u32 i = threadIdx.x; i < 1024; i += blockDim.x) {
    (&t[0][0])[i] = (&tabs->t[0][0])[i];
    (&pw[0][0])[i] = (&tabs->pw[0][0])[i];
  }
  __syncthreads();
  // message = head (n % 4 bytes, run from the real init state by chunk 0) || body of whole uint32 words
  const u64 head = n & 3, nw = n >> 2;
  const u64 CH = 16;                                   // words per chunk
  const u64 nchunks = (nw + CH - 1) / CH;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  u32 x = 0;
  for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < (nchunks ? nchunks : 1); c += stride) {
    u32 crc = 0;
    if (c == 0) {
      crc = 0xFFFFFFFFu;                               // carries the init term for the whole message
      for (u64 i = 0; i < head; i++) crc = crc_byte(t, crc, d[i]);
    }
    const u64 w0 = c * CH, w1 = min(nw, w0 + CH);
    for (u64 wi = w0; wi < w1; wi++) {
      const u8* p = d + head + wi * 4;
      const u32 v = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
      crc = crc_word(t, crc, v);
    }
    u64 after = nw - w1;
    while (after) {
      const u32 part = (u32)(after > 0xFFFFFFFFull ? 0xFFFFFFFFull : after);
      crc = gf_mul(crc, gf_xpow32(pw, part));
      after -= part;
    }
    x ^= crc;
  }
  x = __reduce_xor_sync(FULL_MASK, x);
  if ((threadIdx.x & 31) == 0 && x) atomicXor(acc, x);
}
__global__ void k_not(u32* v) { *v = ~*v; }

void launch_crc_bytes(const u8* d, u64 n, const CrcTables* d_tables, const CrcTables& h_tables, u32* d_out, cudaStream_t st) {
  (void)h_tables;
  CUDA_CHECK(cudaMemsetAsync(d_out, 0, 4, st));
  const u64 nchunks = ((n >> 2) + 15) / 16;
  k_crc_bytes<<<grid_for(nchunks ? nchunks : 1, 256, 8), 256, 0, st>>>(d, n, d_tables, d_out);
  LAUNCH_CHECK();
  k_not<<<1, 1, 0, st>>>(d_out);
  LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// Compressed-domain statistics (src/operations.hpp:321-665 voxel_counts / centroids / bounding_boxes): the reference
// paints the slice's component image and loops over its pixels; here every RUN contributes its length, coordinate sums
// and extent to the tables of its label (index into the sorted unique table) -- no full-width image exists.
// counts u64[nu]; sums u64[nu][3] (x, y, z); bbox u32[nu][6] (xmin, ymin, zmin, xmax, ymax, zmax)
__global__ void __launch_bounds__(256) k_stats_init(u64 nu, ull* counts, ull* sums, u32* bbox) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nu * 6; i += stride) {
    if (i < nu) counts[i] = 0;
    if (i < nu * 3) sums[i] = 0;
    bbox[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;
  }
}
__global__ void __launch_bounds__(256) k_run_stats(Geom g, u32 z_first, const u32* __restrict__ sliceRuns, const u64* __restrict__ runBase,
                                                    const u32* __restrict__ runStart, const u64* __restrict__ runKey, u64 nu,
                                                    ull* counts, ull* sums, u32* bbox) {
  for (u32 z = blockIdx.y; z < g.sz; z += gridDim.y) {
    const u32 n = sliceRuns[z];
    const u64 gb = runBase[z];
    const u32 zabs = z_first + z;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const u64 key = runKey[gb + i];
      if (key >= nu) continue;                                   // corrupt key: the reference would read out of bounds
      const u32 start = runStart[gb + i];
      const u32 end = i + 1 < n ? runStart[gb + i + 1] : (u32)g.sxy;   // runs tile the slice in raster order, one row each
      const u32 len = end - start;
      const u32 y = start / g.sx, x0 = start - y * g.sx, x1 = x0 + len - 1;
      atomicAdd(counts + key, (ull)len);
      atomicAdd(sums + key * 3 + 0, (ull)len * x0 + (ull)len * (len - 1) / 2);
      atomicAdd(sums + key * 3 + 1, (ull)len * y);
      atomicAdd(sums + key * 3 + 2, (ull)len * zabs);
      u32* b = bbox + key * 6;                                   // a plain read first: most runs do not extend the box
      if (x0 < __ldcg(b + 0)) atomicMin(b + 0, x0);
      if (y < __ldcg(b + 1)) atomicMin(b + 1, y);
      if (zabs < __ldcg(b + 2)) atomicMin(b + 2, zabs);
      if (x1 > __ldcg(b + 3)) atomicMax(b + 3, x1);
      if (y > __ldcg(b + 4)) atomicMax(b + 4, y);
      if (zabs > __ldcg(b + 5)) atomicMax(b + 5, zabs);
    }
  }
}
void launch_run_stats(const Geom& g, u32 z_first, const CclBufs& B, u64 total_runs, const u64* runKey, u64 nu, ull* counts, ull* sums,
                      u32* bbox, cudaStream_t st) {
  if (!nu) return;
  k_stats_init<<<grid_for(nu * 6, 256, 8), 256, 0, st>>>(nu, counts, sums, bbox);
  LAUNCH_CHECK();
  if (!total_runs) return;
  const u64 per_slice = (total_runs + g.sz - 1) / g.sz;
  u32 gx = (u32)((per_slice + 255) / 256);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  k_run_stats<<<dim3(gx, g.sz < 65535u ? g.sz : 65535u), 256, 0, st>>>(g, z_first, B.sliceRuns.as<u32>(), B.runBase.as<u64>(),
                                                                        B.runStart.as<u32>(), runKey, nu, counts, sums, bbox);
  LAUNCH_CHECK();
}
