// ckl_markov.cu -- order-N context ("markov") coding of the per-slice codepoint streams.
//
// Reference behaviour reproduced:
//   difference_codepoints / gather_statistics / CircularBuf   src/markov.hpp:97-220
//   stats_to_model (std::sort with a non-strict >= comparator)  src/markov.hpp:222-266
//   to_stored_model (5-bit permutation index, LSB first)        src/markov.hpp:325-380, LUT :43-68
//   encode_markov / markov::compress                            src/markov.hpp:422-489
// The context of a symbol is a sliding window of the previous `order` symbols, so statistics and encoding are
// data-parallel: every thread rebuilds its own context from the codepoint buffer.
#include "ckl_internal.cuh"

struct TraceParamsM {
  Geom g;
  const u64* offs;
  const u8* cp;
  const u32* sliceInfo;
};
static TraceParamsM mk(const Geom& g, TraceBufs& T) {
  TraceParamsM P;
  P.g = g; P.offs = T.offs.as<u64>(); P.cp = T.cp.as<u8>(); P.sliceInfo = T.sliceInfo.as<u32>();
  return P;
}
static u32 grid_slices(u32 sz, u32 per_sm) { return ckl_grid(sz, 1, per_sm, false); }
void launch_exscan_u32_u64(const u32* in, u32 n, u32 stride, u64* out, ull* total_out, u64 add_each, cudaStream_t st);
void launch_write_boc_only(const Geom& g, TraceBufs& T, u8* dst, cudaStream_t st);

// difference-coded symbol k of a slice: d[0] = cp[0], d[k] = (cp[k] - cp[k-1]) mod 4   (markov.hpp:182-189)
__device__ __forceinline__ u32 diff_at(const u8* cp, u32 k) { return k == 0 ? cp[0] : ((u32)cp[k] - (u32)cp[k - 1]) & 3u; }
// context row before symbol k: previous `order` symbols, oldest least significant, zero-filled before the slice
__device__ __forceinline__ u32 ctx_at(const u8* cp, u32 k, int order) {
  u32 ctx = 0;
  for (int m = 1; m <= order; m++) {
    if ((u32)m > k) break;
    ctx |= diff_at(cp, k - m) << (2 * (order - m));
  }
  return ctx;
}

__global__ void __launch_bounds__(256) k_markov_stats(TraceParamsM P, int order, u32* stats) {
  extern __shared__ u32 hist[];
  const u32 rows = 1u << (2 * order);
  const bool use_smem = order <= 5;
  const u64 n1 = (u64)P.g.sz + 1;
  if (use_smem) {
    for (u32 i = threadIdx.x; i < rows * 4; i += blockDim.x) hist[i] = 0;
    __syncthreads();
  }
  for (u32 z = blockIdx.x; z < P.g.sz; z += gridDim.x) {
    const u8* cp = P.cp + P.offs[3 * n1 + z];
    const u32 ncp = P.sliceInfo[(u64)z * 4 + 0];
    for (u32 k = threadIdx.x; k < ncp; k += blockDim.x) {
      const u32 idx = ctx_at(cp, k, order) * 4 + diff_at(cp, k);
      if (use_smem) atomicAdd(hist + idx, 1u); else atomicAdd(stats + idx, 1u);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (u32 i = threadIdx.x; i < rows * 4; i += blockDim.x) if (hist[i]) atomicAdd(stats + i, hist[i]);
  }
}

void launch_markov_stats(const Geom& g, TraceBufs& T, int order, u32* stats, cudaStream_t st) {
  const u64 rows = 1ull << (2 * order);
  CUDA_CHECK(cudaMemsetAsync(stats, 0, rows * 16, st));
  const size_t sm = order <= 5 ? rows * 16 : 0;
  k_markov_stats<<<grid_slices(g.sz, 4), 256, sm, st>>>(mk(g, T), order, stats);
  LAUNCH_CHECK();
}

// model[row][symbol] = rank; stored model = lexicographic index of (symbol of rank 0..3), 5 bits per row
__global__ void __launch_bounds__(256) k_markov_model(u32 rows, const u32* __restrict__ stats, u8* __restrict__ model, u32* stored_words) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  u32 cnt[4];
  for (int i = 0; i < 4; i++) cnt[i] = stats[(u64)r * 4 + i];
  // libstdc++ insertion sort with comp(a,b) = a.count >= b.count  (ties: later symbol moves ahead)
  int sym[4] = {0, 1, 2, 3};
  for (int i = 1; i < 4; i++) {
    const int v = sym[i];
    if (cnt[v] >= cnt[sym[0]]) {
      for (int j = i; j > 0; j--) sym[j] = sym[j - 1];
      sym[0] = v;
    } else {
      int j = i;
      while (cnt[v] >= cnt[sym[j - 1]]) { sym[j] = sym[j - 1]; j--; }
      sym[j] = v;
    }
  }
  for (int k = 0; k < 4; k++) model[(u64)r * 4 + sym[k]] = (u8)k;
  // lexicographic rank of the permutation (sym[0], sym[1], sym[2], sym[3])  == index into markov.hpp:43-68 LUT
  const int a = sym[0];
  const int b = sym[1] - (sym[1] > a ? 1 : 0);
  const int c = sym[2] - (sym[2] > a ? 1 : 0) - (sym[2] > sym[1] ? 1 : 0);
  const u32 idx = (u32)(a * 6 + b * 2 + c);
  const u64 bit = (u64)r * 5;
  const u32 sh = (u32)(bit & 31);
  atomicOr(stored_words + (bit >> 5), idx << sh);
  if (sh > 27) atomicOr(stored_words + (bit >> 5) + 1, idx >> (32 - sh));
}

void launch_markov_model(int order, const u32* stats, u8* model, u8* stored, u64 stored_bytes, cudaStream_t st) {
  const u32 rows = 1u << (2 * order);
  CUDA_CHECK(cudaMemsetAsync(stored, 0, ((stored_bytes + 3) / 4 + 1) * 4, st));
  k_markov_model<<<(rows + 255) / 256, 256, 0, st>>>(rows, stats, model, (u32*)stored);
  LAUNCH_CHECK();
}

// per-slice bitstream into word-aligned scratch: 2 raw bits for the first symbol, then unary-ish rank codes
//   rank 0 -> 0, 1 -> 10, 2 -> 110, 3 -> 111   written LSB first (patterns 0, 1, 3, 7)
__global__ void __launch_bounds__(256) k_markov_encode(TraceParamsM P, int order, const u8* __restrict__ model,
                                                        const u64* __restrict__ scratchOff, u32* scratch, u64* bitlen,
                                                        u32* sliceInfo) {
  __shared__ u32 sm[33];
  const u64 n1 = (u64)P.g.sz + 1;
  for (u32 z = blockIdx.x; z < P.g.sz; z += gridDim.x) {
    const u8* cp = P.cp + P.offs[3 * n1 + z];
    const u32 ncp = P.sliceInfo[(u64)z * 4 + 0];
    u32* out = scratch + scratchOff[z];
    u64 carry = 2;
    if (ncp && threadIdx.x == 0) atomicOr(out, (u32)cp[0]);
    for (u32 k0 = 1; k0 < ncp; k0 += blockDim.x) {
      const u32 k = k0 + threadIdx.x;
      u32 len = 0, pat = 0;
      if (k < ncp) {
        const u32 rank = model[(u64)ctx_at(cp, k, order) * 4 + diff_at(cp, k)];
        len = rank == 0 ? 1 : rank == 1 ? 2 : 3;
        pat = rank == 0 ? 0 : rank == 1 ? 1 : rank == 2 ? 3 : 7;
      }
      u32 tot;
      const u32 ex = block_excl_scan(len, sm, tot);
      if (pat) {
        const u64 bit = carry + ex;
        const u32 sh = (u32)(bit & 31);
        atomicOr(out + (bit >> 5), pat << sh);
        if (sh + len > 32) atomicOr(out + (bit >> 5) + 1, pat >> (32 - sh));
      }
      carry += tot;
    }
    if (threadIdx.x == 0) {
      const u64 bits = ncp ? carry : 0;
      bitlen[z] = bits;
      sliceInfo[(u64)z * 4 + 3] = sliceInfo[(u64)z * 4 + 2] + (u32)((bits + 7) / 8);
    }
    __syncthreads();
  }
}

__global__ void k_markov_caps(u32 sz, const u32* __restrict__ sliceInfo, u32* __restrict__ caps) {
  const u32 z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < sz) caps[z] = (u32)((2ull + 3ull * sliceInfo[(u64)z * 4 + 0] + 31) / 32 + 1);
}

void launch_markov_sizes(const Geom& g, TraceBufs& T, MarkovBufs& M, ull* scal, cudaStream_t st) {
  // scratch capacity per slice (words), offsets, then encode into scratch; sizes fall out of the encode
  M.bitlen.ensure((u64)g.sz * 8 + (u64)g.sz * 4);
  M.scratchOff.ensure(((u64)g.sz + 1) * 8);
  u32* caps = (u32*)(M.bitlen.as<u64>() + g.sz);
  k_markov_caps<<<(g.sz + 255) / 256, 256, 0, st>>>(g.sz, T.sliceInfo.as<u32>(), caps);
  LAUNCH_CHECK();
  launch_exscan_u32_u64(caps, g.sz, 1, M.scratchOff.as<u64>(), &scal[SC_CPCAP], 0, st);
}

// bitstreams of all slices into the word-aligned scratch (sizes land in sliceInfo[.codeBytes]); launch_markov_copy places them
void launch_markov_encode(const Geom& g, TraceBufs& T, int order, const u8* model, MarkovBufs& M, cudaStream_t st) {
  k_markov_encode<<<grid_slices(g.sz, 8), 256, 0, st>>>(mk(g, T), order, model, M.scratchOff.as<u64>(), M.scratch.as<u32>(),
                                                        M.bitlen.as<u64>(), T.sliceInfo.as<u32>());
  LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) k_markov_copy(Geom g, const u32* __restrict__ sliceInfo, const u64* __restrict__ codeOff,
                                                      const u64* __restrict__ scratchOff, const u32* __restrict__ scratch,
                                                      u8* __restrict__ dst) {
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    const u32 boc = sliceInfo[(u64)z * 4 + 2], total = sliceInfo[(u64)z * 4 + 3];
    const u8* src = (const u8*)(scratch + scratchOff[z]);
    u8* out = dst + codeOff[z] + boc;
    for (u32 b = threadIdx.x; b < total - boc; b += blockDim.x) out[b] = src[b];
  }
}
// bitstreams only (the caller places the beginning-of-chain indices itself)
void launch_markov_copy_body(const Geom& g, TraceBufs& T, MarkovBufs& M, u8* dst, cudaStream_t st) {
  k_markov_copy<<<grid_slices(g.sz, 8), 256, 0, st>>>(g, T.sliceInfo.as<u32>(), T.codeOff.as<u64>(), M.scratchOff.as<u64>(),
                                                      M.scratch.as<u32>(), dst);
  LAUNCH_CHECK();
}
void launch_markov_copy(const Geom& g, TraceBufs& T, MarkovBufs& M, u8* dst, cudaStream_t st) {
  launch_write_boc_only(g, T, dst, st);
  k_markov_copy<<<grid_slices(g.sz, 8), 256, 0, st>>>(g, T.sliceInfo.as<u32>(), T.codeOff.as<u64>(), M.scratchOff.as<u64>(),
                                                      M.scratch.as<u32>(), dst);
  LAUNCH_CHECK();
}
