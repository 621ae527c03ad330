// ckl_labels.cu -- flat label table: component -> label gather, global sorted unique table, keys, serialisation.
//
// Reference behaviour reproduced: labels::encode_flat global part, src/labels.hpp:90-154
//   uniq = sort+unique(mapping); keys[i] = index of mapping[i] in uniq; serialise
//   u64 n_uniq | uniq (stored width) | N_z (byte_width(sx*sy)) | keys (byte_width(n_uniq))
// CUB (CCCL) radix sort / select are used for the sort+unique of the (small) component-label list.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "ckl_internal.cuh"

static u32 grid1d(u64 n, u32 bs) {
  u64 b = (n + bs - 1) / bs;
  if (b < 1) b = 1;
  if (b > 148ull * 32 * (u64)g_ckl_grid_mult) b = 148ull * 32 * (u64)g_ckl_grid_mult;
  return (u32)b;
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather_mapping(const T* __restrict__ L, Geom g, u64 ncomp, const u64* __restrict__ compBase,
                                                         const u32* __restrict__ compPix, u64* __restrict__ mapping) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ncomp; i += stride) {
    u32 lo = 0, hi = g.sz;
    while (hi - lo > 1) { const u32 m = (lo + hi) >> 1; if (compBase[m] <= i) lo = m; else hi = m; }
    mapping[i] = (u64)L[(u64)lo * g.sxy + compPix[i]];
  }
}

void launch_gather_mapping(const void* labels, int width, const Geom& g, const CclBufs& B, u64 ncomp, u64* mapping, cudaStream_t st) {
  if (!ncomp) return;
  const u32 grid = grid1d(ncomp, 256);
  const u64* cb = B.compBase.as<u64>();
  const u32* px = B.compPix.as<u32>();
  switch (width) {
    case 1: k_gather_mapping<u8><<<grid, 256, 0, st>>>((const u8*)labels, g, ncomp, cb, px, mapping); break;
    case 2: k_gather_mapping<u16><<<grid, 256, 0, st>>>((const u16*)labels, g, ncomp, cb, px, mapping); break;
    case 4: k_gather_mapping<u32><<<grid, 256, 0, st>>>((const u32*)labels, g, ncomp, cb, px, mapping); break;
    default: k_gather_mapping<u64><<<grid, 256, 0, st>>>((const u64*)labels, g, ncomp, cb, px, mapping); break;
  }
  LAUNCH_CHECK();
}

// count_dev != nullptr: the number of unique labels is left in *count_dev (device, u64) and NOT read back -- no stream drain;
// the caller picks it up with its next scalar read-back.  Returns 0 in that case.
u64 labels_sort_unique(LabelBufs& L, u64 n, int key_bits, cudaStream_t st, ull* count_dev) {
  if (n == 0) {
    if (count_dev) CUDA_CHECK(cudaMemsetAsync(count_dev, 0, 8, st));
    return 0;
  }
  if (n > 0x7FFFFFFFull) throw CklError(CKL_ERR_ARG, "crackle_b200: more than 2^31 components in one shard");
  L.sorted.ensure(n * 8);
  L.uniq.ensure(n * 8);
  L.flags.ensure(16);
  const int end_bit = key_bits < 1 ? 1 : (key_bits > 64 ? 64 : key_bits);     // every key is below 2^key_bits
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, L.mapping.as<u64>(), L.sorted.as<u64>(), (int)n, 0, end_bit, st);
  cub::DeviceSelect::Unique(nullptr, t2, L.sorted.as<u64>(), L.uniq.as<u64>(), L.flags.as<u64>(), (int)n, st);
  L.tmp.ensure(t1 > t2 ? t1 : t2);
  size_t t = L.tmp.cap;
  CUDA_CHECK(cub::DeviceRadixSort::SortKeys(L.tmp.p, t, L.mapping.as<u64>(), L.sorted.as<u64>(), (int)n, 0, end_bit, st));
  t = L.tmp.cap;
  CUDA_CHECK(cub::DeviceSelect::Unique(L.tmp.p, t, L.sorted.as<u64>(), L.uniq.as<u64>(), L.flags.as<u64>(), (int)n, st));
  if (count_dev) {
    CUDA_CHECK(cudaMemcpyAsync(count_dev, L.flags.p, 8, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  u64 count = 0;
  CUDA_CHECK(cudaMemcpyAsync(&count, L.flags.p, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(ckl_sync(st));
  return count;
}

__device__ __forceinline__ void store_le(u8* p, u64 v, int w) {
  for (int i = 0; i < w; i++) p[i] = (u8)(v >> (8 * i));
}

__global__ void __launch_bounds__(256) k_write_keys(const u64* __restrict__ mapping, u64 n, const u64* __restrict__ uniq, u64 nu,
                                                     int kw, u8* __restrict__ dst) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 v = mapping[i];
    u64 lo = 0, hi = nu;
    while (lo < hi) { const u64 m = (lo + hi) >> 1; if (uniq[m] < v) lo = m + 1; else hi = m; }
    store_le(dst + i * (u64)kw, lo, kw);
  }
}
void launch_write_keys(const u64* mapping, u64 n, const u64* uniq, u64 n_uniq, int key_width, u8* dst, cudaStream_t st) {
  if (!n) return;
  k_write_keys<<<grid1d(n, 256), 256, 0, st>>>(mapping, n, uniq, n_uniq, key_width, dst);
  LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) k_write_le_u64(const u64* __restrict__ src, u64 n, int w, u8* __restrict__ dst) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) store_le(dst + i * (u64)w, src[i], w);
}
void launch_write_uniq(const u64* uniq, u64 n_uniq, int stored_width, u8* dst, cudaStream_t st) {
  if (!n_uniq) return;
  k_write_le_u64<<<grid1d(n_uniq, 256), 256, 0, st>>>(uniq, n_uniq, stored_width, dst);
  LAUNCH_CHECK();
}
__global__ void __launch_bounds__(256) k_write_le_u32(const u32* __restrict__ src, u64 n, int w, u8* __restrict__ dst) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) store_le(dst + i * (u64)w, src[i], w);
}
void launch_write_le_u32(const u32* src, u64 n, int width, u8* dst, cudaStream_t st) {
  if (!n) return;
  k_write_le_u32<<<grid1d(n, 256), 256, 0, st>>>(src, n, width, dst);
  LAUNCH_CHECK();
}
