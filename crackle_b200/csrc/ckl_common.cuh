// ckl_common.cuh -- shared types, error handling, GF(2)/CRC-32C arithmetic and small device helpers
// for the B200-native crackle hot path.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/crackle_b200.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned long long ull;

struct CklError : std::runtime_error {
  int code;
  CklError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(x)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess)                                                                         \
      throw CklError(CKL_ERR_CUDA, std::string("crackle_b200: CUDA error: ") + cudaGetErrorString(e_) + \
                                       " (" #x ") at " __FILE__ ":" + std::to_string(__LINE__));  \
  } while (0)

// every kernel launch is followed by LAUNCH_CHECK(): error check + a process-wide launch counter (bench.py reports it)
extern unsigned long long g_ckl_launches;
// grid-cap multiplier of the grid-stride kernels: > 1 while z-chunks run concurrently, so blocks are short and the
// hardware block scheduler can interleave the chunks by stream priority instead of queueing behind persistent grids
extern int g_ckl_grid_mult;
// every host-side wait on a compute stream goes through ckl_sync(): a process-wide drain counter (bench.py reports it per call)
extern unsigned long long g_ckl_syncs;
inline cudaError_t ckl_sync(cudaStream_t st) { __atomic_fetch_add(&g_ckl_syncs, 1ull, __ATOMIC_RELAXED); return cudaStreamSynchronize(st); }
#define LAUNCH_CHECK() do { __atomic_fetch_add(&g_ckl_launches, 1ull, __ATOMIC_RELAXED); CUDA_CHECK(cudaGetLastError()); } while (0)

// 64-bit index / 32-bit divisor: the full 64-bit division costs ~70 instructions; almost every index fits 32 bits
__host__ __device__ __forceinline__ u64 fdiv(u64 a, u32 d) { return (a >> 32) ? a / d : (u64)((u32)a / d); }

// Grow-only device buffer (the context keeps these across calls so steady-state calls do no cudaMalloc).
struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  void ensure(size_t n) {
    if (n <= cap) return;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t c = n + n / 16 + 256;
    cudaError_t e = cudaMalloc(&p, c);
    if (e != cudaSuccess) {
      p = nullptr;
      throw CklError(CKL_ERR_NOMEM, std::string("crackle_b200: cudaMalloc of ") + std::to_string(c) + " bytes failed: " + cudaGetErrorString(e));
    }
    cap = c;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  ~DBuf() { release(); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};

// lib.hpp:236-247 compute_byte_width
__host__ __device__ inline int ckl_byte_width(u64 x) { return x <= 0xFFull ? 1 : x <= 0xFFFFull ? 2 : x <= 0xFFFFFFFFull ? 4 : 8; }

// ---------------------------------------------------------------------------------------------------------
// CRC-32C (Castagnoli, reflected 0x82F63B78, init ~0, final ~: third_party/fastcrc/crc32c_portable.h:195-218)
// as GF(2)[x] arithmetic.  Bit 31 of a word is x^0 (zlib convention for reflected CRCs).
#define CKL_CRC_POLY 0x82F63B78u

// a(x) * b(x) mod P(x)
__host__ __device__ inline u32 gf_mul(u32 a, u32 b) {
  u32 p = 0;
#pragma unroll 8
  for (int i = 0; i < 32; i++) {
    p ^= b & (0u - ((a >> (31 - i)) & 1u));
    b = (b >> 1) ^ (CKL_CRC_POLY & (0u - (b & 1u)));
  }
  return p;
}

// Host-built tables: slicing-by-4 byte tables and x^(32*d*256^k) for shifting a raw CRC past d*256^k uint32 words.
struct CrcTables {
  u32 t[4][256];      // t[0] = plain byte table; t[k][i] = t[0] applied to byte i followed by k zero bytes
  u32 pw[4][256];     // pw[k][d] = x^(32 * d * 256^k) mod P   (pw[k][0] = x^0 = 0x80000000)
};

inline void crc_build_tables(CrcTables& T) {
  for (u32 i = 0; i < 256; i++) {
    u32 c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CKL_CRC_POLY : c >> 1;
    T.t[0][i] = c;
  }
  for (int k = 1; k < 4; k++)
    for (u32 i = 0; i < 256; i++) T.t[k][i] = (T.t[k - 1][i] >> 8) ^ T.t[0][T.t[k - 1][i] & 0xFF];
  // x^32: shift x^0 by 32 single-bit steps
  u32 x32 = 0x80000000u;
  for (int i = 0; i < 32; i++) x32 = (x32 >> 1) ^ (CKL_CRC_POLY & (0u - (x32 & 1u)));
  u32 base = x32;   // x^(32 * 256^k)
  for (int k = 0; k < 4; k++) {
    T.pw[k][0] = 0x80000000u;
    for (u32 d = 1; d < 256; d++) T.pw[k][d] = gf_mul(T.pw[k][d - 1], base);
    base = gf_mul(T.pw[k][255], base);
  }
}

// x^(32*n) mod P from the digit tables (host or device; `pw` may live in shared memory)
__host__ __device__ inline u32 gf_xpow32(const u32 (*pw)[256], u32 n) {
  u32 r = pw[0][n & 0xFF];
  u32 d1 = (n >> 8) & 0xFF, d2 = (n >> 16) & 0xFF, d3 = n >> 24;
  if (d1) r = gf_mul(r, pw[1][d1]);
  if (d2) r = gf_mul(r, pw[2][d2]);
  if (d3) r = gf_mul(r, pw[3][d3]);
  return r;
}

// one uint32 word through the raw CRC register (slicing-by-4)
__host__ __device__ inline u32 crc_word(const u32 (*t)[256], u32 crc, u32 v) {
  crc ^= v;
  return t[3][crc & 0xFF] ^ t[2][(crc >> 8) & 0xFF] ^ t[1][(crc >> 16) & 0xFF] ^ t[0][crc >> 24];
}
__host__ __device__ inline u32 crc_byte(const u32 (*t)[256], u32 crc, u8 v) {
  return t[0][(crc ^ v) & 0xFF] ^ (crc >> 8);
}

// host CRC-32C of a byte buffer (small sections: z-index; also used by tests of the tables)
inline u32 crc32c_host(const CrcTables& T, const u8* d, u64 n) {
  u32 c = 0xFFFFFFFFu;
  for (u64 i = 0; i < n; i++) c = crc_byte(T.t, c, d[i]);
  return ~c;
}
// finalise a RAW (zero-init, no final xor) register of a message of `nwords` uint32 words into the standard CRC
inline u32 crc_finalize_words(const CrcTables& T, u32 raw, u64 nwords) {
  // state_final = shift(0xFFFFFFFF, len) ^ raw ; crc = ~state_final.   nwords may exceed 2^32: split.
  u32 sh = 0x80000000u;
  u64 n = nwords;
  while (n) { u32 part = (u32)(n > 0xFFFFFFFFull ? 0xFFFFFFFFull : n); sh = gf_mul(sh, gf_xpow32(T.pw, part)); n -= part; }
  return ~(gf_mul(sh, 0xFFFFFFFFu) ^ raw);
}

// crc8 of the header: crc.hpp:23-37
inline u8 crc8_header(const u8* d, u64 n) {
  u8 c = 0xFF;
  while (n--) {
    c ^= *d++;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (u8)((c >> 1) ^ 0xE7) : (u8)(c >> 1);
  }
  return c;
}

// ---------------------------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__
#define FULL_MASK 0xFFFFFFFFu

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
  const u32 lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(FULL_MASK, v, o);
    if (lane >= (u32)o) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one u32 per thread.  `smem` needs 33 words.  Returns exclusive prefix and the
// block total via `total`.  Requires blockDim.x to be a multiple of 32 (<= 1024); all threads must call.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* smem, u32& total) {
  const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  u32 inc = warp_incl_scan(v);
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    u32 s = lane < nw ? smem[lane] : 0;
    u32 si = warp_incl_scan(s);
    smem[lane] = si - s;
    if (lane == 31) smem[32] = si;
  }
  __syncthreads();
  u32 r = smem[wid] + inc - v;
  total = smem[32];
  __syncthreads();
  return r;
}

// union-find over uint32 ids with "smaller id wins" (roots are component minima).  Loads go to L2 (other SMs
// update parents with atomics).
__device__ __forceinline__ u32 uf_find(const u32* par, u32 a) {
  u32 p;
  while ((p = __ldcg(par + a)) != a) a = p;
  return a;
}
__device__ __forceinline__ void uf_unite(u32* par, u32 a, u32 b) {
  for (;;) {
    a = uf_find(par, a);
    b = uf_find(par, b);
    if (a == b) return;
    if (a < b) { u32 t = a; a = b; b = t; }   // a > b: hang a under b
    u32 old = atomicMin(par + a, b);
    if (old == a) return;
    a = old;                                  // a was no longer a root; keep merging its old parent with b
  }
}
#endif
