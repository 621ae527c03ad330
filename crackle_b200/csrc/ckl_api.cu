// ckl_api.cu -- host orchestration and the C-ABI of libcrackle_b200.so.
//
// Stream assembly follows crackle::compress_helper (src/crackle.hpp:34-217) and the prologue of
// crackle::decompress (src/crackle.hpp:503-582); header emit/parse follows src/header.hpp:98-267.
// Error strings match the reference's std::runtime_error texts so a binding can re-raise them unchanged.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "ckl_internal.cuh"

void launch_exscan_u32_u64(const u32* in, u32 n, u32 stride, u64* out, ull* total_out, u64 add_each, cudaStream_t st);
void launch_markov_copy(const Geom& g, TraceBufs& T, MarkovBufs& M, u8* dst, cudaStream_t st);
void launch_markov_copy_body(const Geom& g, TraceBufs& T, MarkovBufs& M, u8* dst, cudaStream_t st);
void launch_reencode_codepoints(const Geom& g, const u8* stream, int order, DecodeBufs& D, u64 total_events, u32* sliceInfo, cudaStream_t st);
void launch_reencode_unpack(const Geom& g, DecodeBufs& D, const u32* sliceInfo, const u64* cpOff, u8* cp, u64 total_words, cudaStream_t st);
void launch_reencode_emit(const Geom& g, const u8* stream, DecodeBufs& D, const u32* sliceInfo, const u64* cpOff, const u8* cp,
                          const u64* codeOff, int pack0, u8* dst, cudaStream_t st);
bool launch_trace_nodes(const Geom& g, TraceBufs& T, ull* scal, u64 total_nodes, cudaStream_t st);
void launch_trace_paths(const Geom& g, TraceBufs& T, ull* scal, u32 max_nodes, cudaStream_t st);
void launch_trace_replay(const Geom& g, TraceBufs& T, ull* scal, u32 max_nodes, cudaStream_t st);
void launch_trace_post(const Geom& g, TraceBufs& T, ull* scal, u64 total_ev_cap, cudaStream_t st);
void launch_vcg(const Geom& g, const u32* DV, const u32* DH, int permissible, const u32* key, u8* out, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
struct ShardJob {
  bool active = false;
  Geom g{};
  const void* labels = nullptr;    // device pointer
  int width = 0;
  int permissible = 0, stored_width = 0, order = 0;
  u64 runs = 0, ncomp = 0, nuniq_local = 0, ncp = 0;
  u64 label_or = 0;                // OR of the shard's labels (same bit length as its largest label)
  u64 keys_bytes = 0, codes_bytes = 0;
  u64 codes_bytes0 = 0;            // order-0 code size of the shard, known at the end of the encode stage
  int key_width = 0;
  const u64* guniq = nullptr;      // global sorted unique table (device) the keys are written against
  u64 nuniq_global = 0;
  bool queued = false;             // encode stage queued, CCL / label results read back, tracer possibly still running
  bool encoded = false, finished = false;
};

unsigned long long g_ckl_launches = 0;
unsigned long long g_ckl_syncs = 0;
int g_ckl_grid_mult = 1;

int ckl_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}
u32 ckl_grid(u64 items, u32 per_block, u32 blocks_per_sm, bool fine) {
  u64 need = (items + per_block - 1) / per_block;
  const u64 cap = (u64)ckl_num_sms() * blocks_per_sm * (u64)(fine ? g_ckl_grid_mult : 1);
  if (need < 1) need = 1;
  return (u32)(need < cap ? need : cap);
}

// The chunk pipeline drives up to 2 streams per chunk; with the default of 8 hardware work queues the streams would
// alias and serialise each other.  Honoured only if the CUDA context is created after this library is loaded.
namespace { struct EnvInit { EnvInit() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); } } g_env_init; }

// optional per-stage timing with CUDA events on the context's stream
struct Prof {
  bool on = false;
  int id = 0;                    // 0 = the context itself, k + 1 = its k-th chunk context
  cudaEvent_t base = nullptr;    // CKL_TIMELINE=1: stage begin / end times relative to this event go to stderr
  struct Rec { std::string name; cudaEvent_t a, b; };
  std::vector<Rec> pending;
  std::vector<std::pair<std::string, double>> acc;   // name -> accumulated ms
  std::vector<u64> cnt;
  void begin(const char* name, cudaStream_t st) {
    if (!on) return;
    Rec r; r.name = name;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    pending.push_back(r);
  }
  void end(cudaStream_t st) { if (on && !pending.empty()) cudaEventRecord(pending.back().b, st); }
  void collect() {
    for (auto& r : pending) {
      cudaEventSynchronize(r.b);
      float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
      if (base) {
        float t0 = 0, t1 = 0;
        if (cudaEventElapsedTime(&t0, base, r.a) == cudaSuccess && cudaEventElapsedTime(&t1, base, r.b) == cudaSuccess)
          fprintf(stderr, "TL %d %-20s %9.3f %9.3f\n", id, r.name.c_str(), t0, t1);
        else cudaGetLastError();
      }
      size_t i = 0;
      for (; i < acc.size(); i++) if (acc[i].first == r.name) break;
      if (i == acc.size()) { acc.emplace_back(r.name, 0.0); cnt.push_back(0); }
      acc[i].second += ms; cnt[i]++;
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    pending.clear();
  }
};
#define STAGE(c, name, ...) do { (c)->prof.begin(name, (c)->st); __VA_ARGS__; (c)->prof.end((c)->st); } while (0)

// Staggered start of concurrent z-chunks: chunk k's decode stage starts (on the GPU) when chunk k-1's has finished, so the
// chunks run one stage apart -- a latency-bound stage of one beside a bandwidth- or issue-bound stage of another -- instead
// of in lockstep, where identical kernels only share the machine.
struct Stagger {
  std::vector<cudaEvent_t> ev;
  std::atomic<int> seq{0};           // number of chunks that have recorded their event (host-side ordering of record / wait)
};

struct ckl_ctx {
  Prof prof;
  Stagger* stg = nullptr;            // set on chunk contexts for the duration of a staggered call
  int stg_index = 0;
  int device = 0;
  cudaStream_t st = nullptr, own_st = nullptr;
  cudaStream_t st2 = nullptr;            // side stream: the tracing chain runs beside the CCL / label chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool ext_stream = false;
  std::string err;
  CrcTables htab;
  CrcTables* dtab = nullptr;
  ull* scal = nullptr;     // device scalars
  ull* hscal = nullptr;    // pinned mirror
  ull* hscal2 = nullptr;   // second mirror, read through the side stream
  DBuf labels_dev, DV, DH;
  CclBufs ccl;
  TraceBufs tr;
  MarkovBufs mk;
  LabelBufs lb;
  LabelBufs lb_merge;      // scratch of ckl_sort_unique_u64 (its `mapping` is borrowed from the caller per call)
  DecodeBufs dc;
  DBuf result; u64 result_bytes = 0;
  DBuf stream_dev, out_dev, tmp32, keys, codes;
  ShardJob job;
  // z-chunk pipelining (single-GPU compress / decompress of large volumes): child contexts, each with its own streams
  // and workspace, process disjoint z-ranges concurrently so the latency-bound stages of one chunk (chain replay,
  // decode chains) overlap the bandwidth-bound stages of the others (edge extraction, paint).
  std::vector<ckl_ctx*> kids;
  bool is_kid = false;
  int chunks = 0;                          // 0 = automatic, 1 = off, K = force K chunks
  cudaEvent_t ev_done = nullptr;
  struct HostStager* stager = nullptr;     // pinned staging ring for bulk copies from / to PAGEABLE host memory (lazy)
  bool stager_failed = false;              // the ring could not be allocated: pageable copies stay with the driver's staging
};

static void set_err(char* err, size_t n, const std::string& m) {
  if (err && n) { strncpy(err, m.c_str(), n - 1); err[n - 1] = 0; }
}

static void read_scalars(ckl_ctx* c) {
  CUDA_CHECK(cudaMemcpyAsync(c->hscal, c->scal, SC_COUNT * sizeof(ull), cudaMemcpyDeviceToHost, c->st));
  CUDA_CHECK(ckl_sync(c->st));
}

// ---------------------------------------------------------------------------------------------------------
// Bulk copies between PAGEABLE host memory and the device -- what a caller of the Python interface hands over (numpy
// arrays, bytes objects; the reference's crackle.compress / decompress take and return exactly those).  cudaMemcpyAsync on
// pageable memory is staged by the driver through one bounce buffer on the calling thread: measured on the bench box
// 10 GB/s up and 4 GB/s down (into untouched pages of a fresh output array) for a 8.6 GB volume, against 55 GB/s for pinned
// memory.  Here HS_THREADS host threads each move their share of the range through two pinned stages on their own stream:
// a thread copies one stage (first touch of the destination pages included, in parallel) while the DMA engine moves
// the other.  Pinned or registered memory, device pointers and small copies keep the plain asynchronous copy.
struct ChunkErr { int code = 0; std::string msg; };
#define HS_THREADS 8
#define HS_STAGE (4ull << 20)
#define HS_MIN_BYTES (16ull << 20)
struct HostStager {
  u8* pin = nullptr;
  cudaStream_t st[HS_THREADS] = {};
  cudaEvent_t ev[HS_THREADS][2] = {};
  u8* stage(int t, int s) const { return pin + ((u64)t * 2 + s) * HS_STAGE; }
  void create() {
    CUDA_CHECK(cudaMallocHost(&pin, (u64)HS_THREADS * 2 * HS_STAGE));
    for (int t = 0; t < HS_THREADS; t++) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking));
      for (int s = 0; s < 2; s++) CUDA_CHECK(cudaEventCreateWithFlags(&ev[t][s], cudaEventDisableTiming));
    }
  }
  ~HostStager() {
    for (int t = 0; t < HS_THREADS; t++) {
      for (int s = 0; s < 2; s++) if (ev[t][s]) cudaEventDestroy(ev[t][s]);
      if (st[t]) cudaStreamDestroy(st[t]);
    }
    if (pin) cudaFreeHost(pin);
  }
};
static bool host_ptr_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}
// host <-> device copy of `bytes` ordered after everything queued on `st`.  Pinned host memory / small sizes: queued on `st`
// like cudaMemcpyAsync.  Large pageable ranges: staged as described above; returns when the bytes have arrived.
static void copy_host(ckl_ctx* c, void* dst, const void* src, u64 bytes, bool to_device, cudaStream_t st) {
  if (!bytes) return;
  const void* hptr = to_device ? src : dst;
  if (bytes < HS_MIN_BYTES || !host_ptr_pageable(hptr)) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
    return;
  }
  if (!c->stager && !c->stager_failed) {
    HostStager* hs = new HostStager();
    try { hs->create(); c->stager = hs; }
    catch (const CklError&) { delete hs; cudaGetLastError(); c->stager_failed = true; }   // no pinned memory to be had: plain copies
  }
  if (!c->stager) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
    return;
  }
  HostStager& S = *c->stager;
  CUDA_CHECK(ckl_sync(st));               // down: the producer of the device buffer; up: its previous readers
  const u64 per = (((bytes + HS_THREADS - 1) / HS_THREADS) + 4095) & ~4095ull;
  ChunkErr errs[HS_THREADS];
  auto worker = [&](int t) {
    try {
      CUDA_CHECK(cudaSetDevice(c->device));
      const u64 lo = std::min(bytes, per * (u64)t), hi = std::min(bytes, lo + per);
      const u64 nb = (hi - lo + HS_STAGE - 1) / HS_STAGE;
      auto blk = [&](u64 i, u64& off, u64& n) { off = lo + i * HS_STAGE; n = std::min<u64>(HS_STAGE, hi - off); };
      u64 off, n;
      if (to_device) {
        for (u64 i = 0; i < nb; i++) {
          const int s = (int)(i & 1);
          blk(i, off, n);
          if (i >= 2) CUDA_CHECK(cudaEventSynchronize(S.ev[t][s]));      // the DMA out of this stage two blocks ago
          memcpy(S.stage(t, s), (const u8*)src + off, n);
          CUDA_CHECK(cudaMemcpyAsync((u8*)dst + off, S.stage(t, s), n, cudaMemcpyHostToDevice, S.st[t]));
          CUDA_CHECK(cudaEventRecord(S.ev[t][s], S.st[t]));
        }
      } else {
        auto issue = [&](u64 i) {
          u64 o, m;
          blk(i, o, m);
          CUDA_CHECK(cudaMemcpyAsync(S.stage(t, (int)(i & 1)), (const u8*)src + o, m, cudaMemcpyDeviceToHost, S.st[t]));
          CUDA_CHECK(cudaEventRecord(S.ev[t][i & 1], S.st[t]));
        };
        if (nb) issue(0);
        for (u64 i = 0; i < nb; i++) {
          if (i + 1 < nb) issue(i + 1);                                  // its stage was emptied in the previous round
          blk(i, off, n);
          CUDA_CHECK(cudaEventSynchronize(S.ev[t][i & 1]));
          memcpy((u8*)dst + off, S.stage(t, (int)(i & 1)), n);
        }
      }
      CUDA_CHECK(cudaStreamSynchronize(S.st[t]));
    } catch (const CklError& e) { errs[t].code = e.code; errs[t].msg = e.what(); cudaGetLastError(); }
    catch (const std::exception& e) { errs[t].code = CKL_ERR_CUDA; errs[t].msg = e.what(); }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < HS_THREADS; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  for (int t = 0; t < HS_THREADS; t++)
    if (errs[t].code) throw CklError(errs[t].code, errs[t].msg);
}

// ---------------------------------------------------------------------------------------------------------
// header (src/header.hpp)
static int ilog2w(int w) { return w == 1 ? 0 : w == 2 ? 1 : w == 4 ? 2 : 3; }
static void header_bytes_v1(u8* b, int data_width, int stored_width, int crack_format, int fortran, int order,
                            u32 sx, u32 sy, u32 sz, u64 num_label_bytes) {
  memcpy(b, "crkl", 4);
  b[4] = 1;
  const u32 fmt = (u32)ilog2w(data_width) | ((u32)ilog2w(stored_width) << 2) | ((u32)crack_format << 4) | (0u << 5) |
                  ((u32)(fortran ? 1 : 0) << 7) | (0u << 8) | (((u32)order & 15u) << 9) | (0u << 13);
  b[5] = (u8)fmt; b[6] = (u8)(fmt >> 8);
  for (int i = 0; i < 4; i++) { b[7 + i] = (u8)(sx >> (8 * i)); b[11 + i] = (u8)(sy >> (8 * i)); b[15 + i] = (u8)(sz >> (8 * i)); }
  b[19] = 31;                                                   // log2(grid_size = 2^31), crackle.hpp:87
  for (int i = 0; i < 8; i++) b[20 + i] = (u8)(num_label_bytes >> (8 * i));
  b[28] = crc8_header(b + 5, 23);
}
static u64 le_host(const u8* p, int w) { u64 v = 0; for (int i = 0; i < w; i++) v |= (u64)p[i] << (8 * i); return v; }

static int parse_header(const u8* b, u64 n, ckl_header_info* h, std::string& err) {
  if (n < 29) {   // crackle.hpp:513-517
    err = "crackle: Input too small to be a valid stream. Bytes: " + std::to_string(n);
    return CKL_ERR_STREAM;
  }
  if (memcmp(b, "crkl", 4) != 0 || b[4] > 1) {   // header.hpp:99-104
    err = "crackle: Data stream is not valid. Unable to decompress.";
    return CKL_ERR_STREAM;
  }
  h->format_version = b[4];
  const u32 fmt = (u32)le_host(b + 5, 2);
  h->sx = (u32)le_host(b + 7, 4); h->sy = (u32)le_host(b + 11, 4); h->sz = (u32)le_host(b + 15, 4);
  h->num_label_bytes = h->format_version == 0 ? le_host(b + 20, 4) : le_host(b + 20, 8);
  h->data_width = 1u << (fmt & 3); h->stored_data_width = 1u << ((fmt >> 2) & 3);
  h->crack_format = (fmt >> 4) & 1; h->label_format = (fmt >> 5) & 3; h->fortran_order = (fmt >> 7) & 1;
  h->is_signed = (fmt >> 8) & 1; h->markov_model_order = (fmt >> 9) & 15; h->is_sorted = !((fmt >> 13) & 1);
  if (h->format_version > 0 && crc8_header(b + 5, 23) != b[28]) {   // header.hpp:145-149
    err = "crackle: CRC8 check failed. Header may be corrupted. (~4.1% chance of a false positive for a single bit flip).";
    return CKL_ERR_STREAM;
  }
  return CKL_OK;
}
static u64 model_bytes_for(int order) {   // header.hpp:284-297
  if (order == 0) return 0;
  return (((u64)1 << (2 * order)) * 5 + 4) / 8;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int ckl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int ckl_ctx_create(int device, ckl_ctx** out) {
  if (!out) return CKL_ERR_ARG;
  *out = nullptr;
  if (ckl_device_count() <= device || device < 0) return CKL_ERR_CUDA;
  ckl_ctx* c = new ckl_ctx();
  try {
    c->device = device;
    CUDA_CHECK(cudaSetDevice(device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->own_st, cudaStreamNonBlocking));
    c->st = c->own_st;
    {
      // the tracing chain is the critical path of compress: its side stream gets the highest priority so its blocks
      // are scheduled ahead of the concurrent CCL / label kernels
      int lo = 0, hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, hi));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    crc_build_tables(c->htab);
    CUDA_CHECK(cudaMalloc(&c->dtab, sizeof(CrcTables)));
    CUDA_CHECK(cudaMemcpy(c->dtab, &c->htab, sizeof(CrcTables), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&c->scal, SC_COUNT * sizeof(ull)));
    CUDA_CHECK(cudaMallocHost(&c->hscal, SC_COUNT * sizeof(ull)));
    CUDA_CHECK(cudaMallocHost(&c->hscal2, SC_COUNT * sizeof(ull)));
  } catch (const CklError& e) {
    delete c;
    return e.code;
  }
  *out = c;
  return CKL_OK;
}

extern "C" void ckl_ctx_destroy(ckl_ctx* c) {
  if (!c) return;
  for (ckl_ctx* k : c->kids) ckl_ctx_destroy(k);
  c->kids.clear();
  cudaSetDevice(c->device);
  if (c->st) cudaStreamSynchronize(c->st);
  if (c->st2) { cudaStreamSynchronize(c->st2); cudaStreamDestroy(c->st2); }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->prof.base && !c->is_kid) cudaEventDestroy(c->prof.base);
  if (c->own_st) cudaStreamDestroy(c->own_st);
  if (c->dtab) cudaFree(c->dtab);
  if (c->scal) cudaFree(c->scal);
  if (c->hscal) cudaFreeHost(c->hscal);
  if (c->hscal2) cudaFreeHost(c->hscal2);
  delete c->stager;
  delete c;
}
extern "C" const char* ckl_ctx_error(const ckl_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" const char* crackle_b200_version(void) { return "crackle_b200 0.1 (sm_100a)"; }

#define API_BEGIN(c)                      \
  if (!(c)) return CKL_ERR_ARG;           \
  (c)->err.clear();                       \
  try {                                   \
    CUDA_CHECK(cudaSetDevice((c)->device));
#define API_END(c)                                                                     \
  }                                                                                    \
  catch (const CklError& e) { (c)->err = e.what(); (c)->job.active = false; if ((c)->st2) cudaStreamSynchronize((c)->st2); cudaGetLastError(); return e.code; } \
  catch (const std::exception& e) { (c)->err = e.what(); (c)->job.active = false; if ((c)->st2) cudaStreamSynchronize((c)->st2); return CKL_ERR_CUDA; }      \
  return CKL_OK;

// ---------------------------------------------------------------------------------------------------------
// sharded compress stages
static void shard_begin_impl(ckl_ctx* c, const void* labels, int on_device, int width, u64 sx, u64 sy, u64 sz, ckl_shard_summary* s) {
  if (width != 1 && width != 2 && width != 4 && width != 8) throw CklError(CKL_ERR_ARG, "crackle_b200: data_width must be 1, 2, 4 or 8");
  if (sx == 0 || sy == 0 || sz == 0) throw CklError(CKL_ERR_ARG, "crackle_b200: empty shard");
  if (sx > 0xFFFFFFFFull || sy > 0xFFFFFFFFull || sz > 0xFFFFFFFFull || sx * sy >= (1ull << 30))
    throw CklError(CKL_ERR_ARG, "crackle_b200: slice too large (sx*sy must be < 2^30)");
  ShardJob& J = c->job;
  J = ShardJob();
  J.g.sx = (u32)sx; J.g.sy = (u32)sy; J.g.sz = (u32)sz; J.g.W = (u32)((sx + 31) / 32); J.g.sxy = sx * sy;
  J.width = width;
  const u64 voxels = sx * sy * sz;
  if (on_device) J.labels = labels;
  else {
    c->labels_dev.ensure(voxels * (u64)width);
    if (c->stg) {                    // staggered chunk: the host->device copies of the chunks go one after the other (they share
                                     // PCIe anyway), so chunk k computes while chunk k+1 is still in flight
      while (c->stg->seq.load(std::memory_order_acquire) < c->stg_index) std::this_thread::yield();
      if (c->stg_index > 0) CUDA_CHECK(cudaStreamWaitEvent(c->st, c->stg->ev[c->stg_index - 1], 0));
    }
    copy_host(c, c->labels_dev.p, labels, voxels * (u64)width, true, c->st);
    if (c->stg) {
      CUDA_CHECK(cudaEventRecord(c->stg->ev[c->stg_index], c->st));
      int expect = c->stg_index;
      c->stg->seq.compare_exchange_strong(expect, c->stg_index + 1, std::memory_order_release);
    }
    J.labels = c->labels_dev.p;
  }
  const Geom& g = J.g;
  c->DV.ensure(g.words() * 4);
  c->DH.ensure(g.words() * 4);
  CUDA_CHECK(cudaMemsetAsync(c->scal, 0, SC_COUNT * sizeof(ull), c->st));
  STAGE(c, "edges", launch_edges(J.labels, width, g, c->DV.as<u32>(), c->DH.as<u32>(), c->scal, c->st));
  STAGE(c, "ccl_count", launch_ccl_count(g, c->DV.as<u32>(), c->ccl, c->scal, c->st));
  // first / last voxel of the shard (for the pixel pair straddling shard boundaries)
  u64 fl[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(&fl[0], J.labels, width, cudaMemcpyDeviceToHost, c->st));
  CUDA_CHECK(cudaMemcpyAsync(&fl[1], (const u8*)J.labels + (voxels - 1) * (u64)width, width, cudaMemcpyDeviceToHost, c->st));
  read_scalars(c);
  J.runs = c->hscal[SC_RUNS];
  s->max_label = c->hscal[SC_MAX];
  J.label_or = c->hscal[SC_MAX];
  s->pairs = c->hscal[SC_PAIRS];
  s->first_voxel = fl[0];
  s->last_voxel = fl[1];
  s->voxels = voxels;
  s->reserved[0] = s->reserved[1] = s->reserved[2] = 0;
  J.active = true;
}

// The encode stage in two halves.  shard_encode_queue queues both chains and returns once the CCL / label chain is through
// (component and unique-label counts known, the shard's sorted unique table ready) while the tracing chain may still be
// running on its side streams; shard_encode_join waits for the tracer and reads the code sizes.  A z-sharded caller puts its
// metadata and unique-table exchange between the two, beside the serial chain replay.
static void shard_encode_queue(ckl_ctx* c, int permissible, int stored_width, int order) {
  ShardJob& J = c->job;
  if (!J.active) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_encode without ckl_shard_begin");
  J.queued = J.encoded = J.finished = false;
  if (order < 0 || order > 12) throw CklError(CKL_ERR_ARG, "crackle_b200: markov_model_order must be in [0, 12]");
  J.permissible = permissible ? 1 : 0;
  J.stored_width = stored_width;
  J.order = order;
  const Geom& g = J.g;
  cudaStream_t st = c->st;
  // Two independent chains consume the planes: crack-code tracing (side stream) and CCL -> labels (main stream).
  cudaStream_t st2 = c->st2;
  CUDA_CHECK(cudaEventRecord(c->ev_fork, st));
  CUDA_CHECK(cudaStreamWaitEvent(st2, c->ev_fork, 0));
  c->prof.begin("trace_prepare", st2);
  launch_trace_prepare(g, c->DV.as<u32>(), c->DH.as<u32>(), J.permissible, c->tr, c->scal, st2);
  c->prof.end(st2);
  c->ccl.parent.ensure(J.runs * 4);
  c->ccl.runStart.ensure(J.runs * 4);
  c->ccl.compRank.ensure(J.runs * 4);
  c->ccl.compPix.ensure(J.runs * 4);
  // The tracing chain is the critical path and, up to the replay, as issue-bound as the CCL chain: it runs first and
  // alone.  The CCL / label chain is queued once the replay is -- that kernel keeps one warp per scheduler busy a fifth of
  // the time, and the CCL work fits into the issue slots it leaves free.
  CUDA_CHECK(cudaMemcpyAsync(c->hscal2, c->scal, SC_COUNT * sizeof(ull), cudaMemcpyDeviceToHost, st2));
  CUDA_CHECK(ckl_sync(st2));
  const u64 evCap = c->hscal2[SC_SYMCAP], stackCap = c->hscal2[SC_STACKCAP], chainCap = c->hscal2[SC_CHAINCAP], cpCap = c->hscal2[SC_CPCAP];
  const u64 nodes = c->hscal2[SC_NODES];
  const u32 maxNodes = (u32)c->hscal2[SC_MAXNODES];
  if (maxNodes >= (1u << 28)) throw CklError(CKL_ERR_ARG, "crackle_b200: slice has too many crack-graph nodes");
  c->tr.nodeVertex.ensure(nodes * 4 + 16);
  c->tr.nodeP.ensure(nodes * 4 + 16);
  c->tr.nodeAdj.ensure(nodes + 16);
  c->tr.seFar.ensure(nodes * 16 + 16);
  c->tr.seLen.ensure(nodes * 16 + 16);
  c->tr.ev.ensure(evCap * 4 + 16);
  c->tr.evRec.ensure(evCap * 16 + 16);
  c->tr.evCp.ensure((evCap + g.sz) * 4 + 16);
  c->tr.stack.ensure(stackCap * 8);
  c->tr.chain.ensure(chainCap * sizeof(ChainRec));
  c->tr.cp.ensure(cpCap);
  CUDA_CHECK(cudaMemsetAsync(c->tr.cp.p, 0, cpCap, st2));    // k_expand ORs the partial words at the ends of a super-edge into place
  c->prof.begin("trace_nodes", st2);
  const bool any_nodes = launch_trace_nodes(g, c->tr, c->scal, nodes, st2);
  c->prof.end(st2);
  if (any_nodes) {
    c->prof.begin("trace_paths", st2);
    launch_trace_paths(g, c->tr, c->scal, maxNodes, st2);
    c->prof.end(st2);
    c->prof.begin("trace_replay", st2);                    // k_replay alone: the serial, latency-bound kernel
    launch_trace_replay(g, c->tr, c->scal, maxNodes, st2);
    c->prof.end(st2);
  }
  c->prof.begin("trace_post", st2);
  launch_trace_post(g, c->tr, c->scal, evCap, st2);
  c->prof.end(st2);
  CUDA_CHECK(cudaEventRecord(c->ev_join, st2));
  // connected components, component ranks, crcs, component labels
  STAGE(c, "ccl_solve", launch_ccl_solve(g, c->DV.as<u32>(), c->DH.as<u32>(), c->ccl, c->dtab, c->scal, J.runs, st));
  read_scalars(c);
  J.ncomp = c->hscal[SC_COMPONENTS];
  const u32 init_term = gf_mul(gf_xpow32(c->htab.pw, (u32)g.sxy), 0xFFFFFFFFu);
  STAGE(c, "ccl_finish", launch_ccl_finish(g, c->ccl, J.runs, c->dtab, init_term, nullptr, st));
  c->lb.mapping.ensure(J.ncomp * 8 + 8);
  c->prof.begin("labels_sort_unique", st);
  launch_gather_mapping(J.labels, J.width, g, c->ccl, J.ncomp, c->lb.mapping.as<u64>(), st);
  // radix passes only over the bits the shard's labels have (40-bit ids in uint64 voxels: 5 passes instead of 8)
  int label_bits = 1;
  while (label_bits < 64 && (J.label_or >> label_bits)) label_bits++;
  labels_sort_unique(c->lb, J.ncomp, label_bits, st, &c->scal[SC_UNIQUE]);      // count picked up by the read-back below
  c->prof.end(st);
  J.queued = true;
}
static void shard_encode_join(ckl_ctx* c) {
  ShardJob& J = c->job;
  if (!J.queued) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_encode_wait without ckl_shard_encode_async");
  const Geom& g = J.g;
  cudaStream_t st = c->st;
  const int order = J.order;
  CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_join, 0));     // join: everything below sees the tracer's results
  launch_code_sizes_order0(g, c->tr, c->scal, st);        // order-0 code offsets / total: final unless a markov model re-codes them
  read_scalars(c);
  if (c->hscal[SC_ERROR]) throw CklError(CKL_ERR_CUDA, "crackle_b200: internal tracer capacity error " + std::to_string(c->hscal[SC_ERROR]));
  J.ncp = c->hscal[SC_CODEPOINTS];
  J.codes_bytes0 = c->hscal[SC_CODE_BYTES];
  J.nuniq_local = c->hscal[SC_UNIQUE];
  if (order > 0) {
    const u64 rows = 1ull << (2 * order);
    c->mk.stats.ensure(rows * 16);
    STAGE(c, "markov_stats", launch_markov_stats(g, c->tr, order, c->mk.stats.as<u32>(), st));
  }
  J.queued = false;
  J.encoded = true;
}
static void shard_encode_impl(ckl_ctx* c, int permissible, int stored_width, int order) {
  shard_encode_queue(c, permissible, stored_width, order);
  shard_encode_join(c);
}

// keys_dst / codes_dst: device destinations (may point into the final stream)
static void shard_finish_impl(ckl_ctx* c, const u64* guniq_dev, u64 nuniq_global, const u32* gstats_dev, int order,
                              bool materialise) {
  ShardJob& J = c->job;
  if (!J.encoded) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_finish without ckl_shard_encode");
  const Geom& g = J.g;
  cudaStream_t st = c->st;
  J.order = order;
  J.key_width = ckl_byte_width(nuniq_global);
  J.keys_bytes = J.ncomp * (u64)J.key_width;
  J.guniq = guniq_dev; J.nuniq_global = nuniq_global;
  if (materialise) {
    c->keys.ensure(J.keys_bytes + 8);
    launch_write_keys(c->lb.mapping.as<u64>(), J.ncomp, guniq_dev, nuniq_global, J.key_width, c->keys.as<u8>(), st);
  }
  if (order == 0) {                                      // sizes and offsets were scanned at the end of the encode stage
    J.codes_bytes = J.codes_bytes0;
    J.finished = true;
    if (!materialise) return;
    c->codes.ensure(J.codes_bytes + 8);
    launch_pack_order0(g, c->tr, c->codes.as<u8>(), st);
    return;
  }
  if (order > 0) {
    const u64 rows = 1ull << (2 * order);
    c->mk.model.ensure(rows * 4);
    c->mk.stored.ensure(model_bytes_for(order) + 16);
    launch_markov_model(order, gstats_dev, c->mk.model.as<u8>(), c->mk.stored.as<u8>(), model_bytes_for(order), st);
    launch_markov_sizes(g, c->tr, c->mk, c->scal, st);
    const u64 scratch_words = (3 * J.ncp) / 32 + 3ull * g.sz + 64;
    c->mk.scratch.ensure(scratch_words * 4);
    CUDA_CHECK(cudaMemsetAsync(c->mk.scratch.p, 0, scratch_words * 4, st));
    STAGE(c, "markov_encode", launch_markov_encode(g, c->tr, order, c->mk.model.as<u8>(), c->mk, st));
  }
  launch_code_sizes_order0(g, c->tr, c->scal, st);     // scans sliceInfo[.codeBytes] (either format)
  read_scalars(c);
  J.codes_bytes = c->hscal[SC_CODE_BYTES];
  J.finished = true;
  if (!materialise) return;                              // single-GPU path places keys and codes in the stream itself
  c->codes.ensure(J.codes_bytes + 8);
  if (order > 0) launch_markov_copy(g, c->tr, c->mk, c->codes.as<u8>(), st);
  else launch_pack_order0(g, c->tr, c->codes.as<u8>(), st);
}

__global__ void k_gather_stride4(const u32* __restrict__ src, u32 n, u32 off, u32* __restrict__ dst) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[(u64)i * 4 + off];
}
__global__ void k_store_bytes_u32(u8* dst, const u32* src) {
  const u32 v = *src;
  dst[0] = (u8)v; dst[1] = (u8)(v >> 8); dst[2] = (u8)(v >> 16); dst[3] = (u8)(v >> 24);
}

extern "C" int ckl_shard_begin(ckl_ctx* c, const void* labels, int labels_on_device, int data_width, uint64_t sx, uint64_t sy,
                               uint64_t sz_local, ckl_shard_summary* summary) {
  API_BEGIN(c)
  if (!summary) throw CklError(CKL_ERR_ARG, "crackle_b200: null summary");
  shard_begin_impl(c, labels, labels_on_device, data_width, sx, sy, sz_local, summary);
  API_END(c)
}
extern "C" int ckl_shard_encode(ckl_ctx* c, int permissible, int stored_width, int markov_model_order, uint64_t* n_unique_local,
                                uint64_t* n_components_local, uint64_t* n_codepoints_local) {
  API_BEGIN(c)
  shard_encode_impl(c, permissible, stored_width, markov_model_order);
  if (n_unique_local) *n_unique_local = c->job.nuniq_local;
  if (n_components_local) *n_components_local = c->job.ncomp;
  if (n_codepoints_local) *n_codepoints_local = c->job.ncp;
  API_END(c)
}
extern "C" int ckl_shard_encode_async(ckl_ctx* c, int permissible, int stored_width, int markov_model_order, uint64_t* n_unique_local,
                                      uint64_t* n_components_local) {
  API_BEGIN(c)
  shard_encode_queue(c, permissible, stored_width, markov_model_order);
  read_scalars(c);                                        // drains the main stream only: the CCL / label chain
  c->job.nuniq_local = c->hscal[SC_UNIQUE];
  if (n_unique_local) *n_unique_local = c->job.nuniq_local;
  if (n_components_local) *n_components_local = c->job.ncomp;
  API_END(c)
}
extern "C" int ckl_shard_encode_wait(ckl_ctx* c, uint64_t* n_codepoints_local, uint64_t* codes_bytes_order0) {
  API_BEGIN(c)
  shard_encode_join(c);
  if (n_codepoints_local) *n_codepoints_local = c->job.ncp;
  if (codes_bytes_order0) *codes_bytes_order0 = c->job.codes_bytes0;
  API_END(c)
}
extern "C" int ckl_shard_unique(ckl_ctx* c, uint64_t* dst, int dst_on_device) {
  API_BEGIN(c)
  if (!c->job.encoded && !c->job.queued) throw CklError(CKL_ERR_ARG, "crackle_b200: no encoded shard");
  if (c->job.nuniq_local) {
    CUDA_CHECK(cudaMemcpyAsync(dst, c->lb.uniq.p, c->job.nuniq_local * 8, dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->st));
    CUDA_CHECK(ckl_sync(c->st));
  }
  API_END(c)
}
extern "C" int ckl_shard_stats(ckl_ctx* c, uint32_t* dst, int dst_on_device) {
  API_BEGIN(c)
  if (!c->job.encoded || c->job.order <= 0) throw CklError(CKL_ERR_ARG, "crackle_b200: no markov statistics for this shard");
  const u64 rows = 1ull << (2 * c->job.order);
  CUDA_CHECK(cudaMemcpyAsync(dst, c->mk.stats.p, rows * 16, dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->st));
  CUDA_CHECK(ckl_sync(c->st));
  API_END(c)
}
extern "C" int ckl_shard_finish(ckl_ctx* c, const uint64_t* global_unique, int unique_on_device, uint64_t n_unique_global,
                                const uint32_t* global_stats, int stats_on_device, ckl_shard_pieces* pieces) {
  API_BEGIN(c)
  if (!pieces) throw CklError(CKL_ERR_ARG, "crackle_b200: null pieces");
  // the context keeps its own copy of the global table: ckl_shard_pack / ckl_shard_fetch write the keys against it later
  c->lb.sorted.ensure(n_unique_global * 8 + 8);
  if (n_unique_global)
    CUDA_CHECK(cudaMemcpyAsync(c->lb.sorted.p, global_unique, n_unique_global * 8,
                               unique_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->st));
  const u64* gu = c->lb.sorted.as<u64>();
  int order = global_stats ? c->job.order : 0;
  const u32* gs = global_stats;
  if (order > 0 && !stats_on_device) {
    const u64 rows = 1ull << (2 * order);
    CUDA_CHECK(cudaMemcpyAsync(c->mk.stats.p, global_stats, rows * 16, cudaMemcpyHostToDevice, c->st));
    gs = c->mk.stats.as<u32>();
  }
  shard_finish_impl(c, gu, n_unique_global, gs, order, true);
  pieces->keys_bytes = c->job.keys_bytes;
  pieces->codes_bytes = c->job.codes_bytes;
  pieces->sz_local = c->job.g.sz;
  API_END(c)
}
extern "C" int ckl_shard_model(ckl_ctx* c, uint8_t* dst, int dst_on_device, uint64_t capacity, uint64_t* model_bytes) {
  API_BEGIN(c)
  ShardJob& J = c->job;
  if (!J.finished) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_model without ckl_shard_finish");
  const u64 n = model_bytes_for(J.order);
  if (model_bytes) *model_bytes = n;
  if (dst && n) {
    if (capacity < n) throw CklError(CKL_ERR_ARG, "crackle_b200: model buffer too small");
    CUDA_CHECK(cudaMemcpyAsync(dst, c->mk.stored.p, n, dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->st));
    CUDA_CHECK(ckl_sync(c->st));
  }
  API_END(c)
}
extern "C" int ckl_shard_fetch(ckl_ctx* c, uint8_t* keys, uint64_t* components_per_slice, uint32_t* code_sizes, uint32_t* slice_crcs,
                               uint8_t* codes, int dst_on_device) {
  API_BEGIN(c)
  ShardJob& J = c->job;
  if (!J.finished) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_fetch without ckl_shard_finish");
  const cudaMemcpyKind k = dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const u32 sz = J.g.sz;
  if (keys && J.keys_bytes) CUDA_CHECK(cudaMemcpyAsync(keys, c->keys.p, J.keys_bytes, k, c->st));
  if (codes && J.codes_bytes) CUDA_CHECK(cudaMemcpyAsync(codes, c->codes.p, J.codes_bytes, k, c->st));
  if (slice_crcs) CUDA_CHECK(cudaMemcpyAsync(slice_crcs, c->ccl.sliceCrc.p, (u64)sz * 4, k, c->st));
  std::vector<u32> nz;
  if (components_per_slice) {
    if (dst_on_device) throw CklError(CKL_ERR_ARG, "crackle_b200: components_per_slice is host-only");
    nz.resize(sz);
    CUDA_CHECK(cudaMemcpyAsync(nz.data(), c->ccl.nz.p, (u64)sz * 4, cudaMemcpyDeviceToHost, c->st));
  }
  if (code_sizes) {
    c->tmp32.ensure((u64)sz * 4);
    k_gather_stride4<<<(sz + 255) / 256, 256, 0, c->st>>>(c->tr.sliceInfo.as<u32>(), sz, 3, c->tmp32.as<u32>());
    LAUNCH_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(code_sizes, c->tmp32.p, (u64)sz * 4, k, c->st));
  }
  CUDA_CHECK(ckl_sync(c->st));
  if (components_per_slice) for (u32 z = 0; z < sz; z++) components_per_slice[z] = nz[z];
  API_END(c)
}

extern "C" int ckl_shard_info(ckl_ctx* c, ckl_shard_counts* out) {
  API_BEGIN(c)
  if (!out) throw CklError(CKL_ERR_ARG, "crackle_b200: null output");
  const ShardJob& J = c->job;
  if (!J.encoded) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_info without ckl_shard_encode");
  out->n_unique_local = J.nuniq_local; out->n_components = J.ncomp; out->n_codepoints = J.ncp;
  out->codes_bytes_order0 = J.codes_bytes0; out->sz_local = J.g.sz; out->runs = J.runs;
  out->codes_bytes = J.finished ? J.codes_bytes : 0; out->keys_bytes = J.finished ? J.keys_bytes : 0;
  API_END(c)
}

// block layout of ckl_shard_pack (offsets from the block start; every section starts 4-byte aligned when the block does):
//   N_z u32[sz] | code sizes u32[sz] | slice crcs u32[sz] | keys[keys_bytes] | codes[codes_bytes]
static u64 shard_block_bytes(const ShardJob& J) { return 12ull * J.g.sz + J.keys_bytes + J.codes_bytes; }

extern "C" int ckl_shard_pack(ckl_ctx* c, uint8_t* dst, uint64_t capacity, uint64_t* bytes) {
  API_BEGIN(c)
  ShardJob& J = c->job;
  if (!J.finished) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_pack without ckl_shard_finish");
  const u64 need = shard_block_bytes(J);
  if (bytes) *bytes = need;
  if (dst) {
    if (capacity < need) throw CklError(CKL_ERR_ARG, "crackle_b200: pack buffer too small");
    if ((u64)dst & 3) throw CklError(CKL_ERR_ARG, "crackle_b200: pack buffer must be 4-byte aligned");
    const Geom& g = J.g;
    cudaStream_t st = c->st;
    u32* small = reinterpret_cast<u32*>(dst);
    CUDA_CHECK(cudaMemcpyAsync(small, c->ccl.nz.p, 4ull * g.sz, cudaMemcpyDeviceToDevice, st));
    k_gather_stride4<<<(g.sz + 255) / 256, 256, 0, st>>>(c->tr.sliceInfo.as<u32>(), g.sz, 3, small + g.sz);
    LAUNCH_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(small + 2ull * g.sz, c->ccl.sliceCrc.p, 4ull * g.sz, cudaMemcpyDeviceToDevice, st));
    u8* keys = dst + 12ull * g.sz;
    launch_write_keys(c->lb.mapping.as<u64>(), J.ncomp, J.guniq, J.nuniq_global, J.key_width, keys, st);
    u8* codes = keys + J.keys_bytes;
    if (J.order > 0) launch_markov_copy(g, c->tr, c->mk, codes, st);
    else launch_pack_order0(g, c->tr, codes, st);
  }
  API_END(c)
}

// small per-slice arrays of up to ASM_MAX_BLOCKS gathered blocks -> their place in the stream, one launch
#define ASM_MAX_BLOCKS 16
struct AsmParams {
  const u32* src[ASM_MAX_BLOCKS];     // block start: N_z | code sizes | crcs
  u32 sz[ASM_MAX_BLOCKS], z0[ASM_MAX_BLOCKS];
  int n, cw;
  u8 *dst_nz, *dst_z, *dst_crc;
};
__global__ void __launch_bounds__(256) k_assemble_small(AsmParams P) {
  for (int b = blockIdx.y; b < P.n; b += gridDim.y) {
    const u32 sz = P.sz[b];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * sz; i += gridDim.x * blockDim.x) {
      const u32 which = i / sz, z = i - which * sz;
      const u32 v = P.src[b][i];
      const u64 zz = (u64)P.z0[b] + z;
      u8* d = which == 0 ? P.dst_nz + zz * (u64)P.cw : (which == 1 ? P.dst_z : P.dst_crc) + zz * 4;
      const int w = which == 0 ? P.cw : 4;
      for (int k = 0; k < w; k++) d[k] = (u8)(v >> (8 * k));
    }
  }
}

extern "C" int ckl_shard_assemble(ckl_ctx* c, const uint8_t* gathered, const ckl_shard_block* blocks, int n_blocks,
                                  const uint64_t* global_unique, int unique_on_device, uint64_t nu, int data_width, int stored,
                                  int permissible, int fortran_order, int order, uint64_t sx, uint64_t sy, uint64_t* out_bytes) {
  API_BEGIN(c)
  if (!gathered || !blocks || n_blocks <= 0) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_assemble: no blocks");
  if (order < 0 || order > 12) throw CklError(CKL_ERR_ARG, "crackle_b200: markov_model_order must be in [0, 12]");
  cudaStream_t st = c->st;
  u64 sz = 0, ncomp = 0, codes_bytes = 0;
  for (int b = 0; b < n_blocks; b++) {
    if (blocks[b].offset & 3) throw CklError(CKL_ERR_ARG, "crackle_b200: block offsets must be 4-byte aligned");
    sz += blocks[b].sz_local; ncomp += blocks[b].n_components; codes_bytes += blocks[b].codes_bytes;
  }
  if (sz > 0xFFFFFFFFull) throw CklError(CKL_ERR_ARG, "crackle_b200: dimension exceeds uint32");
  const u64 sxy = sx * sy;
  const int kw = ckl_byte_width(nu), cw = ckl_byte_width(sxy);
  const u64 labels_bytes = 8 + nu * (u64)stored + sz * (u64)cw + ncomp * (u64)kw;
  const u64 off_z = 29, off_lab = off_z + 4ull * (sz + 1), off_model = off_lab + labels_bytes;
  const u64 off_nz = off_lab + 8 + nu * (u64)stored, off_keys = off_nz + sz * (u64)cw;
  const u64 off_codes = off_model + model_bytes_for(order);
  const u64 off_crcs = off_codes + codes_bytes + 4;
  const u64 total = off_crcs + 4ull * sz;
  c->result.ensure(total + 16);
  c->tmp32.ensure(64);
  u8* R = c->result.as<u8>();
  u8 hb[29];
  header_bytes_v1(hb, data_width, stored, permissible, fortran_order, order, (u32)sx, (u32)sy, (u32)sz, labels_bytes);
  CUDA_CHECK(cudaMemcpyAsync(R, hb, 29, cudaMemcpyHostToDevice, st));
  u64 z0 = 0, kpos = off_keys, cpos = off_codes;
  for (int b0 = 0; b0 < n_blocks; b0 += ASM_MAX_BLOCKS) {
    AsmParams P{};
    P.n = std::min(ASM_MAX_BLOCKS, n_blocks - b0); P.cw = cw;
    P.dst_nz = R + off_nz; P.dst_z = R + off_z; P.dst_crc = R + off_crcs;
    u32 maxsz = 1;
    for (int i = 0; i < P.n; i++) {
      const ckl_shard_block& B = blocks[b0 + i];
      if (B.keys_bytes != B.n_components * (u64)kw) throw CklError(CKL_ERR_ARG, "crackle_b200: block key bytes do not match the global key width");
      P.src[i] = reinterpret_cast<const u32*>(gathered + B.offset);
      P.sz[i] = (u32)B.sz_local; P.z0[i] = (u32)z0;
      maxsz = std::max(maxsz, (u32)B.sz_local);
      const u8* keys = gathered + B.offset + 12ull * B.sz_local;
      if (B.keys_bytes) CUDA_CHECK(cudaMemcpyAsync(R + kpos, keys, B.keys_bytes, cudaMemcpyDeviceToDevice, st));
      if (B.codes_bytes) CUDA_CHECK(cudaMemcpyAsync(R + cpos, keys + B.keys_bytes, B.codes_bytes, cudaMemcpyDeviceToDevice, st));
      z0 += B.sz_local; kpos += B.keys_bytes; cpos += B.codes_bytes;
    }
    k_assemble_small<<<dim3((3 * maxsz + 255) / 256, P.n), 256, 0, st>>>(P);
    LAUNCH_CHECK();
  }
  // z-index crc (crackle.hpp:173-185), unique table, markov model, labels crc (crackle.hpp:187, 211)
  u32* crc_tmp = c->tmp32.as<u32>();
  launch_crc_bytes(R + off_z, 4ull * sz, c->dtab, c->htab, crc_tmp, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_z + 4ull * sz, crc_tmp);
  LAUNCH_CHECK();
  u8 nub[8];
  for (int i = 0; i < 8; i++) nub[i] = (u8)(nu >> (8 * i));
  CUDA_CHECK(cudaMemcpyAsync(R + off_lab, nub, 8, cudaMemcpyHostToDevice, st));
  if (nu) {
    const u64* gu = global_unique;
    if (!unique_on_device) {
      c->lb.uniq.ensure(nu * 8 + 8);
      CUDA_CHECK(cudaMemcpyAsync(c->lb.uniq.p, global_unique, nu * 8, cudaMemcpyHostToDevice, st));
      gu = c->lb.uniq.as<u64>();
    }
    launch_write_uniq(gu, nu, stored, R + off_lab + 8, st);
  }
  if (order > 0) {
    if (!c->job.finished || c->job.order != order) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_shard_assemble needs this context's ckl_shard_finish for the markov model");
    CUDA_CHECK(cudaMemcpyAsync(R + off_model, c->mk.stored.p, model_bytes_for(order), cudaMemcpyDeviceToDevice, st));
  }
  launch_crc_bytes(R + off_lab, labels_bytes, c->dtab, c->htab, crc_tmp + 1, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_codes + codes_bytes, crc_tmp + 1);
  LAUNCH_CHECK();
  c->result_bytes = total;                     // queued on the context's stream; ckl_result_copy / ckl_decompress order after it
  if (out_bytes) *out_bytes = total;
  c->job.active = false;
  API_END(c)
}


// ---------------------------------------------------------------------------------------------------------
// z-chunk pipelining helpers

static void ensure_kids(ckl_ctx* c, int K) {
  while ((int)c->kids.size() < K) {
    ckl_ctx* k = nullptr;
    const int rc = ckl_ctx_create(c->device, &k);
    if (rc) throw CklError(rc, "crackle_b200: failed to create a chunk context");
    k->is_kid = true;
    k->chunks = 1;
    // stream priorities stagger the chunks: chunk 0 runs ahead, each later chunk fills what the earlier ones leave
    // idle (numerically lower = higher priority); inside a chunk the tracing chain outranks the CCL / label chain
    int lo = 0, hi = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int idx = (int)c->kids.size();
    const int p2 = std::min(hi + idx, lo), p1 = std::min(hi + idx + 1, lo);
    cudaStreamDestroy(k->st2);
    cudaStreamDestroy(k->own_st);
    k->st2 = k->own_st = k->st = nullptr;
    c->kids.push_back(k);          // owned from here on (destroyed with the parent even if stream creation fails)
    CUDA_CHECK(cudaStreamCreateWithPriority(&k->own_st, cudaStreamNonBlocking, p1));
    k->st = k->own_st;
    CUDA_CHECK(cudaStreamCreateWithPriority(&k->st2, cudaStreamNonBlocking, p2));
  }
}
struct GridMultScope {        // finer grids while chunks run concurrently (see g_ckl_grid_mult)
  int saved;
  GridMultScope() : saved(g_ckl_grid_mult) {
    g_ckl_grid_mult = 8;
  }
  ~GridMultScope() { g_ckl_grid_mult = saved; }
};
// chunk count for a volume of `sz` slices: explicit setting, CKL_CHUNKS, or automatic (large volumes only)
// Automatic policy: only when the volume crosses PCIe (host pointers) -- there the chunks overlap copies with compute.
// Device-resident volumes are not chunked: measured on B200 (tools/chunk_matrix.sh) the stages of different chunks compete
// for the same SM resources (shared memory of the replay, issue slots of the walkers) and the total does not shrink.
static int pick_chunks(const ckl_ctx* c, u64 sxy, u64 sz, u64 bytes_per_voxel, bool host_io) {
  if (c->is_kid) return 1;
  int K = c->chunks;
  if (K == 0) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("CKL_CHUNKS"); env = e ? atoi(e) : 0; if (env < 0) env = 0; }
    K = env;
  }
  if (K == 0) K = (host_io && sxy * sz * bytes_per_voxel >= (1ull << 28) && sz >= 64) ? 4 : 1;
  if ((u64)K > sz) K = (int)sz;
  if (K > 64) K = 64;
  return K < 1 ? 1 : K;
}
// run fn(k) for k in [0, K) on K host threads (each drives its own child context); first error in chunk order wins
static void run_chunks(ckl_ctx* c, int K, const std::function<void(int)>& fn) {
  std::vector<ChunkErr> errs(K);
  std::vector<std::thread> th;
  auto body = [&](int k) {
    try {
      CUDA_CHECK(cudaSetDevice(c->device));
      fn(k);
    } catch (const CklError& e) { errs[k].code = e.code; errs[k].msg = e.what(); cudaGetLastError(); }
    catch (const std::exception& e) { errs[k].code = CKL_ERR_CUDA; errs[k].msg = e.what(); }
  };
  for (int k = 1; k < K; k++) th.emplace_back(body, k);
  body(0);
  for (auto& t : th) t.join();
  for (int k = 0; k < K; k++)
    if (errs[k].code) {
      for (int j = 0; j < K; j++) { cudaStreamSynchronize(c->kids[j]->st); if (c->kids[j]->st2) cudaStreamSynchronize(c->kids[j]->st2); c->kids[j]->job.active = false; }
      throw CklError(errs[k].code, errs[k].msg);
    }
}
// children start after everything already queued on the parent's stream
static bool timeline_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CKL_TIMELINE"); v = (e && atoi(e) > 0) ? 1 : 0; }
  return v == 1;
}
static void timeline_base(ckl_ctx* c) {
  if (!c->prof.on || !timeline_on() || c->is_kid) return;
  if (!c->prof.base) CUDA_CHECK(cudaEventCreate(&c->prof.base));
  CUDA_CHECK(cudaEventRecord(c->prof.base, c->st));
  fprintf(stderr, "TL base\n");
}
static void fork_kids(ckl_ctx* c, int K) {
  CUDA_CHECK(cudaEventRecord(c->ev_fork, c->st));
  for (int k = 0; k < K; k++) {
    c->kids[k]->prof.on = c->prof.on;
    c->kids[k]->prof.id = k + 1;
    c->kids[k]->prof.base = c->prof.base;
    CUDA_CHECK(cudaStreamWaitEvent(c->kids[k]->st, c->ev_fork, 0));
  }
}
// the parent's stream continues after everything the children queued
static void join_kids(ckl_ctx* c, int K) {
  for (int k = 0; k < K; k++) {
    CUDA_CHECK(cudaEventRecord(c->kids[k]->ev_done, c->kids[k]->st));
    CUDA_CHECK(cudaStreamWaitEvent(c->st, c->kids[k]->ev_done, 0));
  }
}
static void merge_kid_prof(ckl_ctx* c, int K) {
  if (!c->prof.on) return;
  std::vector<std::pair<std::string, double>> sum;
  for (int k = 0; k < K; k++) {
    Prof& p = c->kids[k]->prof;
    p.collect();
    for (size_t i = 0; i < p.acc.size(); i++) {
      size_t j = 0;
      for (; j < sum.size(); j++) if (sum[j].first == p.acc[i].first) break;
      if (j == sum.size()) sum.emplace_back(p.acc[i].first, 0.0);
      sum[j].second += p.acc[i].second;
    }
    p.acc.clear(); p.cnt.clear();
  }
  for (auto& e : sum) {      // one call of the parent = the chunks' stage times added up (they overlap in wall time)
    size_t i = 0;
    for (; i < c->prof.acc.size(); i++) if (c->prof.acc[i].first == e.first) break;
    if (i == c->prof.acc.size()) { c->prof.acc.emplace_back(e.first, 0.0); c->prof.cnt.push_back(0); }
    c->prof.acc[i].second += e.second; c->prof.cnt[i]++;
  }
}

__global__ void k_add_u32(u32* __restrict__ dst, const u32* __restrict__ src, u64 n) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];                    // uint32 wrap == the reference's atomic<uint32> counters
}

// Single-GPU compress of a large volume as K z-chunks: each chunk runs the shard stages (edges -> CCL / tracing ->
// labels) on its own child context and streams; the chunks are then merged exactly like z-shards on different GPUs
// (global pixel-pair / max-label decisions, merged sorted unique table, summed markov statistics) and every chunk
// writes its keys, N_z, crack codes and CRCs straight into its place in the one output stream.
static void compress_chunked(ckl_ctx* c, const void* labels, int labels_on_device, int width, u64 sx, u64 sy, u64 sz, int fortran_order,
                             int order_in, int K, uint64_t* out_bytes) {
  ensure_kids(c, K);
  GridMultScope fine_grids;
  const u64 sxy = sx * sy, voxels = sxy * sz;
  std::vector<u64> z0(K + 1);
  for (int k = 0; k <= K; k++) z0[k] = sz * (u64)k / (u64)K;
  std::vector<ckl_shard_summary> sum(K);
  std::vector<int> guess(K);
  fork_kids(c, K);
  // phase A: everything that does not need a global decision.  The crack format (pixel pairs) is guessed from the
  // chunk's own statistics and verified below.  Host-resident input: the chunks' uploads are staggered.
  Stagger stg;
  if (!labels_on_device)
    for (int k = 0; k < K; k++) { stg.ev.push_back(c->kids[k]->ev_done); c->kids[k]->stg = &stg; c->kids[k]->stg_index = k; }
  try {
    run_chunks(c, K, [&](int k) {
      ckl_ctx* q = c->kids[k];
      const u8* src = (const u8*)labels + z0[k] * sxy * (u64)width;
      try {
        shard_begin_impl(q, src, labels_on_device, width, sx, sy, z0[k + 1] - z0[k], &sum[k]);
      } catch (...) { stg.seq.store(1 << 30, std::memory_order_release); throw; }      // never leave a later chunk spinning
      q->stg = nullptr;
      guess[k] = (i64)sum[k].pairs < (i64)sum[k].voxels / 2;
      shard_encode_impl(q, guess[k], ckl_byte_width(sum[k].max_label), order_in);
    });
  } catch (...) { for (int k = 0; k < K; k++) c->kids[k]->stg = nullptr; throw; }
  for (int k = 0; k < K; k++) c->kids[k]->stg = nullptr;
  u64 pairs = 0, maxl = 0;
  for (int k = 0; k < K; k++) {
    pairs += sum[k].pairs;
    if (k && sum[k].first_voxel == sum[k - 1].last_voxel) pairs++;      // the pair straddling the chunk boundary (lib.hpp:249-256)
    if (sum[k].max_label > maxl) maxl = sum[k].max_label;
  }
  const int permissible = (i64)pairs < (i64)voxels / 2;                   // crackle.hpp:50-55
  const int stored = ckl_byte_width(maxl);                               // crackle.hpp:233-235
  {
    std::vector<int> redo;
    for (int k = 0; k < K; k++) if (guess[k] != permissible) redo.push_back(k);
    if (!redo.empty())
      run_chunks(c, K, [&](int k) {
        if (guess[k] == permissible) return;
        ckl_ctx* q = c->kids[k];
        const u8* src = (const u8*)labels + z0[k] * sxy * (u64)width;
        shard_begin_impl(q, src, labels_on_device, width, sx, sy, z0[k + 1] - z0[k], &sum[k]);
        shard_encode_impl(q, permissible, stored, order_in);
      });
  }
  cudaStream_t st = c->st;
  join_kids(c, K);
  u64 ncomp = 0, ncp = 0, nloc = 0;
  std::vector<u64> compBase(K + 1), codeBase(K + 1);
  for (int k = 0; k < K; k++) { compBase[k] = ncomp; ncomp += c->kids[k]->job.ncomp; ncp += c->kids[k]->job.ncp; nloc += c->kids[k]->job.nuniq_local; }
  compBase[K] = ncomp;
  int order = order_in;
  if (order > 0 && ncp == 0) order = 0;                                  // crackle.hpp:107-118
  // global sorted unique label table = sort + unique of the concatenated per-chunk tables (labels.hpp:99-109)
  u64 nu = 0;
  if (nloc) {
    c->lb.mapping.ensure(nloc * 8 + 8);
    u64 o = 0;
    for (int k = 0; k < K; k++) {
      const u64 n = c->kids[k]->job.nuniq_local;
      if (n) CUDA_CHECK(cudaMemcpyAsync(c->lb.mapping.as<u64>() + o, c->kids[k]->lb.uniq.p, n * 8, cudaMemcpyDeviceToDevice, st));
      o += n;
    }
    nu = labels_sort_unique(c->lb, nloc, stored * 8, st);
  }
  if (order > 0) {                                                       // markov.hpp:193-220: counters summed over all slices
    const u64 cells = 4ull << (2 * order);
    c->mk.stats.ensure(cells * 4);
    CUDA_CHECK(cudaMemcpyAsync(c->mk.stats.p, c->kids[0]->mk.stats.p, cells * 4, cudaMemcpyDeviceToDevice, st));
    for (int k = 1; k < K; k++) {
      k_add_u32<<<(u32)((cells + 255) / 256), 256, 0, st>>>(c->mk.stats.as<u32>(), c->kids[k]->mk.stats.as<u32>(), cells);
      LAUNCH_CHECK();
    }
  }
  fork_kids(c, K);
  // phase B: keys against the global table, markov coding against the global model, code sizes
  run_chunks(c, K, [&](int k) {
    shard_finish_impl(c->kids[k], c->lb.uniq.as<u64>(), nu, order > 0 ? c->mk.stats.as<u32>() : nullptr, order, false);
  });
  u64 codes_bytes = 0;
  for (int k = 0; k < K; k++) { codeBase[k] = codes_bytes; codes_bytes += c->kids[k]->job.codes_bytes; }
  const int kw = ckl_byte_width(nu), cw = ckl_byte_width(sxy);
  const u64 labels_bytes = 8 + nu * (u64)stored + sz * (u64)cw + ncomp * (u64)kw;
  const u64 off_z = 29, off_lab = off_z + 4ull * (sz + 1), off_model = off_lab + labels_bytes;
  const u64 off_nz = off_lab + 8 + nu * (u64)stored, off_keys = off_nz + sz * (u64)cw;
  const u64 off_codes = off_model + model_bytes_for(order);
  const u64 off_crcs = off_codes + codes_bytes + 4;
  const u64 total = off_crcs + 4ull * sz;
  c->result.ensure(total + 16);
  c->tmp32.ensure(sz * 4 + 16);
  u8* R = c->result.as<u8>();
  // phase C: every chunk places its pieces (asynchronous launches on the chunk's own stream)
  for (int k = 0; k < K; k++) {
    ckl_ctx* q = c->kids[k];
    const Geom& g = q->job.g;
    cudaStream_t qs = q->st;
    k_gather_stride4<<<(g.sz + 255) / 256, 256, 0, qs>>>(q->tr.sliceInfo.as<u32>(), g.sz, 3, c->tmp32.as<u32>() + z0[k]);
    LAUNCH_CHECK();
    launch_write_le_u32(q->ccl.nz.as<u32>(), g.sz, cw, R + off_nz + z0[k] * (u64)cw, qs);
    launch_write_keys(q->lb.mapping.as<u64>(), q->job.ncomp, c->lb.uniq.as<u64>(), nu, kw, R + off_keys + compBase[k] * (u64)kw, qs);
    if (order > 0) launch_markov_copy(g, q->tr, q->mk, R + off_codes + codeBase[k], qs);
    else launch_pack_order0(g, q->tr, R + off_codes + codeBase[k], qs);
    launch_write_le_u32(q->ccl.sliceCrc.as<u32>(), g.sz, 4, R + off_crcs + 4ull * z0[k], qs);
  }
  join_kids(c, K);
  // header, z index + its crc (crackle.hpp:173-185), unique table, markov model, labels crc (crackle.hpp:187, 211)
  u8 hb[29];
  header_bytes_v1(hb, width, stored, permissible, fortran_order, order, (u32)sx, (u32)sy, (u32)sz, labels_bytes);
  CUDA_CHECK(cudaMemcpyAsync(R, hb, 29, cudaMemcpyHostToDevice, st));
  launch_write_le_u32(c->tmp32.as<u32>(), sz, 4, R + off_z, st);
  u32* crc_tmp = c->tmp32.as<u32>() + sz;
  launch_crc_bytes(R + off_z, 4ull * sz, c->dtab, c->htab, crc_tmp, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_z + 4ull * sz, crc_tmp);
  LAUNCH_CHECK();
  u8 nub[8];
  for (int i = 0; i < 8; i++) nub[i] = (u8)(nu >> (8 * i));
  CUDA_CHECK(cudaMemcpyAsync(R + off_lab, nub, 8, cudaMemcpyHostToDevice, st));
  launch_write_uniq(c->lb.uniq.as<u64>(), nu, stored, R + off_lab + 8, st);
  if (order > 0) CUDA_CHECK(cudaMemcpyAsync(R + off_model, c->kids[0]->mk.stored.p, model_bytes_for(order), cudaMemcpyDeviceToDevice, st));
  launch_crc_bytes(R + off_lab, labels_bytes, c->dtab, c->htab, crc_tmp + 1, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_codes + codes_bytes, crc_tmp + 1);
  LAUNCH_CHECK();
  CUDA_CHECK(ckl_sync(st));
  merge_kid_prof(c, K);
  c->prof.collect();
  for (int k = 0; k < K; k++) c->kids[k]->job.active = false;
  c->result_bytes = total;
  if (out_bytes) *out_bytes = total;
}

// ---------------------------------------------------------------------------------------------------------
// single-GPU compress = the three shard stages + stream assembly on the device
extern "C" int ckl_compress(ckl_ctx* c, const void* labels, int labels_on_device, int data_width, uint64_t sx, uint64_t sy, uint64_t sz,
                            int fortran_order, int markov_model_order, uint64_t* out_bytes) {
  API_BEGIN(c)
  if (data_width != 1 && data_width != 2 && data_width != 4 && data_width != 8)
    throw CklError(CKL_ERR_ARG, "crackle_b200: data_width must be 1, 2, 4 or 8");
  if (sx > 0xFFFFFFFFull || sy > 0xFFFFFFFFull || sz > 0xFFFFFFFFull) throw CklError(CKL_ERR_ARG, "crackle_b200: dimension exceeds uint32");
  if (markov_model_order < 0 || markov_model_order > 12) throw CklError(CKL_ERR_ARG, "crackle_b200: markov_model_order must be in [0, 12]");
  const u64 voxels = sx * sy * sz;
  cudaStream_t st = c->st;
  timeline_base(c);
  if (voxels == 0) {   // crackle.hpp:96-98: header only (crack format from pairs(0) < 0 == false)
    u8 hb[29];
    header_bytes_v1(hb, data_width, 1, 0, fortran_order, markov_model_order, (u32)sx, (u32)sy, (u32)sz, 0);
    c->result.ensure(29);
    CUDA_CHECK(cudaMemcpyAsync(c->result.p, hb, 29, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(ckl_sync(st));
    c->result_bytes = 29;
    if (out_bytes) *out_bytes = 29;
    return CKL_OK;
  }
  {
    if (sx * sy >= (1ull << 30)) throw CklError(CKL_ERR_ARG, "crackle_b200: slice too large (sx*sy must be < 2^30)");
    const int K = pick_chunks(c, sx * sy, sz, (u64)data_width, !labels_on_device);
    if (K > 1) { compress_chunked(c, labels, labels_on_device, data_width, sx, sy, sz, fortran_order, markov_model_order, K, out_bytes); return CKL_OK; }
  }
  ckl_shard_summary s;
  shard_begin_impl(c, labels, labels_on_device, data_width, sx, sy, sz, &s);
  const int permissible = (i64)s.pairs < (i64)voxels / 2;        // crackle.hpp:50-55
  const int stored = ckl_byte_width(s.max_label);                // crackle.hpp:233-235
  shard_encode_impl(c, permissible, stored, markov_model_order);
  ShardJob& J = c->job;
  int order = markov_model_order;
  if (order > 0 && J.ncp == 0) order = 0;                        // crackle.hpp:107-118
  const Geom& g = J.g;
  const u64 nu = J.nuniq_local;
  const int kw = ckl_byte_width(nu), cw = ckl_byte_width(g.sxy);
  const u64 labels_bytes = 8 + nu * (u64)stored + (u64)g.sz * cw + J.ncomp * (u64)kw;
  const u64 off_z = 29, off_lab = off_z + 4ull * (g.sz + 1), off_model = off_lab + labels_bytes;
  const u64 off_codes = off_model + model_bytes_for(order);
  // keys go straight into the stream; codes are placed once their total size is known
  const u64 off_keys = off_lab + 8 + nu * (u64)stored + (u64)g.sz * cw;
  shard_finish_impl(c, c->lb.uniq.as<u64>(), nu, order > 0 ? c->mk.stats.as<u32>() : nullptr, order, false);
  const u64 total = off_codes + J.codes_bytes + 4 + 4ull * g.sz;
  c->result.ensure(total + 16);
  u8* R = c->result.as<u8>();
  // header
  u8 hb[29];
  header_bytes_v1(hb, data_width, stored, permissible, fortran_order, order, g.sx, g.sy, g.sz, labels_bytes);
  CUDA_CHECK(cudaMemcpyAsync(R, hb, 29, cudaMemcpyHostToDevice, st));
  // z index: u32 code size per slice + crc32c of those bytes (crackle.hpp:173-185)
  c->tmp32.ensure((u64)g.sz * 4 + 16);
  k_gather_stride4<<<(g.sz + 255) / 256, 256, 0, st>>>(c->tr.sliceInfo.as<u32>(), g.sz, 3, c->tmp32.as<u32>());
  LAUNCH_CHECK();
  launch_write_le_u32(c->tmp32.as<u32>(), g.sz, 4, R + off_z, st);
  u32* crc_tmp = c->tmp32.as<u32>() + g.sz;
  launch_crc_bytes(R + off_z, 4ull * g.sz, c->dtab, c->htab, crc_tmp, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_z + 4ull * g.sz, crc_tmp);
  LAUNCH_CHECK();
  // labels section (labels.hpp:123-152)
  u8 nub[8];
  for (int i = 0; i < 8; i++) nub[i] = (u8)(nu >> (8 * i));
  CUDA_CHECK(cudaMemcpyAsync(R + off_lab, nub, 8, cudaMemcpyHostToDevice, st));
  launch_write_uniq(c->lb.uniq.as<u64>(), nu, stored, R + off_lab + 8, st);
  launch_write_le_u32(c->ccl.nz.as<u32>(), g.sz, cw, R + off_lab + 8 + nu * (u64)stored, st);
  launch_write_keys(c->lb.mapping.as<u64>(), J.ncomp, c->lb.uniq.as<u64>(), nu, kw, R + off_keys, st);
  // markov model + crack codes
  if (order > 0) {
    CUDA_CHECK(cudaMemcpyAsync(R + off_model, c->mk.stored.p, model_bytes_for(order), cudaMemcpyDeviceToDevice, st));
    launch_markov_copy(g, c->tr, c->mk, R + off_codes, st);
  } else {
    STAGE(c, "pack_order0", launch_pack_order0(g, c->tr, R + off_codes, st));
  }
  // trailing crcs (crackle.hpp:187, 211-214)
  launch_crc_bytes(R + off_lab, labels_bytes, c->dtab, c->htab, crc_tmp + 1, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_codes + J.codes_bytes, crc_tmp + 1);
  launch_write_le_u32(c->ccl.sliceCrc.as<u32>(), g.sz, 4, R + off_codes + J.codes_bytes + 4, st);
  LAUNCH_CHECK();
  // Device-resident input on a caller-owned stream (ckl_ctx_set_stream): the result is complete in stream order
  // (ckl_result_copy drains; a ckl_decompress of ckl_result_device() on this context is ordered behind it) and the caller
  // orders its own use of the input on that stream -- no drain here.  Otherwise the call returns with the work finished.
  if (!(labels_on_device && c->ext_stream) || c->prof.on) CUDA_CHECK(ckl_sync(st));
  c->prof.collect();
  c->result_bytes = total;
  if (out_bytes) *out_bytes = total;
  J.active = false;
  API_END(c)
}

extern "C" int ckl_result_copy(ckl_ctx* c, void* dst, int dst_on_device, uint64_t capacity) {
  API_BEGIN(c)
  if (capacity < c->result_bytes) throw CklError(CKL_ERR_ARG, "crackle_b200: result buffer too small");
  if (c->result_bytes) {
    if (dst_on_device) CUDA_CHECK(cudaMemcpyAsync(dst, c->result.p, c->result_bytes, cudaMemcpyDeviceToDevice, c->st));
    else copy_host(c, dst, c->result.p, c->result_bytes, false, c->st);
    CUDA_CHECK(ckl_sync(c->st));
  }
  API_END(c)
}
extern "C" const void* ckl_result_device(ckl_ctx* c, uint64_t* bytes) {
  if (!c) return nullptr;
  if (bytes) *bytes = c->result_bytes;
  return c->result.p;
}

// ---------------------------------------------------------------------------------------------------------
// decompress
__global__ void k_crc_compare(const u32* __restrict__ computed, const u8* __restrict__ stored, u32 n, ull* scal) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8* p = stored + (u64)i * 4;
  const u32 s = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
  if (s != computed[i]) atomicMin(&scal[SC_CRC_BAD], (ull)i);
}

// hbin: host view of the stream (may be null), dbin: device view (may be null -> uploaded); at least one is given.
static std::vector<u8> decode_stored_model(const u8* ms, u64 mbytes, int order);
// stats != nullptr: no paint -- the per-label tables of ckl_label_stats are produced instead (out / label are unused)
struct StatsOut { u64 *labels, *counts, *sums; u32* bbox; int on_device; u64 capacity; u64 n_unique; };
struct VcgOut { int connectivity; u8* out; int on_device; u64 capacity; };      // voxel connectivity graph instead of labels
static void decompress_impl(ckl_ctx* c, const u8* hbin, const u8* dbin, uint64_t num_bytes, int64_t z_start, int64_t z_end,
                            int has_label, uint64_t label, void* out, int out_on_device, uint64_t out_capacity,
                            StatsOut* stats = nullptr, VcgOut* vcg = nullptr) {
  cudaStream_t st = c->st;
  timeline_base(c);
  if (num_bytes < 29) throw CklError(CKL_ERR_STREAM, "crackle: Input too small to be a valid stream. Bytes: " + std::to_string(num_bytes));
  // the small sections are parsed on the host; a device-resident stream is read back piecewise (never as a whole)
  // (the first read brings the first 64 KB over in one go: header, z index and the label count of any ordinary stream)
  std::vector<u8> buf_head, buf_z, buf_lab, buf_nz, buf_model, prefix;
  auto fetch = [&](u64 offset, u64 n, std::vector<u8>& buf) -> const u8* {
    if (hbin) return hbin + offset;
    if (prefix.empty()) {
      prefix.resize((size_t)std::min<u64>(num_bytes, 65536));
      CUDA_CHECK(cudaMemcpyAsync(prefix.data(), dbin, prefix.size(), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(ckl_sync(st));
    }
    if (offset + n <= prefix.size()) return prefix.data() + offset;
    buf.resize(n ? n : 1);
    if (n) {
      CUDA_CHECK(cudaMemcpyAsync(buf.data(), dbin + offset, n, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(ckl_sync(st));
    }
    return buf.data();
  };
  const u8* hhead = fetch(0, 29, buf_head);
  ckl_header_info h;
  std::string perr;
  int rc = parse_header(hhead, num_bytes, &h, perr);
  if (rc) throw CklError(rc, perr);
  if (h.label_format != 0) throw CklError(CKL_ERR_UNSUPPORTED, "crackle_b200: pin label formats are outside the flat-label hot path; use the reference decoder");
  const i64 sz = (i64)h.sz;
  // crackle.hpp:527-537
  z_start = std::max(std::min(z_start, sz - 1), (i64)0);
  z_end = z_end < 0 ? sz : z_end;
  z_end = std::max(std::min(z_end, sz), (i64)0);
  if (z_start >= z_end) throw CklError(CKL_ERR_STREAM, "crackle: Invalid range: " + std::to_string(z_start) + " - " + std::to_string(z_end));
  const u64 szr = (u64)(z_end - z_start);
  const u64 sx = h.sx, sy = h.sy, sxy = sx * sy;
  const u64 voxels = sxy * szr;
  if (voxels == 0) return;
  if (sxy >= (1ull << 30)) throw CklError(CKL_ERR_ARG, "crackle_b200: slice too large (sx*sy must be < 2^30)");
  const int ow = has_label ? 1 : (int)h.data_width;
  if (!stats && !vcg && out_capacity < voxels * (u64)ow) throw CklError(CKL_ERR_ARG, "crackle_b200: output buffer too small");
  if (vcg && vcg->capacity < voxels) throw CklError(CKL_ERR_ARG, "crackle_b200: output buffer too small");
  const u64 hbytes = h.format_version == 0 ? 24 : 29;
  const u64 zbytes = 4ull * (h.sz + (h.format_version == 0 ? 0 : 1));
  if (hbytes + zbytes > num_bytes) throw CklError(CKL_ERR_STREAM, "crackle: get_crack_code_offsets: Unable to read past end of buffer.");
  const u8* hz = fetch(hbytes, zbytes, buf_z);                  // z index
  if (h.format_version > 0) {   // crackle.hpp:276-291
    const u32 stored = (u32)le_host(hz + 4ull * h.sz, 4);
    const u32 computed = crc32c_host(c->htab, hz, 4ull * h.sz);
    if (stored != computed)
      throw CklError(CKL_ERR_STREAM, "crackle: grid index crc32c did not match. stored: " + std::to_string(stored) + " computed: " + std::to_string(computed));
  }
  const int order = (int)h.markov_model_order;
  // the header field holds 0..15; the encoders (this one and, in practice, the reference: 4^order rows of statistics) stop far
  // below 13, where the model alone would be > 80 MB of stream and the 32-bit row arithmetic of the kernels ends
  if (order > 12) throw CklError(CKL_ERR_UNSUPPORTED, "crackle_b200: markov_model_order " + std::to_string(order) + " is not supported (maximum 12)");
  const u64 mbytes = model_bytes_for(order);
  std::vector<u64> off((u64)sz + 1);
  off[0] = hbytes + zbytes + h.num_label_bytes + mbytes;
  for (i64 z = 0; z < sz; z++) off[z + 1] = off[z] + le_host(hz + 4ull * z, 4);
  if (off[sz] > num_bytes) throw CklError(CKL_ERR_STREAM, "crackle: get_crack_codes: Unable to read past end of buffer.");
  if (h.format_version > 0 && off[sz] + 4 + 4ull * h.sz > num_bytes)
    throw CklError(CKL_ERR_STREAM, "crackle: get_crack_codes: Unable to read past end of buffer.");
  // labels section (labels.hpp:453-506)
  const u64 lab_off = hbytes + zbytes;
  if (h.num_label_bytes < 8) throw CklError(CKL_ERR_STREAM, "crackle: labels section too small.");
  const u64 nu = le_host(fetch(lab_off, 8, buf_lab), 8);
  const int sw = (int)h.stored_data_width, kw = ckl_byte_width(nu), cw = ckl_byte_width(sxy);
  const u64 uniq_off = lab_off + 8, nz_off = uniq_off + nu * (u64)sw, keys_off = nz_off + (u64)cw * h.sz;
  if (nu > h.num_label_bytes || keys_off > lab_off + h.num_label_bytes)
    throw CklError(CKL_ERR_STREAM, "crackle: labels section is inconsistent with the header.");
  const u64 n_keys = (lab_off + h.num_label_bytes - keys_off) / (u64)kw;
  const u8* hnz = fetch(nz_off, (u64)cw * h.sz, buf_nz);       // components per slice
  std::vector<u64> keyBase(szr), stackOff(szr + 1), codeOff(szr + 1);
  {
    u64 kb = 0;
    for (i64 z = 0; z < z_start; z++) kb += le_host(hnz + (u64)cw * z, cw);
    stackOff[0] = 0;
    for (u64 i = 0; i < szr; i++) {
      const u64 z = (u64)z_start + i;
      keyBase[i] = kb;
      kb += le_host(hnz + (u64)cw * z, cw);
      codeOff[i] = off[z];
      stackOff[i + 1] = stackOff[i] + 2 * (off[z + 1] - off[z]) + 4;
    }
    codeOff[szr] = off[z_end];
  }
  // markov model: symbol of rank per context row (markov.hpp:382-420)
  std::vector<u8> model;
  if (order > 0) model = decode_stored_model(fetch(lab_off + h.num_label_bytes, mbytes, buf_model), mbytes, order);
  // device copies
  const u8* dstream = dbin;
  if (!dstream) {
    c->stream_dev.ensure(num_bytes + 8);
    copy_host(c, c->stream_dev.p, hbin, num_bytes, true, st);
    dstream = c->stream_dev.as<u8>();
  }
  // Large Fortran-order outputs: K z-chunks on child contexts (each a z-range decode into its own part of the output),
  // so one chunk's decode chains / CCL overlap another chunk's paint and, for host outputs, its device->host copy.
  {
    const int K = (h.fortran_order && !stats && !vcg) ? pick_chunks(c, sxy, szr, (u64)ow, !out_on_device) : 1;
    if (K > 1) {
      ensure_kids(c, K);
      GridMultScope fine_grids;
      fork_kids(c, K);
      Stagger stg;
      for (int k = 0; k < K; k++) { stg.ev.push_back(c->kids[k]->ev_join); c->kids[k]->stg = &stg; c->kids[k]->stg_index = k; }
      try {
        run_chunks(c, K, [&](int k) {
          const i64 a = z_start + (i64)(szr * (u64)k / (u64)K), b = z_start + (i64)(szr * (u64)(k + 1) / (u64)K);
          const u64 o = (u64)(a - z_start) * sxy * (u64)ow;
          try {
            decompress_impl(c->kids[k], hbin, dstream, num_bytes, a, b, has_label, label, (u8*)out + o, out_on_device, (u64)(b - a) * sxy * (u64)ow);
          } catch (...) { stg.seq.store(1 << 30, std::memory_order_release); throw; }     // never leave a later chunk spinning
        });
      } catch (...) { for (int k = 0; k < K; k++) c->kids[k]->stg = nullptr; throw; }
      for (int k = 0; k < K; k++) c->kids[k]->stg = nullptr;
      join_kids(c, K);
      CUDA_CHECK(ckl_sync(st));
      merge_kid_prof(c, K);
      return;
    }
  }
  DecodeBufs& D = c->dc;
  D.codeOff.ensure((szr + 1) * 8); D.keyBase.ensure(szr * 8); D.stackOff.ensure((szr + 1) * 8);
  CUDA_CHECK(cudaMemcpyAsync(D.codeOff.p, codeOff.data(), (szr + 1) * 8, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(D.keyBase.p, keyBase.data(), szr * 8, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(D.stackOff.p, stackOff.data(), (szr + 1) * 8, cudaMemcpyHostToDevice, st));
  if (order > 0) {
    D.model.ensure(model.size());
    CUDA_CHECK(cudaMemcpyAsync(D.model.p, model.data(), model.size(), cudaMemcpyHostToDevice, st));
  }
  Geom g;
  g.sx = h.sx; g.sy = h.sy; g.sz = (u32)szr; g.W = (u32)((sx + 31) / 32); g.sxy = sxy;
  c->DV.ensure(g.words() * 4);
  c->DH.ensure(g.words() * 4);
  CUDA_CHECK(cudaMemsetAsync(c->scal, 0, SC_COUNT * sizeof(ull), st));
  const ull none = ~0ull;
  CUDA_CHECK(cudaMemcpyAsync(&c->scal[SC_CRC_BAD], &none, 8, cudaMemcpyHostToDevice, st));
  // per-slice word capacities for the scan-parallel decoder (from the code sizes; the descriptors are built on the device)
  std::vector<u64> wordOff(szr + 1);
  bool parallel_ok = true;
  {
    u64 tw = 0;
    for (u64 i = 0; i < szr; i++) {
      const u64 clen = codeOff[i + 1] - codeOff[i];
      const u64 ncp_cap = order == 0 ? clen * 4 : clen * 8;
      if (ncp_cap >= (1ull << 30)) parallel_ok = false;
      wordOff[i] = tw;
      tw += (ncp_cap + 15) / 16 + 1;
    }
    wordOff[szr] = tw;
  }
  const u64 total_words = wordOff[szr];
  if (c->stg) {                      // staggered chunk: wait (on the GPU) for the previous chunk's decode stage
    while (c->stg->seq.load(std::memory_order_acquire) < c->stg_index) std::this_thread::yield();
    if (c->stg_index > 0) CUDA_CHECK(cudaStreamWaitEvent(st, c->stg->ev[c->stg_index - 1], 0));
  }
  c->prof.begin("d_decode", st);
  bool decoded = false;
  if (parallel_ok) {
    D.slices.ensure(szr * sizeof(DecSlice));
    D.wordOff.ensure((szr + 1) * 8);
    CUDA_CHECK(cudaMemcpyAsync(D.wordOff.p, wordOff.data(), (szr + 1) * 8, cudaMemcpyHostToDevice, st));
    launch_decode_slices_init(g, dstream, D.codeOff.as<u64>(), D.wordOff.as<u64>(), D.slices.as<DecSlice>(), c->scal, st);
    launch_decode_classify(g, dstream, order, D.model.as<u8>(), D, total_words, c->scal, st);
    read_scalars(c);
    if (!c->hscal[SC_FIRST]) {     // no opposite-move run longer than a word (never produced by the encoder)
      launch_decode_mark(g, dstream, order, D, c->hscal[SC_LAST], total_words, c->DV.as<u32>(), c->DH.as<u32>(), c->scal, st);
      decoded = true;
    }
  }
  if (!decoded) {
    D.stack.ensure(stackOff[szr] * 4 + 16);
    launch_decode_slices(g, dstream, D.codeOff.as<u64>(), (int)h.crack_format, order, D.model.as<u8>(), c->DV.as<u32>(), c->DH.as<u32>(),
                         D.stack.as<u32>(), D.stackOff.as<u64>(), c->scal, st);
  }
  c->prof.end(st);
  if (c->stg) {
    CUDA_CHECK(cudaEventRecord(c->stg->ev[c->stg_index], st));
    int expect = c->stg_index;
    c->stg->seq.compare_exchange_strong(expect, c->stg_index + 1, std::memory_order_release);
  }
  launch_planes_from_cracks(g, (int)h.crack_format, c->DV.as<u32>(), c->DH.as<u32>(), st);
  STAGE(c, "d_ccl_count", launch_ccl_count(g, c->DV.as<u32>(), c->ccl, c->scal, st));
  read_scalars(c);
  if (c->hscal[SC_ERROR]) {
    if (h.crack_format) throw CklError(CKL_ERR_STREAM, "crackle: decode_permissible_crack_code: index out of range.");
    throw CklError(CKL_ERR_STREAM, "crackle: decode_impermissible_crack_code: index out of range.");
  }
  // voxel connectivity graph (operations.hpp:667-826): the 2-D bits come straight from the planes; connectivity 6 goes on
  // through the CCL / label stages for the keys of every voxel (key volume in out_dev, the graph bytes behind it)
  auto emit_vcg = [&](const u32* keyvol) {
    u8* dv = vcg->out;
    if (!vcg->on_device) {
      const u64 off = keyvol ? voxels * 4 : 0;             // the key volume (connectivity 6) lives in front of the graph bytes
      c->out_dev.ensure(off + voxels + 64);
      dv = c->out_dev.as<u8>() + off;
    }
    STAGE(c, "d_vcg", launch_vcg(g, c->DV.as<u32>(), c->DH.as<u32>(), (int)h.crack_format, keyvol, dv, st));
    if (!vcg->on_device) copy_host(c, vcg->out, dv, voxels, false, st);
  };
  if (vcg && (vcg->connectivity == 4 || h.sz == 1)) {     // operations.hpp:756-758
    emit_vcg(nullptr);
    read_scalars(c);
    if (!c->is_kid) c->prof.collect();
    return;
  }
  const u64 runs = c->hscal[SC_RUNS];
  c->ccl.parent.ensure(runs * 4); c->ccl.runStart.ensure(runs * 4); c->ccl.compRank.ensure(runs * 4);
  c->ccl.compPix.ensure(runs * 4);
  STAGE(c, "d_ccl_solve", launch_ccl_solve(g, c->DV.as<u32>(), c->DH.as<u32>(), c->ccl, c->dtab, c->scal, runs, st));
  const u32 init_term = gf_mul(gf_xpow32(c->htab.pw, (u32)g.sxy), 0xFFFFFFFFu);
  D.runLabel.ensure(runs * 8 + 8);
  CclDecodeSrc dsrc;
  dsrc.uniq = dstream + uniq_off; dsrc.keys = dstream + keys_off; dsrc.n_uniq = nu; dsrc.n_keys = n_keys; dsrc.sw = sw; dsrc.kw = kw;
  dsrc.keyBase = D.keyBase.as<u64>(); dsrc.runLabel = D.runLabel.as<u64>();
  D.uniq64.ensure(nu * 8 + 8);
  D.keys64.ensure(n_keys * 8 + 8);
  launch_unpack_le(dsrc.uniq, sw, nu, D.uniq64.as<u64>(), st);
  launch_unpack_le(dsrc.keys, kw, n_keys, D.keys64.as<u64>(), st);
  dsrc.uniq64 = D.uniq64.as<u64>(); dsrc.keys64 = D.keys64.as<u64>();
  dsrc.keys_only = stats != nullptr || vcg != nullptr;
  STAGE(c, "d_ccl_finish", launch_ccl_finish(g, c->ccl, runs, c->dtab, init_term, &dsrc, st));
  if (h.format_version > 0) {   // crackle.hpp:599-611
    k_crc_compare<<<(g.sz + 255) / 256, 256, 0, st>>>(c->ccl.sliceCrc.as<u32>(), dstream + num_bytes - 4ull * h.sz + 4ull * (u64)z_start, g.sz, c->scal);
    LAUNCH_CHECK();
  }
  // the verdict of the crc check is read with the final drain: the paint is queued behind it without a host round trip
  auto check_crc = [&]() {
    if (h.format_version == 0) return;
    if (c->hscal[SC_CRC_BAD] != none) {
      const u64 zi = c->hscal[SC_CRC_BAD];
      u32 computed = 0;
      CUDA_CHECK(cudaMemcpy(&computed, c->ccl.sliceCrc.as<u32>() + zi, 4, cudaMemcpyDeviceToHost));
      std::vector<u8> buf_crc;
      const u32 stored = (u32)le_host(fetch(num_bytes - 4ull * h.sz + 4ull * ((u64)z_start + zi), 4, buf_crc), 4);
      throw CklError(CKL_ERR_STREAM, "crackle: crack code crc mismatch on z=" + std::to_string((u64)z_start + zi) + " computed: " +
                                         std::to_string(computed) + " stored: " + std::to_string(stored));
    }
  };
  if (stats) {      // operations.hpp:321-665: per-label voxel counts, coordinate sums and bounding boxes, straight from the runs
    stats->n_unique = nu;
    if (stats->capacity < nu) throw CklError(CKL_ERR_ARG, "crackle_b200: statistics buffers hold fewer entries than the stream has labels");
    c->out_dev.ensure(nu * (8 + 24 + 24) + 64);
    ull* d_counts = c->out_dev.as<ull>();
    ull* d_sums = d_counts + nu;
    u32* d_bbox = reinterpret_cast<u32*>(d_sums + 3 * nu);
    STAGE(c, "d_stats", launch_run_stats(g, (u32)z_start, c->ccl, runs, D.runLabel.as<u64>(), nu, d_counts, d_sums, d_bbox, st));
    const cudaMemcpyKind k = stats->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (nu) {
      if (stats->labels) CUDA_CHECK(cudaMemcpyAsync(stats->labels, D.uniq64.p, nu * 8, k, st));
      if (stats->counts) CUDA_CHECK(cudaMemcpyAsync(stats->counts, d_counts, nu * 8, k, st));
      if (stats->sums) CUDA_CHECK(cudaMemcpyAsync(stats->sums, d_sums, nu * 24, k, st));
      if (stats->bbox) CUDA_CHECK(cudaMemcpyAsync(stats->bbox, d_bbox, nu * 24, k, st));
    }
    read_scalars(c);
    check_crc();
    if (!c->is_kid) c->prof.collect();
    return;
  }
  if (vcg) {        // connectivity 6: labels of vertically adjacent voxels agree <=> their unique-table keys do
    c->out_dev.ensure(voxels * 5 + 64);
    u32* keyvol = c->out_dev.as<u32>();
    STAGE(c, "d_paint", launch_paint(g, c->DV.as<u32>(), c->ccl, D.runLabel.as<u64>(), 4, 0, 0, 1, keyvol, st));
    emit_vcg(keyvol);
    read_scalars(c);
    check_crc();
    if (!c->is_kid) c->prof.collect();
    return;
  }
  void* dout = out;
  if (!out_on_device) { c->out_dev.ensure(voxels * (u64)ow); dout = c->out_dev.p; }
  STAGE(c, "d_paint", launch_paint(g, c->DV.as<u32>(), c->ccl, D.runLabel.as<u64>(), ow, has_label, label, (int)h.fortran_order, dout, st));
  if (!out_on_device) copy_host(c, out, dout, voxels * (u64)ow, false, st);
  read_scalars(c);                                       // the one drain at the end of the call
  check_crc();
  if (!c->is_kid) c->prof.collect();
}

extern "C" int ckl_decompress(ckl_ctx* c, const void* binary, int binary_on_device, uint64_t num_bytes, int64_t z_start, int64_t z_end,
                              int has_label, uint64_t label, void* out, int out_on_device, uint64_t out_capacity) {
  API_BEGIN(c)
  decompress_impl(c, binary_on_device ? nullptr : (const u8*)binary, binary_on_device ? (const u8*)binary : nullptr, num_bytes, z_start, z_end,
                  has_label, label, out, out_on_device, out_capacity);
  API_END(c)
}

extern "C" int ckl_voxel_connectivity_graph(ckl_ctx* c, const void* binary, int binary_on_device, uint64_t num_bytes, int64_t z_start,
                                            int64_t z_end, int connectivity, uint8_t* out, int out_on_device, uint64_t out_capacity) {
  API_BEGIN(c)
  if (connectivity != 4 && connectivity != 6)      // operations.hpp:675-679 (the reference appends the byte count to the text)
    throw CklError(CKL_ERR_ARG, "crackle: voxel_connectivity_graph: only connectivity 4 and 6 are currently supported." + std::to_string(num_bytes));
  VcgOut vo{connectivity, out, out_on_device, out_capacity};
  decompress_impl(c, binary_on_device ? nullptr : (const u8*)binary, binary_on_device ? (const u8*)binary : nullptr, num_bytes, z_start, z_end,
                  0, 0, nullptr, 1, 0, nullptr, &vo);
  API_END(c)
}

extern "C" int ckl_label_stats(ckl_ctx* c, const void* binary, int binary_on_device, uint64_t num_bytes, int64_t z_start, int64_t z_end,
                               uint64_t* labels, uint64_t* counts, uint64_t* sums, uint32_t* bbox, int out_on_device,
                               uint64_t capacity_entries, uint64_t* n_unique) {
  API_BEGIN(c)
  StatsOut so{labels, counts, sums, bbox, out_on_device, capacity_entries, 0};
  try {
    decompress_impl(c, binary_on_device ? nullptr : (const u8*)binary, binary_on_device ? (const u8*)binary : nullptr, num_bytes, z_start, z_end,
                    0, 0, nullptr, 1, 0, &so);
  } catch (...) { if (n_unique) *n_unique = so.n_unique; throw; }       // the label count is reported even when the buffers are too small
  if (n_unique) *n_unique = so.n_unique;
  API_END(c)
}

// symbol of rank per context row from the stored model (markov.hpp:382-420)
static std::vector<u8> decode_stored_model(const u8* ms, u64 mbytes, int order) {
  u8 lut[24]; int k = 0;
  for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) for (int cc = 0; cc < 4; cc++) for (int d = 0; d < 4; d++) {
    if (a == b || a == cc || a == d || b == cc || b == d || cc == d) continue;
    lut[k++] = (u8)(a | b << 2 | cc << 4 | d << 6);
  }
  const u64 rows = 1ull << (2 * order);
  std::vector<u8> model(rows * 4);
  for (u64 r = 0; r < rows; r++) {
    const u64 bit = r * 5;
    u32 v = ms[bit >> 3];
    if ((bit >> 3) + 1 < mbytes) v |= (u32)ms[(bit >> 3) + 1] << 8;
    const u8 packed = lut[((v >> (bit & 7)) & 31) % 24];
    for (int q = 0; q < 4; q++) model[r * 4 + q] = (packed >> (2 * q)) & 3;
  }
  return model;
}

// crackle::reencode_with_markov_order (src/crackle.hpp:860-984): same stream, crack codes re-coded with another context
// model order.  Labels section, labels crc and slice crcs are carried over verbatim (so every label format is accepted);
// the crack codes go codepoints -> (statistics -> model -> bitstreams | 2-bit packing) on the device.
extern "C" int ckl_reencode(ckl_ctx* c, const void* binary, int binary_on_device, uint64_t num_bytes, int new_order, uint64_t* out_bytes) {
  API_BEGIN(c)
  cudaStream_t st = c->st;
  const u8* hbin = binary_on_device ? nullptr : (const u8*)binary;
  const u8* dbin = binary_on_device ? (const u8*)binary : nullptr;
  if (num_bytes < 29) throw CklError(CKL_ERR_STREAM, "crackle: Input too small to be a valid stream. Bytes: " + std::to_string(num_bytes));
  if (new_order < 0 || new_order > 12) throw CklError(CKL_ERR_ARG, "crackle_b200: markov_model_order must be in [0, 12]");
  std::vector<u8> buf_head, buf_z, buf_model;
  auto fetch = [&](u64 offset, u64 n, std::vector<u8>& buf) -> const u8* {
    if (hbin) return hbin + offset;
    buf.resize(n ? n : 1);
    if (n) { CUDA_CHECK(cudaMemcpyAsync(buf.data(), dbin + offset, n, cudaMemcpyDeviceToHost, st)); CUDA_CHECK(ckl_sync(st)); }
    return buf.data();
  };
  const u8* hhead = fetch(0, 29, buf_head);
  ckl_header_info h;
  std::string perr;
  int rc = parse_header(hhead, num_bytes, &h, perr);
  if (rc) throw CklError(rc, perr);
  const int order = (int)h.markov_model_order;
  if (order > 12) throw CklError(CKL_ERR_UNSUPPORTED, "crackle_b200: markov_model_order " + std::to_string(order) + " is not supported (maximum 12)");
  const u8* dstream = dbin;
  if (!dstream) {
    c->stream_dev.ensure(num_bytes + 8);
    copy_host(c, c->stream_dev.p, hbin, num_bytes, true, st);
    dstream = c->stream_dev.as<u8>();
  }
  const u64 sxy = (u64)h.sx * h.sy;
  if (order == new_order || sxy * h.sz == 0) {                     // crackle.hpp:888-890: a copy
    c->result.ensure(num_bytes + 16);
    CUDA_CHECK(cudaMemcpyAsync(c->result.p, dstream, num_bytes, cudaMemcpyDeviceToDevice, st));
    CUDA_CHECK(ckl_sync(st));
    c->result_bytes = num_bytes;
    if (out_bytes) *out_bytes = num_bytes;
    return CKL_OK;
  }
  if (sxy >= (1ull << 30)) throw CklError(CKL_ERR_ARG, "crackle_b200: slice too large (sx*sy must be < 2^30)");
  const bool v1 = h.format_version > 0;
  const u64 sz = h.sz;
  const u64 hbytes = v1 ? 29 : 24, zbytes = 4ull * (sz + (v1 ? 1 : 0));
  if (hbytes + zbytes > num_bytes) throw CklError(CKL_ERR_STREAM, "crackle: get_crack_code_offsets: Unable to read past end of buffer.");
  const u8* hz = fetch(hbytes, zbytes, buf_z);
  const u64 mbytes = model_bytes_for(order), mbytes_new = model_bytes_for(new_order);
  std::vector<u64> codeOff(sz + 1), wordOff(sz + 1);
  codeOff[0] = hbytes + zbytes + h.num_label_bytes + mbytes;
  u64 tw = 0;
  for (u64 z = 0; z < sz; z++) {
    const u64 clen = le_host(hz + 4 * z, 4);
    codeOff[z + 1] = codeOff[z] + clen;
    const u64 cap = order == 0 ? clen * 4 : clen * 8;
    if (cap >= (1ull << 30)) throw CklError(CKL_ERR_ARG, "crackle_b200: crack code of a slice is too large");
    wordOff[z] = tw;
    tw += (cap + 15) / 16 + 1;
  }
  wordOff[sz] = tw;
  const u64 tail = v1 ? 4 + 4 * sz : 0;
  if (codeOff[sz] + tail > num_bytes) throw CklError(CKL_ERR_STREAM, "crackle: get_crack_codes: Unable to read past end of buffer.");
  Geom g;
  g.sx = h.sx; g.sy = h.sy; g.sz = (u32)sz; g.W = (u32)((g.sx + 31) / 32); g.sxy = sxy;
  DecodeBufs& D = c->dc;
  D.codeOff.ensure((sz + 1) * 8); D.wordOff.ensure((sz + 1) * 8); D.slices.ensure(sz * sizeof(DecSlice));
  CUDA_CHECK(cudaMemcpyAsync(D.codeOff.p, codeOff.data(), (sz + 1) * 8, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(D.wordOff.p, wordOff.data(), (sz + 1) * 8, cudaMemcpyHostToDevice, st));
  std::vector<u8> model;
  if (order > 0) {
    model = decode_stored_model(fetch(hbytes + zbytes + h.num_label_bytes, mbytes, buf_model), mbytes, order);
    D.model.ensure(model.size());
    CUDA_CHECK(cudaMemcpyAsync(D.model.p, model.data(), model.size(), cudaMemcpyHostToDevice, st));
  }
  CUDA_CHECK(cudaMemsetAsync(c->scal, 0, SC_COUNT * sizeof(ull), st));
  launch_decode_slices_init(g, dstream, D.codeOff.as<u64>(), D.wordOff.as<u64>(), D.slices.as<DecSlice>(), c->scal, st);
  launch_decode_classify(g, dstream, order, D.model.as<u8>(), D, tw, c->scal, st);
  read_scalars(c);
  if (c->hscal[SC_ERROR]) throw CklError(CKL_ERR_STREAM, "crackle: crack code index is larger than its slice's code.");
  if (c->hscal[SC_FIRST]) throw CklError(CKL_ERR_UNSUPPORTED, "crackle_b200: crack code holds an escape run no encoder emits; re-encode it with the reference");
  TraceBufs& T = c->tr;
  const u64 n1 = sz + 1;
  T.sliceInfo.ensure(sz * 16); T.offs.ensure(n1 * 8 * 4); T.codeOff.ensure(n1 * 8);
  launch_reencode_codepoints(g, dstream, order, D, c->hscal[SC_LAST], T.sliceInfo.as<u32>(), st);
  u64* cpOff = T.offs.as<u64>() + 3 * n1;
  launch_exscan_u32_u64(T.sliceInfo.as<u32>(), g.sz, 4, cpOff, &c->scal[SC_CODEPOINTS], 0, st);
  T.cp.ensure(tw * 16 + 16);                                         // upper bound of the codepoint total, known without a read-back
  launch_reencode_unpack(g, D, T.sliceInfo.as<u32>(), cpOff, T.cp.as<u8>(), tw, st);
  if (new_order > 0) {                                               // markov.hpp:166-266, :422-489
    const u64 rows = 1ull << (2 * new_order);
    c->mk.stats.ensure(rows * 16); c->mk.model.ensure(rows * 4); c->mk.stored.ensure(mbytes_new + 16);
    launch_markov_stats(g, T, new_order, c->mk.stats.as<u32>(), st);
    launch_markov_model(new_order, c->mk.stats.as<u32>(), c->mk.model.as<u8>(), c->mk.stored.as<u8>(), mbytes_new, st);
    launch_markov_sizes(g, T, c->mk, c->scal, st);
    const u64 scratch_words = (3 * tw * 16) / 32 + 3ull * g.sz + 64;
    c->mk.scratch.ensure(scratch_words * 4);
    CUDA_CHECK(cudaMemsetAsync(c->mk.scratch.p, 0, scratch_words * 4, st));
    launch_markov_encode(g, T, new_order, c->mk.model.as<u8>(), c->mk, st);
  }
  launch_code_sizes_order0(g, T, c->scal, st);
  read_scalars(c);
  const u64 codes_bytes = c->hscal[SC_CODE_BYTES];
  const u64 off_z = hbytes, off_lab = off_z + zbytes, off_model = off_lab + h.num_label_bytes, off_codes = off_model + mbytes_new;
  const u64 total = off_codes + codes_bytes + tail;
  c->result.ensure(total + 16);
  u8* R = c->result.as<u8>();
  u8 hb[29];
  memcpy(hb, hhead, hbytes);
  const u32 fmt = ((u32)le_host(hb + 5, 2) & ~(15u << 9)) | ((u32)new_order << 9);
  hb[5] = (u8)fmt; hb[6] = (u8)(fmt >> 8);
  if (v1) hb[28] = crc8_header(hb + 5, 23);
  CUDA_CHECK(cudaMemcpyAsync(R, hb, hbytes, cudaMemcpyHostToDevice, st));
  c->tmp32.ensure(sz * 4 + 16);
  k_gather_stride4<<<(g.sz + 255) / 256, 256, 0, st>>>(T.sliceInfo.as<u32>(), g.sz, 3, c->tmp32.as<u32>());
  LAUNCH_CHECK();
  launch_write_le_u32(c->tmp32.as<u32>(), sz, 4, R + off_z, st);
  if (v1) {
    u32* crc_tmp = c->tmp32.as<u32>() + sz;
    launch_crc_bytes(R + off_z, 4 * sz, c->dtab, c->htab, crc_tmp, st);
    k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_z + 4 * sz, crc_tmp);
    LAUNCH_CHECK();
  }
  CUDA_CHECK(cudaMemcpyAsync(R + off_lab, dstream + off_lab, h.num_label_bytes, cudaMemcpyDeviceToDevice, st));
  if (new_order > 0) CUDA_CHECK(cudaMemcpyAsync(R + off_model, c->mk.stored.p, mbytes_new, cudaMemcpyDeviceToDevice, st));
  launch_reencode_emit(g, dstream, D, T.sliceInfo.as<u32>(), cpOff, T.cp.as<u8>(), T.codeOff.as<u64>(), new_order == 0, R + off_codes, st);
  if (new_order > 0) launch_markov_copy_body(g, T, c->mk, R + off_codes, st);
  if (tail) CUDA_CHECK(cudaMemcpyAsync(R + off_codes + codes_bytes, dstream + codeOff[sz], tail, cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(ckl_sync(st));
  c->result_bytes = total;
  if (out_bytes) *out_bytes = total;
  API_END(c)
}

// ---------------------------------------------------------------------------------------------------------
// zstack / z-range extraction of streams without decoding (crackle/operations.py:424-548 zstack, :258-295
// _zstack_flat_labels, :551-662 zsplit / zshatter): label tables are merged / re-derived on the device (sort + unique,
// binary-search keys), crack codes, N_z, z-index entries and slice crcs are byte copies.
struct FlatStreamView {              // host-side description of one flat-label, order-0, version-1 stream
  ckl_header_info h;
  u64 nbytes, nu, n_keys, codes_bytes;
  u64 off_z, off_lab, off_uniq, off_nz, off_keys, off_codes, off_crcs;
  int sw, kw, cw;
  std::vector<u64> code_off;         // sz + 1 offsets relative to off_codes
  std::vector<u64> key_base;         // sz + 1 prefix of components per slice
  const u8* dev;                     // device view
};
static void ensure_copy(ckl_ctx* c, std::vector<u8>& dst, const u8* hbin, const u8* dbin, u64 off, u64 n) {
  dst.resize(n ? n : 1);
  if (!n) return;
  if (hbin) memcpy(dst.data(), hbin + off, n);
  else { CUDA_CHECK(cudaMemcpyAsync(dst.data(), dbin + off, n, cudaMemcpyDeviceToHost, c->st)); CUDA_CHECK(ckl_sync(c->st)); }
}
static FlatStreamView view_flat_stream(ckl_ctx* c, const u8* hbin, const u8* dbin, u64 nbytes, const char* who) {
  FlatStreamView v;
  v.nbytes = nbytes; v.dev = dbin;
  if (nbytes < 29) throw CklError(CKL_ERR_STREAM, "crackle: Input too small to be a valid stream. Bytes: " + std::to_string(nbytes));
  std::vector<u8> head, zi, small;
  ensure_copy(c, head, hbin, dbin, 0, 29);
  std::string perr;
  int rc = parse_header(head.data(), nbytes, &v.h, perr);
  if (rc) throw CklError(rc, perr);
  const ckl_header_info& h = v.h;
  if (h.format_version != 1) throw CklError(CKL_ERR_UNSUPPORTED, std::string("crackle_b200: ") + who + " needs format version 1 streams");
  if (h.label_format != 0) throw CklError(CKL_ERR_UNSUPPORTED, std::string("crackle_b200: ") + who + " covers the flat label format; use the reference for pins");
  if (h.markov_model_order != 0) throw CklError(CKL_ERR_ARG, std::string("crackle_b200: ") + who + " needs markov order 0 streams (ckl_reencode first, like operations.zstack does)");
  const u64 sz = h.sz, sxy = (u64)h.sx * h.sy;
  v.off_z = 29; v.off_lab = 29 + 4 * (sz + 1);
  if (v.off_lab + 8 > nbytes || h.num_label_bytes < 8) throw CklError(CKL_ERR_STREAM, "crackle: labels section too small.");
  ensure_copy(c, zi, hbin, dbin, v.off_z, 4 * sz);
  v.code_off.resize(sz + 1);
  v.code_off[0] = 0;
  for (u64 z = 0; z < sz; z++) v.code_off[z + 1] = v.code_off[z] + le_host(zi.data() + 4 * z, 4);
  v.codes_bytes = v.code_off[sz];
  ensure_copy(c, small, hbin, dbin, v.off_lab, 8);
  v.nu = le_host(small.data(), 8);
  v.sw = (int)h.stored_data_width; v.kw = ckl_byte_width(v.nu); v.cw = ckl_byte_width(sxy);
  v.off_uniq = v.off_lab + 8; v.off_nz = v.off_uniq + v.nu * (u64)v.sw; v.off_keys = v.off_nz + (u64)v.cw * sz;
  v.off_codes = v.off_lab + h.num_label_bytes; v.off_crcs = v.off_codes + v.codes_bytes + 4;
  if (v.nu > h.num_label_bytes || v.off_keys > v.off_codes || v.off_crcs + 4 * sz > nbytes)
    throw CklError(CKL_ERR_STREAM, "crackle: labels section is inconsistent with the header.");
  v.n_keys = (v.off_codes - v.off_keys) / (u64)v.kw;
  std::vector<u8> nz;
  ensure_copy(c, nz, hbin, dbin, v.off_nz, (u64)v.cw * sz);
  v.key_base.resize(sz + 1);
  v.key_base[0] = 0;
  for (u64 z = 0; z < sz; z++) v.key_base[z + 1] = v.key_base[z] + le_host(nz.data() + (u64)v.cw * z, v.cw);
  if (v.key_base[sz] > v.n_keys) throw CklError(CKL_ERR_STREAM, "crackle: labels section is inconsistent with the header.");
  return v;
}
__global__ void __launch_bounds__(256) k_keys_to_labels(const u64* __restrict__ keys64, u64 n, const u64* __restrict__ uniq64, u64 nu,
                                                         u64* __restrict__ out) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const u64 k = keys64[i]; out[i] = k < nu ? uniq64[k] : 0; }
}
// labels of components [k0, k1) of stream v into dst (u64 each)
static void stream_component_labels(ckl_ctx* c, const FlatStreamView& v, u64 k0, u64 k1, DBuf& uniq64, DBuf& keys64, u64* dst) {
  const u64 n = k1 - k0;
  if (!n) return;
  uniq64.ensure(v.nu * 8 + 8); keys64.ensure(n * 8 + 8);
  launch_unpack_le(v.dev + v.off_uniq, v.sw, v.nu, uniq64.as<u64>(), c->st);
  launch_unpack_le(v.dev + v.off_keys + k0 * (u64)v.kw, v.kw, n, keys64.as<u64>(), c->st);
  k_keys_to_labels<<<(u32)std::min<u64>((n + 255) / 256, 148 * 16), 256, 0, c->st>>>(keys64.as<u64>(), n, uniq64.as<u64>(), v.nu, dst);
  LAUNCH_CHECK();
}
// tail of a stream under construction: header, z-index crc, unique table, labels crc (everything else is already in place)
static void finish_flat_stream(ckl_ctx* c, u8* R, const ckl_header_info& h0, u32 sz, int data_width, int stored, u64 nu, const u64* guniq,
                               u64 labels_bytes, u64 off_codes, u64 codes_bytes) {
  cudaStream_t st = c->st;
  u8 hb[29];
  header_bytes_v1(hb, data_width, stored, (int)h0.crack_format, (int)h0.fortran_order, 0, h0.sx, h0.sy, sz, labels_bytes);
  u32 fmt = (u32)le_host(hb + 5, 2) | ((u32)h0.is_signed << 8) | ((u32)(h0.is_sorted ? 0 : 1) << 13);
  hb[5] = (u8)fmt; hb[6] = (u8)(fmt >> 8);
  hb[28] = crc8_header(hb + 5, 23);
  CUDA_CHECK(cudaMemcpyAsync(R, hb, 29, cudaMemcpyHostToDevice, st));
  const u64 off_z = 29, off_lab = off_z + 4ull * (sz + 1);
  c->tmp32.ensure(64);
  u32* crc_tmp = c->tmp32.as<u32>();
  launch_crc_bytes(R + off_z, 4ull * sz, c->dtab, c->htab, crc_tmp, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_z + 4ull * sz, crc_tmp);
  LAUNCH_CHECK();
  u8 nub[8];
  for (int i = 0; i < 8; i++) nub[i] = (u8)(nu >> (8 * i));
  CUDA_CHECK(cudaMemcpyAsync(R + off_lab, nub, 8, cudaMemcpyHostToDevice, st));
  launch_write_uniq(guniq, nu, stored, R + off_lab + 8, st);
  launch_crc_bytes(R + off_lab, labels_bytes, c->dtab, c->htab, crc_tmp + 1, st);
  k_store_bytes_u32<<<1, 1, 0, st>>>(R + off_codes + codes_bytes, crc_tmp + 1);
  LAUNCH_CHECK();
}
static u64 read_max_label(ckl_ctx* c, const u64* guniq, u64 nu) {
  u64 m = 0;
  if (nu) { CUDA_CHECK(cudaMemcpyAsync(&m, guniq + nu - 1, 8, cudaMemcpyDeviceToHost, c->st)); CUDA_CHECK(ckl_sync(c->st)); }
  return m;
}

extern "C" int ckl_zstack(ckl_ctx* c, int n, const void* const* binaries, const uint64_t* sizes, int on_device, uint64_t* out_bytes) {
  API_BEGIN(c)
  if (n <= 0 || !binaries || !sizes) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_zstack: no inputs");
  cudaStream_t st = c->st;
  std::vector<FlatStreamView> V;
  std::vector<DBuf> up(on_device ? 0 : n);
  u64 sz = 0, ncomp = 0, nloc = 0, codes = 0;
  int data_width = 1;
  for (int i = 0; i < n; i++) {
    const u8* hb = on_device ? nullptr : (const u8*)binaries[i];
    const u8* db = on_device ? (const u8*)binaries[i] : nullptr;
    if (!on_device) {
      up[i].ensure(sizes[i] + 8);
      CUDA_CHECK(cudaMemcpyAsync(up[i].p, hb, sizes[i], cudaMemcpyHostToDevice, st));
      db = up[i].as<u8>();
    }
    V.push_back(view_flat_stream(c, hb, db, sizes[i], "zstack"));
    V.back().dev = db;
    const ckl_header_info &h = V.back().h, &f = V[0].h;
    if (h.sx != f.sx || h.sy != f.sy)                         // operations.py:471-475
      throw CklError(CKL_ERR_ARG, "All images must have the same width and height. Expected sx=" + std::to_string(f.sx) + " sy=" +
                                      std::to_string(f.sy) + " ; Got: sx=" + std::to_string(h.sx) + " sy=" + std::to_string(h.sy));
    if (h.crack_format != f.crack_format) throw CklError(CKL_ERR_ARG, "All crack formats must match.");
    if (h.is_signed != f.is_signed) throw CklError(CKL_ERR_ARG, "All binaries must have the same sign.");
    data_width = std::max(data_width, (int)h.data_width);
    sz += h.sz; ncomp += V.back().key_base[h.sz]; nloc += V.back().nu; codes += V.back().codes_bytes;
  }
  if (sz > 0xFFFFFFFFull) throw CklError(CKL_ERR_ARG, "crackle_b200: dimension exceeds uint32");
  // merged sorted unique table (operations.py:494-497)
  c->lb.mapping.ensure(std::max(nloc, ncomp) * 8 + 8);
  u64 o = 0;
  for (auto& v : V) { launch_unpack_le(v.dev + v.off_uniq, v.sw, v.nu, c->lb.mapping.as<u64>() + o, st); o += v.nu; }
  const u64 nu = labels_sort_unique(c->lb, nloc, 64, st);
  const u64* guniq = c->lb.uniq.as<u64>();
  const int stored = ckl_byte_width(read_max_label(c, guniq, nu));
  const int kw = ckl_byte_width(nu), cw = V[0].cw;
  const u64 labels_bytes = 8 + nu * (u64)stored + sz * (u64)cw + ncomp * (u64)kw;
  const u64 off_z = 29, off_lab = off_z + 4 * (sz + 1), off_nz = off_lab + 8 + nu * (u64)stored, off_keys = off_nz + sz * (u64)cw;
  const u64 off_codes = off_lab + labels_bytes, off_crcs = off_codes + codes + 4, total = off_crcs + 4 * sz;
  c->result.ensure(total + 16);
  u8* R = c->result.as<u8>();
  u64 z0 = 0, k0 = 0, c0 = 0;
  for (auto& v : V) {
    const u64 szi = v.h.sz, nk = v.key_base[szi];
    if (szi) {
      CUDA_CHECK(cudaMemcpyAsync(R + off_z + 4 * z0, v.dev + v.off_z, 4 * szi, cudaMemcpyDeviceToDevice, st));
      CUDA_CHECK(cudaMemcpyAsync(R + off_nz + z0 * (u64)cw, v.dev + v.off_nz, szi * (u64)cw, cudaMemcpyDeviceToDevice, st));
      CUDA_CHECK(cudaMemcpyAsync(R + off_crcs + 4 * z0, v.dev + v.off_crcs, 4 * szi, cudaMemcpyDeviceToDevice, st));
    }
    if (v.codes_bytes) CUDA_CHECK(cudaMemcpyAsync(R + off_codes + c0, v.dev + v.off_codes, v.codes_bytes, cudaMemcpyDeviceToDevice, st));
    if (nk) {
      stream_component_labels(c, v, 0, nk, c->dc.uniq64, c->dc.keys64, c->lb.mapping.as<u64>());
      launch_write_keys(c->lb.mapping.as<u64>(), nk, guniq, nu, kw, R + off_keys + k0 * (u64)kw, st);
    }
    z0 += szi; k0 += nk; c0 += v.codes_bytes;
  }
  finish_flat_stream(c, R, V[0].h, (u32)sz, data_width, stored, nu, guniq, labels_bytes, off_codes, codes);
  CUDA_CHECK(ckl_sync(st));
  c->result_bytes = total;
  if (out_bytes) *out_bytes = total;
  API_END(c)
}

// the stream of slices [z_start, z_end) of `binary`: its own sorted unique table and keys (operations.py:551-615 _zsplit_helper)
extern "C" int ckl_zslice(ckl_ctx* c, const void* binary, int on_device, uint64_t num_bytes, uint64_t z_start, uint64_t z_end,
                          uint64_t* out_bytes) {
  API_BEGIN(c)
  cudaStream_t st = c->st;
  const u8* hb = on_device ? nullptr : (const u8*)binary;
  const u8* db = on_device ? (const u8*)binary : nullptr;
  if (!on_device) {
    c->stream_dev.ensure(num_bytes + 8);
    CUDA_CHECK(cudaMemcpyAsync(c->stream_dev.p, hb, num_bytes, cudaMemcpyHostToDevice, st));
    db = c->stream_dev.as<u8>();
  }
  FlatStreamView v = view_flat_stream(c, hb, db, num_bytes, "zsplit");
  v.dev = db;
  if (z_start >= z_end || z_end > v.h.sz) throw CklError(CKL_ERR_ARG, "crackle_b200: ckl_zslice: z-range outside the stream");
  const u64 sz = z_end - z_start, k0 = v.key_base[z_start], k1 = v.key_base[z_end], nk = k1 - k0;
  const u64 codes = v.code_off[z_end] - v.code_off[z_start];
  c->lb.mapping.ensure(nk * 8 + 8);
  DBuf& labs = c->dc.runLabel;                                 // component labels of the range (kept while mapping is sorted)
  labs.ensure(nk * 8 + 8);
  stream_component_labels(c, v, k0, k1, c->dc.uniq64, c->dc.keys64, labs.as<u64>());
  if (nk) CUDA_CHECK(cudaMemcpyAsync(c->lb.mapping.p, labs.p, nk * 8, cudaMemcpyDeviceToDevice, st));
  const u64 nu = labels_sort_unique(c->lb, nk, 64, st);
  const u64* guniq = c->lb.uniq.as<u64>();
  const int stored = ckl_byte_width(read_max_label(c, guniq, nu));
  const int kw = ckl_byte_width(nu), cw = v.cw;
  const u64 labels_bytes = 8 + nu * (u64)stored + sz * (u64)cw + nk * (u64)kw;
  const u64 off_z = 29, off_lab = off_z + 4 * (sz + 1), off_nz = off_lab + 8 + nu * (u64)stored, off_keys = off_nz + sz * (u64)cw;
  const u64 off_codes = off_lab + labels_bytes, off_crcs = off_codes + codes + 4, total = off_crcs + 4 * sz;
  c->result.ensure(total + 16);
  u8* R = c->result.as<u8>();
  CUDA_CHECK(cudaMemcpyAsync(R + off_z, v.dev + v.off_z + 4 * z_start, 4 * sz, cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(R + off_nz, v.dev + v.off_nz + z_start * (u64)cw, sz * (u64)cw, cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(R + off_crcs, v.dev + v.off_crcs + 4 * z_start, 4 * sz, cudaMemcpyDeviceToDevice, st));
  if (codes) CUDA_CHECK(cudaMemcpyAsync(R + off_codes, v.dev + v.off_codes + v.code_off[z_start], codes, cudaMemcpyDeviceToDevice, st));
  launch_write_keys(labs.as<u64>(), nk, guniq, nu, kw, R + off_keys, st);
  finish_flat_stream(c, R, v.h, (u32)sz, (int)v.h.data_width, stored, nu, guniq, labels_bytes, off_codes, codes);
  CUDA_CHECK(ckl_sync(st));
  c->result_bytes = total;
  if (out_bytes) *out_bytes = total;
  API_END(c)
}

// z-chunk pipelining of the single-GPU paths: 0 = automatic (large volumes), 1 = off, K = always K chunks
extern "C" int ckl_ctx_set_chunks(ckl_ctx* c, int chunks) {
  if (!c || chunks < 0) return CKL_ERR_ARG;
  c->chunks = chunks;
  return CKL_OK;
}

// ---------------------------------------------------------------------------------------------------------
// instrumentation and small device utilities used by the multi-GPU host code
extern "C" int ckl_prof_enable(ckl_ctx* c, int on) {
  if (!c) return CKL_ERR_ARG;
  c->prof.on = on != 0;
  c->prof.acc.clear(); c->prof.cnt.clear();
  return CKL_OK;
}
extern "C" int ckl_prof_read(ckl_ctx* c, char* buf, size_t cap) {
  if (!c || !buf || !cap) return CKL_ERR_ARG;
  c->prof.collect();
  std::string s;
  for (size_t i = 0; i < c->prof.acc.size(); i++)
    s += c->prof.acc[i].first + "=" + std::to_string(c->prof.acc[i].second) + ":" + std::to_string(c->prof.cnt[i]) + ";";
  strncpy(buf, s.c_str(), cap - 1);
  buf[cap - 1] = 0;
  return CKL_OK;
}
extern "C" uint64_t ckl_launch_count(void) { return g_ckl_launches; }
extern "C" uint64_t ckl_sync_count(void) { return g_ckl_syncs; }
// Run this context's work on a caller-owned stream (e.g. torch's current stream) so the caller's stream order and
// CUDA events cover the kernels.  stream == 0 is the legacy default stream (what torch uses unless told otherwise).
extern "C" int ckl_ctx_set_stream(ckl_ctx* c, void* stream) {
  if (!c) return CKL_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->st);
  c->st = (cudaStream_t)stream;
  c->ext_stream = true;
  return CKL_OK;
}
extern "C" int ckl_ctx_own_stream(ckl_ctx* c) {
  if (!c) return CKL_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->st);
  c->st = c->own_st;
  c->ext_stream = false;
  return CKL_OK;
}

extern "C" int ckl_crc32c(ckl_ctx* c, const void* data, int on_device, uint64_t n, uint32_t* out) {
  API_BEGIN(c)
  if (!out) throw CklError(CKL_ERR_ARG, "crackle_b200: null output");
  const u8* d = (const u8*)data;
  if (!on_device) {
    c->stream_dev.ensure(n + 8);
    CUDA_CHECK(cudaMemcpyAsync(c->stream_dev.p, data, n, cudaMemcpyHostToDevice, c->st));
    d = c->stream_dev.as<u8>();
  }
  c->tmp32.ensure(16);
  launch_crc_bytes(d, n, c->dtab, c->htab, c->tmp32.as<u32>(), c->st);
  CUDA_CHECK(cudaMemcpyAsync(out, c->tmp32.p, 4, cudaMemcpyDeviceToHost, c->st));
  CUDA_CHECK(ckl_sync(c->st));
  API_END(c)
}

// in-place sort + unique of a device array of uint64 (the merge step of the global label table)
extern "C" int ckl_sort_unique_u64(ckl_ctx* c, uint64_t* data_device, uint64_t n, int key_bytes, uint64_t* n_unique) {
  API_BEGIN(c)
  if (!n_unique) throw CklError(CKL_ERR_ARG, "crackle_b200: null output");
  LabelBufs& tmp = c->lb_merge;                               // context-owned scratch: steady-state calls do no cudaMalloc
  tmp.mapping.p = data_device; tmp.mapping.cap = n * 8;      // borrowed, not owned
  u64 cnt = 0;
  try {
    cnt = labels_sort_unique(tmp, n, key_bytes * 8, c->st);
    if (cnt) CUDA_CHECK(cudaMemcpyAsync(data_device, tmp.uniq.p, cnt * 8, cudaMemcpyDeviceToDevice, c->st));
    CUDA_CHECK(ckl_sync(c->st));
  } catch (...) { tmp.mapping.p = nullptr; tmp.mapping.cap = 0; throw; }
  tmp.mapping.p = nullptr; tmp.mapping.cap = 0;
  *n_unique = cnt;
  API_END(c)
}

// ---------------------------------------------------------------------------------------------------------
// one-shot host API on a lazily created default context (device = current CUDA device)
static std::mutex g_mu;
static ckl_ctx* g_default = nullptr;
static ckl_ctx* default_ctx(std::string& err, int& code) {
  if (!g_default) {
    int dev = 0;
    if (ckl_device_count() <= 0) { err = "crackle_b200: no CUDA device available (there is no CPU fallback)"; code = CKL_ERR_CUDA; return nullptr; }
    cudaGetDevice(&dev);
    code = ckl_ctx_create(dev, &g_default);
    if (code) { err = "crackle_b200: failed to create a CUDA context"; return nullptr; }
  }
  return g_default;
}

extern "C" int crackle_b200_compress(const void* labels, int data_width, uint64_t sx, uint64_t sy, uint64_t sz, int fortran_order,
                                     int markov_model_order, uint8_t** out, uint64_t* out_bytes, char* err, size_t err_len) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out || !out_bytes) { set_err(err, err_len, "crackle_b200: null output pointer"); return CKL_ERR_ARG; }
  *out = nullptr; *out_bytes = 0;
  std::string e; int code = 0;
  ckl_ctx* c = default_ctx(e, code);
  if (!c) { set_err(err, err_len, e); return code; }
  u64 n = 0;
  code = ckl_compress(c, labels, 0, data_width, sx, sy, sz, fortran_order, markov_model_order, &n);
  if (code) { set_err(err, err_len, c->err); return code; }
  u8* buf = (u8*)malloc(n ? n : 1);
  if (!buf) { set_err(err, err_len, "crackle_b200: out of host memory"); return CKL_ERR_NOMEM; }
  code = ckl_result_copy(c, buf, 0, n);
  if (code) { free(buf); set_err(err, err_len, c->err); return code; }
  *out = buf; *out_bytes = n;
  return CKL_OK;
}

extern "C" int crackle_b200_decompress(const uint8_t* binary, uint64_t num_bytes, int64_t z_start, int64_t z_end, int has_label,
                                       uint64_t label, void* out, uint64_t out_capacity, char* err, size_t err_len) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::string e; int code = 0;
  ckl_ctx* c = default_ctx(e, code);
  if (!c) { set_err(err, err_len, e); return code; }
  code = ckl_decompress(c, binary, 0, num_bytes, z_start, z_end, has_label, label, out, 0, out_capacity);
  if (code) set_err(err, err_len, c->err);
  return code;
}

extern "C" void crackle_b200_free(void* p) { free(p); }

extern "C" int crackle_b200_header(const uint8_t* binary, uint64_t num_bytes, ckl_header_info* info, char* err, size_t err_len) {
  if (!binary || !info) { set_err(err, err_len, "crackle_b200: null argument"); return CKL_ERR_ARG; }
  std::string e;
  const int rc = parse_header(binary, num_bytes, info, e);
  if (rc) set_err(err, err_len, e);
  return rc;
}
