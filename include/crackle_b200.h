/*
 * crackle_b200.h -- C-ABI of the B200-native crackle hot path (libcrackle_b200.so).
 *
 * Plain pointers and sizes only; no C++ / torch types cross this boundary.  These entry points are what the
 * reference's pybind11 boundary (src/fastcrackle.cpp) binds for the per-z-slice path:
 *
 *   crackle_b200_compress    replaces  fastcrackle.compress   (src/fastcrackle.cpp:131-210  ->
 *                                      crackle::compress<LABEL> src/crackle.hpp:220-257), flat labels only
 *   crackle_b200_decompress  replaces  fastcrackle.decompress (src/fastcrackle.cpp:41-129   ->
 *                                      crackle::decompress<LABEL,OUT> src/crackle.hpp:503-663)
 *   crackle_b200_free        mirrors the malloc/free contract of the reference's own C-ABI precedent
 *                            (wasm/crackle_wasm.cc:20-67: crackle_compress / crackle_decompress)
 *
 * The ckl_ctx_* / ckl_* functions are the same operations on an explicit per-GPU context with device-resident
 * buffers (what a z-sharded multi-GPU caller or a benchmark uses so volumes never leave HBM).
 *
 * All functions return 0 on success and a non-zero code on failure; the message text follows the reference's
 * std::runtime_error strings ("crackle: ...") so a binding can re-raise them unchanged.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with CKL_ERR_CUDA.
 */
#ifndef CRACKLE_B200_H
#define CRACKLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CKL_API __attribute__((visibility("default")))
#else
#define CKL_API
#endif

enum {
  CKL_OK = 0,
  CKL_ERR_CUDA = 1,          /* no device / CUDA runtime failure                                     */
  CKL_ERR_ARG = 2,           /* bad argument (unsupported width, sizes, signed input, ...)           */
  CKL_ERR_STREAM = 3,        /* invalid / corrupt .ckl stream (message matches the reference's text) */
  CKL_ERR_UNSUPPORTED = 4,   /* valid stream outside this path (pin label formats): use the reference */
  CKL_ERR_NOMEM = 5
};

typedef struct ckl_ctx ckl_ctx;

/* Parsed .ckl header (src/header.hpp:35-308). */
typedef struct ckl_header_info {
  uint32_t format_version, label_format, crack_format, is_signed;
  uint32_t data_width, stored_data_width, fortran_order, markov_model_order;
  uint32_t sx, sy, sz, is_sorted;
  uint64_t num_label_bytes;
} ckl_header_info;

/* ---- one-shot host API: the drop-in boundary ------------------------------------------------------------ */

/* labels: HOST pointer to sx*sy*sz unsigned integers of `data_width` bytes in Fortran order (x fastest).
 * fortran_order: header flag only (src/crackle.hpp:90).  On success *out is a malloc'ed buffer of *out_bytes
 * (release with crackle_b200_free).  err/err_len: optional message buffer. */
CKL_API int crackle_b200_compress(const void* labels, int data_width, uint64_t sx, uint64_t sy, uint64_t sz,
                          int fortran_order, int markov_model_order, uint8_t** out, uint64_t* out_bytes,
                          char* err, size_t err_len);

/* Decodes slices [z_start, z_end) (clamped exactly like src/crackle.hpp:527-537; z_end < 0 means sz) into the
 * HOST buffer `out` of out_capacity bytes: data_width-byte labels, or 1 byte per voxel (label mask) when
 * has_label != 0.  Output memory order follows the stream's fortran_order flag (src/crackle.hpp:617-656). */
CKL_API int crackle_b200_decompress(const uint8_t* binary, uint64_t num_bytes, int64_t z_start, int64_t z_end,
                            int has_label, uint64_t label, void* out, uint64_t out_capacity,
                            char* err, size_t err_len);

CKL_API void crackle_b200_free(void* p);

/* Host-only header parse (no GPU needed); validates magic, version and crc8 like src/header.hpp:98-150. */
CKL_API int crackle_b200_header(const uint8_t* binary, uint64_t num_bytes, ckl_header_info* info, char* err, size_t err_len);

/* ---- explicit context API (device-resident buffers) ------------------------------------------------------ */

CKL_API int ckl_ctx_create(int device, ckl_ctx** ctx);
CKL_API void ckl_ctx_destroy(ckl_ctx* ctx);
CKL_API const char* ckl_ctx_error(const ckl_ctx* ctx);           /* message of the last failure on this context */
CKL_API int ckl_device_count(void);

/* Compress; `labels` is a host (labels_on_device == 0) or device pointer.  The stream is left in a
 * context-owned device buffer; fetch it with ckl_result_copy / ckl_result_device.  A ckl_ctx is single-threaded (one
 * call at a time).  With a device pointer AND a caller-owned stream (ckl_ctx_set_stream) the call returns when the last
 * kernels are queued: *out_bytes is final, the bytes are complete in stream order, and the caller orders its own
 * reuse of `labels` on that stream. */
CKL_API int ckl_compress(ckl_ctx* ctx, const void* labels, int labels_on_device, int data_width,
                 uint64_t sx, uint64_t sy, uint64_t sz, int fortran_order, int markov_model_order,
                 uint64_t* out_bytes);
CKL_API int ckl_result_copy(ckl_ctx* ctx, void* dst, int dst_on_device, uint64_t capacity);
CKL_API const void* ckl_result_device(ckl_ctx* ctx, uint64_t* bytes);

/* Decompress; `binary` host or device, `out` host or device.  Host pointers (here and in ckl_compress / ckl_result_copy)
 * may be pageable or pinned: pageable ranges of 16 MiB and more are moved through a pinned staging ring by eight host
 * threads, pinned / registered memory is copied in place.  Distinct contexts may be used from distinct threads at once. */
CKL_API int ckl_decompress(ckl_ctx* ctx, const void* binary, int binary_on_device, uint64_t num_bytes,
                   int64_t z_start, int64_t z_end, int has_label, uint64_t label,
                   void* out, int out_on_device, uint64_t out_capacity);

/* Compressed-domain statistics -- voxel_counts, centroids and bounding_boxes of src/operations.hpp:321-665 (the
 * consumers of for_each_z_parallel, :89-182) -- for slices [z_start, z_end) without ever painting the volume: crack
 * decode, per-slice CCL and label map as in ckl_decompress, then every horizontal run adds its length, coordinate sums
 * and extent to its label's entry.  Entries are indexed like the stream's sorted unique label table:
 *   labels u64[n]  counts u64[n]  sums u64[n][3] = (sum x, sum y, sum z)  bbox u32[n][6] = (xmin,ymin,zmin,xmax,ymax,zmax)
 * (centroid = sums / count; a label absent from the z-range keeps count 0, mins 0xFFFFFFFF, maxs 0 like the
 * reference's freshly initialised boxes).  Each output may be NULL; capacity_entries >= the stream's label count
 * (crackle_b200_header + the u64 at the start of the labels section, or call once with all outputs NULL and
 * capacity 0 ... which fails with CKL_ERR_ARG but still reports *n_unique). */
CKL_API int ckl_label_stats(ckl_ctx* ctx, const void* binary, int binary_on_device, uint64_t num_bytes,
                    int64_t z_start, int64_t z_end, uint64_t* labels, uint64_t* counts, uint64_t* sums, uint32_t* bbox,
                    int out_on_device, uint64_t capacity_entries, uint64_t* n_unique);

/* Voxel connectivity graph of slices [z_start, z_end) -- crackle::operations::voxel_connectivity_graph
 * (src/operations.hpp:667-826; binding src/fastcrackle.cpp:538-565; bit layout src/crackcodes.hpp:706-862): one byte per
 * voxel in Fortran order (x fastest), a set bit = the neighbour in that direction is reachable, 00 -z +z -y +y -x +x.
 * connectivity 4: the crack planes read per voxel (no CCL, no labels).  connectivity 6 (streams with sz > 1): +z / -z are
 * set where the labels of vertically adjacent voxels agree, and the outer faces of the first and last decoded slice are
 * open.  Edges on the image border keep the reference's fill value: open for the IMPERMISSIBLE crack format, closed for
 * PERMISSIBLE.  Any other connectivity fails with CKL_ERR_ARG and the reference's text. */
CKL_API int ckl_voxel_connectivity_graph(ckl_ctx* ctx, const void* binary, int binary_on_device, uint64_t num_bytes,
                    int64_t z_start, int64_t z_end, int connectivity, uint8_t* out, int out_on_device, uint64_t out_capacity);

/* Re-code a stream's crack codes with another markov model order without decoding to voxels
 * (crackle::reencode_with_markov_order, src/crackle.hpp:860-984; fastcrackle.reencode_markov src/fastcrackle.cpp:643).
 * Labels, labels crc and slice crcs are carried over verbatim; the result is left in the context's result buffer
 * (ckl_result_copy / ckl_result_device). */
CKL_API int ckl_reencode(ckl_ctx* ctx, const void* binary, int binary_on_device, uint64_t num_bytes, int markov_model_order,
                 uint64_t* out_bytes);

/* Stream surgery on the device, no voxel decode (crackle/operations.py:424-548 zstack with :258-295
 * _zstack_flat_labels; :551-662 zsplit / zshatter through _zsplit_helper): flat-label, markov order 0, format
 * version 1 streams of equal sx, sy and crack format.  ckl_zstack: the stream of the inputs stacked along z (merged
 * sorted unique table, keys re-derived by binary search, crack codes / N_z / z-index entries / slice crcs copied).
 * ckl_zslice: the stream of slices [z_start, z_end) of one input with its own unique table (zsplit(b, z) is the three
 * ranges [0,z), [z,z+1), [z+1,sz); zshatter one range per slice).  Result: ckl_result_copy / ckl_result_device. */
CKL_API int ckl_zstack(ckl_ctx* ctx, int n, const void* const* binaries, const uint64_t* sizes, int on_device, uint64_t* out_bytes);
CKL_API int ckl_zslice(ckl_ctx* ctx, const void* binary, int on_device, uint64_t num_bytes, uint64_t z_start, uint64_t z_end,
               uint64_t* out_bytes);

/* ---- z-sharded multi-GPU compress (one context per GPU; the caller moves the small blobs between ranks,
 *      e.g. with torch.distributed all_gather over NCCL).  Mirrors what operations.zstack /
 *      _zstack_flat_labels (crackle/operations.py:258-295, 424-548) do for independently compressed slabs. --- */

/* Per-shard summary exchanged between ranks (fixed size, little-endian, 64 bytes). */
typedef struct ckl_shard_summary {
  uint64_t max_label;        /* lib::max_label over the shard                                        */
  uint64_t pairs;            /* lib::pixel_pairs inside the shard (excludes the pair with the previous shard) */
  uint64_t first_voxel;      /* value of the shard's first voxel (flat index 0)                      */
  uint64_t last_voxel;       /* value of the shard's last voxel                                      */
  uint64_t voxels;
  uint64_t reserved[3];
} ckl_shard_summary;

/* Stage 1: edge bit-planes + local reductions for a z-slab held on this GPU. */
CKL_API int ckl_shard_begin(ckl_ctx* ctx, const void* labels, int labels_on_device, int data_width,
                    uint64_t sx, uint64_t sy, uint64_t sz_local, ckl_shard_summary* summary);
/* Stage 2 (after the summaries are all-reduced): crack codes, CCL, CRCs, sorted unique labels of the shard.
 * permissible / stored_width are the GLOBAL decisions.  markov stats (if order > 0) are accumulated locally. */
CKL_API int ckl_shard_encode(ckl_ctx* ctx, int permissible, int stored_width, int markov_model_order,
                     uint64_t* n_unique_local, uint64_t* n_components_local, uint64_t* n_codepoints_local);
/* The same stage in two halves, for callers that overlap their exchange with the serial part of the tracer:
 * ckl_shard_encode_async queues the whole stage and returns when the CCL / label chain is through -- the two counts are
 * final and ckl_shard_unique may be called -- while the crack-code tracing may still be running on the context's side
 * streams; ckl_shard_encode_wait joins the tracer and returns the codepoint count and the order-0 code size.
 * ckl_shard_encode == ckl_shard_encode_async followed by ckl_shard_encode_wait.  (Reference precedent for encoding slabs
 * independently and merging their label tables afterwards: crackle/operations.py:258-295, 424-548.) */
CKL_API int ckl_shard_encode_async(ckl_ctx* ctx, int permissible, int stored_width, int markov_model_order,
                           uint64_t* n_unique_local, uint64_t* n_components_local);
CKL_API int ckl_shard_encode_wait(ckl_ctx* ctx, uint64_t* n_codepoints_local, uint64_t* codes_bytes_order0);
/* Copies the shard's sorted unique labels (uint64 each) to dst (host or device). */
CKL_API int ckl_shard_unique(ckl_ctx* ctx, uint64_t* dst, int dst_on_device);
/* Markov statistics of the shard: uint32[4^order * 4] (wrap mod 2^32 like the reference's atomics). */
CKL_API int ckl_shard_stats(ckl_ctx* ctx, uint32_t* dst, int dst_on_device);
/* Stage 3: given the GLOBAL sorted unique table (and, for order > 0, the GLOBAL stats), produce this shard's
 * pieces: keys (key_width bytes per component), N_z table entries, per-slice crack codes, per-slice crcs. */
typedef struct ckl_shard_pieces {
  uint64_t keys_bytes;         /* n_components_local * key_width                                     */
  uint64_t codes_bytes;        /* concatenated crack codes of the shard's slices                     */
  uint64_t sz_local;
} ckl_shard_pieces;
CKL_API int ckl_shard_finish(ckl_ctx* ctx, const uint64_t* global_unique, int unique_on_device, uint64_t n_unique_global,
                     const uint32_t* global_stats, int stats_on_device, ckl_shard_pieces* pieces);
/* Stored markov model bytes built from the global statistics (src/markov.hpp:325-380); 0 bytes for order 0. */
CKL_API int ckl_shard_model(ckl_ctx* ctx, uint8_t* dst, int dst_on_device, uint64_t capacity, uint64_t* model_bytes);
/* Copies the pieces out (each may be NULL to skip): keys, component counts (uint64 per slice), code sizes
 * (uint32 per slice), slice crcs (uint32 per slice), codes. */
CKL_API int ckl_shard_fetch(ckl_ctx* ctx, uint8_t* keys, uint64_t* components_per_slice, uint32_t* code_sizes,
                    uint32_t* slice_crcs, uint8_t* codes, int dst_on_device);

/* Local counts of the shard after ckl_shard_encode (codes_bytes / keys_bytes are valid after ckl_shard_finish;
 * codes_bytes_order0 is the order-0 code size, known right after the encode stage). */
typedef struct ckl_shard_counts {
  uint64_t n_unique_local, n_components, n_codepoints, codes_bytes_order0;
  uint64_t sz_local, runs, keys_bytes, codes_bytes;
} ckl_shard_counts;
CKL_API int ckl_shard_info(ckl_ctx* ctx, ckl_shard_counts* out);

/* One-collective exchange: ckl_shard_pack writes ALL of this shard's pieces into one 4-byte-aligned DEVICE buffer
 *     N_z u32[sz_local] | code sizes u32[sz_local] | slice crcs u32[sz_local] | keys[keys_bytes] | codes[codes_bytes]
 * (dst == NULL: only *bytes is returned) so the ranks can exchange them with a single padded all-gather
 * (torch.distributed.all_gather_into_tensor over NCCL).  ckl_shard_assemble then builds the complete .ckl stream
 * -- header, z index + crc, label table, N_z, keys, markov model, crack codes, labels crc, slice crcs
 * (src/crackle.hpp:171-216, src/labels.hpp:123-152) -- on the device from the gathered blocks, on every rank
 * alike, without a host round trip; the result is fetched with ckl_result_copy / ckl_result_device.  This is the
 * device-side equivalent of operations.zstack for slabs encoded against one global label table
 * (crackle/operations.py:424-548).  For markov_model_order > 0 the stored model is the one this context built in
 * ckl_shard_finish from the global statistics. */
typedef struct ckl_shard_block {
  uint64_t offset;           /* byte offset of the rank's block inside the gathered buffer (multiple of 4) */
  uint64_t sz_local, n_components, keys_bytes, codes_bytes;
} ckl_shard_block;
CKL_API int ckl_shard_pack(ckl_ctx* ctx, uint8_t* dst_device, uint64_t capacity, uint64_t* bytes);
CKL_API int ckl_shard_assemble(ckl_ctx* ctx, const uint8_t* gathered_device, const ckl_shard_block* blocks, int n_blocks,
                       const uint64_t* global_unique, int unique_on_device, uint64_t n_unique_global,
                       int data_width, int stored_width, int permissible, int fortran_order, int markov_model_order,
                       uint64_t sx, uint64_t sy, uint64_t* out_bytes);

/* ---- instrumentation and small device utilities ---------------------------------------------------------- */
/* Per-stage CUDA-event timing on the context's stream.  ckl_prof_read formats "stage=ms_total:calls;..." */
CKL_API int ckl_prof_enable(ckl_ctx* ctx, int on);
CKL_API int ckl_prof_read(ckl_ctx* ctx, char* buf, size_t cap);
/* Run the context's work on a caller-owned CUDA stream (cudaStream_t as void*; 0 = the legacy default stream, which
 * is what torch uses by default).  ckl_ctx_own_stream goes back to the context's private non-blocking stream. */
CKL_API int ckl_ctx_set_stream(ckl_ctx* ctx, void* stream);
CKL_API int ckl_ctx_own_stream(ckl_ctx* ctx);
/* z-chunk pipelining of ckl_compress / ckl_decompress: large volumes are processed as K z-ranges on child contexts
 * (own streams + workspaces) so the latency-bound stages of one range overlap the bandwidth-bound stages of the others;
 * the chunks are merged like the reference merges z-slabs (crackle/operations.py:424-548 zstack) -- output bytes are
 * identical to the unchunked path.  chunks: 0 = automatic (default), 1 = off, K = always K chunks. */
CKL_API int ckl_ctx_set_chunks(ckl_ctx* ctx, int chunks);
/* Number of kernels this library has launched in this process. */
CKL_API uint64_t ckl_launch_count(void);
/* host-side waits on a compute stream (cudaStreamSynchronize) issued by the library since load: the drains of the hot calls */
CKL_API uint64_t ckl_sync_count(void);
/* CRC-32C (Castagnoli) of a host or device buffer, computed on the GPU (src/crc.hpp:39-57). */
CKL_API int ckl_crc32c(ckl_ctx* ctx, const void* data, int on_device, uint64_t n, uint32_t* out);
/* In-place sort + unique of a DEVICE array of uint64 (only the low key_bytes*8 bits are compared). */
CKL_API int ckl_sort_unique_u64(ckl_ctx* ctx, uint64_t* data_device, uint64_t n, int key_bytes, uint64_t* n_unique);

/* Library / build information. */
CKL_API const char* crackle_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CRACKLE_B200_H */
