// ckl_internal.cuh -- internal interfaces between the translation units of libcrackle_b200.so
#pragma once
#include "ckl_common.cuh"

// one place for the SM count and the grid size of grid-stride kernels (ckl_api.cu): min(blocks needed, SMs x blocks per SM
// [x g_ckl_grid_mult while z-chunks run concurrently])
int ckl_num_sms();
u32 ckl_grid(u64 items, u32 per_block, u32 blocks_per_sm, bool fine = true);

struct Geom {
  u32 sx, sy, sz;     // sz = slices held by this context (a z-slab for sharded jobs)
  u32 W;              // 32-pixel words per row = ceil(sx/32)
  u64 sxy;            // sx*sy  (< 2^31)
  __host__ __device__ u64 rows() const { return (u64)sy * sz; }
  __host__ __device__ u64 words() const { return (u64)sy * sz * W; }
};

// ---- device scalars written by kernels and read back in one small copy --------------------------------
enum {
  SC_MAX = 0, SC_PAIRS, SC_RUNS, SC_COMPONENTS, SC_EDGES, SC_SYMCAP, SC_STACKCAP, SC_CHAINCAP, SC_CPCAP,
  SC_CODEPOINTS, SC_CODE_BYTES, SC_ERROR, SC_UNIQUE, SC_FIRST, SC_LAST, SC_CRC_BAD, SC_NODES, SC_MAXNODES, SC_COUNT = 24
};

// ---- planes + CCL (ckl_planes.cu) ------------------------------------------------------------------------
struct CclBufs {
  DBuf wordPrefix, rowRuns, rowBase, sliceRuns, runBase;   // per word / row / slice
  DBuf parent, runStart, compRank;                         // per run (allocated once the run total is known)
  DBuf nz, compBase, compPix;                              // per slice / per component
  DBuf sliceCrc;                                           // raw CRC accumulators, per slice
  DBuf crcH, crcHl; u64 crcHn = 0;                         // H[m] = sum_{j=1..m} x^(32j) mod P for m <= sxy (built once)
};

void launch_edges(const void* labels, int width, const Geom& g, u32* DV, u32* DH, ull* scal, cudaStream_t st);
// phase 1: per-row run prefixes, per-slice run counts and bases; writes scal[SC_RUNS]
void launch_ccl_count(const Geom& g, const u32* DV, CclBufs& B, ull* scal, cudaStream_t st);
// phase 2 (parent/runStart/compRank/compPix must hold `total_runs`): union-find, ranks, N_z, crcs;
// writes scal[SC_COMPONENTS]
void launch_ccl_solve(const Geom& g, const u32* DV, const u32* DH, CclBufs& B, const CrcTables* d_tables,
                      ull* scal, u64 total_runs, cudaStream_t st);
// second half: per-run component ranks, per-slice CRCs and (compress) component first pixels / (decompress, `decode`
// non-null) the label of every run
struct CclDecodeSrc { const u8* uniq; const u8* keys; u64 n_uniq, n_keys; int sw, kw; const u64* keyBase; u64* runLabel;
                      const u64* uniq64 = nullptr; const u64* keys64 = nullptr;      // aligned copies (launch_unpack_le), optional
                      bool keys_only = false; };                                      // runLabel receives the key (unique-table index), not the label
void launch_unpack_le(const u8* src, int width, u64 n, u64* dst, cudaStream_t st);
void launch_ccl_finish(const Geom& g, CclBufs& B, u64 total_runs, const CrcTables* d_tables, u32 crc_init_term,
                       const CclDecodeSrc* decode, cudaStream_t st);
// compressed-domain statistics: per-run reduction into tables indexed like the sorted unique label table
void launch_run_stats(const Geom& g, u32 z_first, const CclBufs& B, u64 total_runs, const u64* runKey, u64 nu, ull* counts, ull* sums,
                      u32* bbox, cudaStream_t st);
// generic device CRC-32C of a byte buffer: result (finalised) written to *d_out
void launch_crc_bytes(const u8* d, u64 n, const CrcTables* d_tables, const CrcTables& h_tables, u32* d_out, cudaStream_t st);
// finalise per-slice raw registers into standard CRCs (in place)
void launch_crc_finalize_slices(u32* sliceCrc, u32 sz, u32 init_term, cudaStream_t st);

// ---- tracing + packing (ckl_trace.cu) ------------------------------------------------------------------
struct TraceBufs {
  DBuf VW;                                      // uint4 = 32 adjacency nibbles per 32-vertex word of the (sx+1) x (sy+1) vertex grid
  DBuf NM, nodeP;                               // node mask per 32-vertex word; padded linear vertex index per node
  DBuf nodePrefix, rowNodes, rowBase;           // node numbering: per vertex word / per vertex row
  DBuf sliceNodes, nodeBase;                    // per slice: node count (u32), first global node index (u64 x (sz+1))
  DBuf nodeVertex, nodeAdj;                     // per node: vertex index (u32), remaining-edge nibble (u8, global replay)
  DBuf seFar, seLen;                            // per (node, direction) slot: far node << 2 | arrival dir, length
  DBuf bounds;                                  // per slice: E, S, C + caps (2 x 4 x u32 x sz)
  DBuf offs;                                    // per slice u64 offsets: events, stack, chain, cp  (4 x (sz+1))
  DBuf ev, evRec, evCp, stack, chain, cp;       // sized from scal[SC_*CAP]
  DBuf sliceInfo;                               // per slice: ncp, nchains, boc_bytes, code_bytes (4 x u32 x sz)
  DBuf codeOff;                                 // u64 x (sz+1) byte offsets of each slice's crack code
};
// symBegin / symEnd / t2f are EVENT indices inside the slice's event list
struct ChainRec { u32 adjStart, symBegin, symEnd, t2f; u32 ncp, outBase, sortedIdx, sortedStart; };

void launch_trace_prepare(const Geom& g, const u32* DV, const u32* DH, int permissible, TraceBufs& T, ull* scal, cudaStream_t st);
// order 0: per-slice code sizes -> codeOff (exclusive scan, total in scal[SC_CODE_BYTES]) then pack into dst
void launch_code_sizes_order0(const Geom& g, TraceBufs& T, ull* scal, cudaStream_t st);
void launch_pack_order0(const Geom& g, TraceBufs& T, u8* dst, cudaStream_t st);

// ---- markov (ckl_markov.cu) ------------------------------------------------------------------------------
struct MarkovBufs {
  DBuf stats;      // u32[4^order][4]
  DBuf model;      // u8[4^order][4]  rank of symbol
  DBuf stored;     // stored model bytes
  DBuf bitlen;     // per slice u64 bit counts
  DBuf scratch;    // per-slice word-aligned bitstreams
  DBuf scratchOff; // u64 x (sz+1) word offsets
};
void launch_markov_stats(const Geom& g, TraceBufs& T, int order, u32* stats, cudaStream_t st);
void launch_markov_model(int order, const u32* stats, u8* model, u8* stored, u64 stored_bytes, cudaStream_t st);
void launch_markov_sizes(const Geom& g, TraceBufs& T, MarkovBufs& M, ull* scal, cudaStream_t st);
void launch_markov_encode(const Geom& g, TraceBufs& T, int order, const u8* model, MarkovBufs& M, cudaStream_t st);

// ---- label table (ckl_labels.cu) -------------------------------------------------------------------------
struct LabelBufs {
  DBuf mapping;    // u64 per component (z order)
  DBuf sorted;     // u64 per component
  DBuf uniq;       // u64 per unique label
  DBuf tmp;        // cub temp storage
  DBuf flags;
};
void launch_gather_mapping(const void* labels, int width, const Geom& g, const CclBufs& B, u64 ncomp, u64* mapping, cudaStream_t st);
// sorts + uniques `mapping` (n items) into L.uniq; returns count on host (synchronises the stream)
u64 labels_sort_unique(LabelBufs& L, u64 n, int key_bits, cudaStream_t st, ull* count_dev = nullptr);
// keys[i] = index of mapping[i] in uniq (n_uniq entries), written little-endian with key_width bytes
void launch_write_keys(const u64* mapping, u64 n, const u64* uniq, u64 n_uniq, int key_width, u8* dst, cudaStream_t st);
// uniq table written little-endian with `stored_width` bytes per entry
void launch_write_uniq(const u64* uniq, u64 n_uniq, int stored_width, u8* dst, cudaStream_t st);
// little-endian packing of u32 / u64 arrays with an arbitrary byte width
void launch_write_le_u32(const u32* src, u64 n, int width, u8* dst, cudaStream_t st);

// ---- decode (ckl_decode.cu) ------------------------------------------------------------------------------
// per decoded slice: absolute stream offsets of its crack code and of the code body (after the BOC index)
struct DecSlice { u64 code, body, wordOff; u32 blen, isz; };
struct DecodeBufs {
  DBuf slices, wordOff;   // DecSlice x szr, word capacity offsets u64 x (szr+1)
  DBuf fields;      // order > 0: packed 2-bit difference fields (u32 per 16)
  DBuf Mw, Sw, Qw, evBaseW;   // per 16-codepoint word: absolute moves, event mask, displacement prefix, events before
  DBuf nev, ncp, evOff, nevUsed;   // per slice
  DBuf evIdx, segQ, segStart, gstack, segSum;   // per event
  DBuf redo;        // per slice: the chain pass must be redone by the serial kernel
  DBuf codeOff;     // u64 x (szr+1): absolute offsets of each decoded slice's crack code inside the stream
  DBuf keyBase;     // u64 x szr: first key index of each decoded slice
  DBuf storedNz;    // u32 x szr
  DBuf model;       // u8[4^order][4]: symbol of rank
  DBuf stack;       // per-slice revisit stacks
  DBuf stackOff;
  DBuf runLabel;    // u64 per run
  DBuf uniq64, keys64;   // aligned copies of the stream's unique-label and key tables
};
void launch_decode_slices(const Geom& g, const u8* stream, const u64* codeOff, int permissible, int order,
                          const u8* model, u32* EV, u32* EH, u32* stack, const u64* stackOff, ull* scal, cudaStream_t st);
void launch_decode_slices_init(const Geom& g, const u8* stream, const u64* codeOff, const u64* wordOff, DecSlice* out, ull* scal, cudaStream_t st);
void launch_decode_classify(const Geom& g, const u8* stream, int order, const u8* model, DecodeBufs& D, u64 total_words, ull* scal,
                            cudaStream_t st);
void launch_decode_mark(const Geom& g, const u8* stream, int order, DecodeBufs& D, u64 total_events, u64 total_words, u32* EV, u32* EH,
                        ull* scal, cudaStream_t st);
void launch_planes_from_cracks(const Geom& g, int permissible, u32* EV, u32* EH, cudaStream_t st);
void launch_paint(const Geom& g, const u32* DV, const CclBufs& B, const u64* runLabel, int out_width, int has_label,
                  u64 label, int fortran_order, void* out, cudaStream_t st);
