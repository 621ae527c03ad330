// ckl_trace.cu -- crack-code chain tracing on the edge bit-planes, exact reproduction of the reference's
// start-point and branch order, followed by the parallel post-passes (initial-branch reversal, spurious branch
// removal, escape coding, chain ordering) and the order-0 packer.
//
// Reference behaviour reproduced:
//   create_crack_codes walk / next_cluster / erase_edge     src/crackcodes.hpp:390-450, :41-64
//   remove_initial_branch                                   src/crackcodes.hpp:185-242
//   remove_spurious_branches                                src/crackcodes.hpp:250-281  (done online in the replay)
//   symbols_to_codepoints                                   src/crackcodes.hpp:128-183
//   write_boc_index / pack_codepoints                       src/crackcodes.hpp:318-372, :455-496
//
// The reference walk is a serial, order-dependent trail over the crack graph (one chain of ~10^5 moves per slice).
// Here it is replayed on the CONTRACTED graph (validated against the oracle by tests/bringup/proto_trace.py):
//   1. k_vw_build     vertex words {r,d,u,n}: right / down / up edge bits of 32 vertices and the node mask n.
//                     node = static degree 1, 3 or 4, or an (R,D)-only corner whose horizontal run to the right
//                     does not end at a vertex with an up edge.  Every component's minimum vertex (the chain
//                     start of next_cluster) has only R / D edges and is such a node; a degree-2 vertex entered
//                     through one edge never branches, so all other vertices are pure pass-through.
//   2. node numbering in raster order (per-row popcount prefix) => next_cluster = smallest node id with edges left
//   3. k_path_walk    one thread per (node, direction): follow degree-2 vertices to the far node -> super-edges
//   4. k_replay       one warp per slice: the reference walk over node records held in shared memory (4 x u16 per
//                     node), emitting an event list (E super-edge / B / T / S) -- ~6x fewer serial steps, no
//                     global-memory latency on the dependent chain
//   5. k_event_post   scans: codepoint offsets per event, chain order by adjusted start vertex, BOC sizes
//   6. k_expand       one thread per event: re-walk the super-edge writing absolute codepoints at their final
//                     position (reversed + flipped inside a removed initial branch); b/t escape pairs by lookback
//
// Crack graph on bit-planes (pixel-indexed, W words per row): right edge of vertex (vx,vy) = EH bit vx of row vy
// (horizontal crack above pixel (vx,vy), vy >= 1); down edge of vertex (vx,vy) = EV bit vx of row vy (vertical
// crack left of pixel (vx,vy), vx >= 1).
#include "ckl_internal.cuh"

#define NONE32 0xFFFFFFFFu
enum { EV_E = 0, EV_B = 1, EV_T = 2, EV_S = 3 };      // event type in bits 30..31, payload (local slot) below

static u32 grid_cap(u64 items, u32 per_block, u32 blocks_per_sm) { return ckl_grid(items, per_block, blocks_per_sm); }
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

struct VGeom {           // vertex grid of one slice
  u32 sxe, sye, Wv;      // (sx+1), (sy+1), 32-vertex words per vertex row
  u32 S;                 // padded vertices per row (32 * Wv): p = y * S + x is the linear vertex index the walkers use
  u64 rowsAll, wordsAll; // over all slices
};
static VGeom vgeom(const Geom& g) {
  VGeom v;
  v.sxe = g.sx + 1; v.sye = g.sy + 1; v.Wv = (v.sxe + 31) / 32; v.S = v.Wv * 32;
  v.rowsAll = (u64)v.sye * g.sz; v.wordsAll = v.rowsAll * v.Wv;
  return v;
}

// one vertex's adjacency nibble (bit 0 right, 1 left, 2 down, 3 up) out of a byte-planar vertex word: the four plane bits
// of vertex b sit at bits b, 8 + b, 16 + b, 24 + b; the multiply gathers them into the top byte (no carries: sums <= 15)
__device__ __forceinline__ u32 vertex_nibble(u32 word, u32 b) { return ((((word >> b) & 0x01010101u) * 0x01020408u) >> 24); }
// 1. vertex words + node mask + per-slice bounds: bounds[z] = {E edges, S node slots with an edge, C start-capable nodes}
// grid = (vertex words of one slice / 256, slices): 32-bit index arithmetic, and a block's three bound sums meet in shared
// memory before they go to the slice's counters.
template <int PERM>      // crack format as a compile-time constant: the IMPERMISSIBLE form is the plain planes
__global__ void __launch_bounds__(256) k_vw_build(Geom g, VGeom vg, const u32* __restrict__ DV, const u32* __restrict__ DH,
                                                   uint4* __restrict__ VW, u32* __restrict__ NM, u32* __restrict__ cnt, u32* bounds) {
  constexpr int perm = PERM;
  __shared__ u32 sb[3];
  const u32 perSlice = vg.Wv * vg.sye;                      // < 2^32: checked by the caller
  const u32 il = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = il < perSlice;
  const u32 y = in ? il / vg.Wv : 0u;
  const u32 w = il - y * vg.Wv;
  for (u32 z = blockIdx.y; z < g.sz; z += gridDim.y) {
    if (threadIdx.x < 3) sb[threadIdx.x] = 0;
    __syncthreads();
    u32 e = 0, s = 0, c = 0;
    if (in) {
      const u64 i = (u64)z * perSlice + il;
      const u64 prow = ((u64)z * g.sy + y) * g.W;           // pixel-plane row (valid when y < sy)
      u32 r = 0, rp = 0, d = 0, u = 0;
      if (y < g.sy) {
        if (w < g.W) { r = crack_h(g, DH, prow + w, y, w, perm); d = crack_v(g, DV, prow + w, w, perm); }
        if (w > 0 && w - 1 < g.W) rp = crack_h(g, DH, prow + w - 1, y, w - 1, perm);
      }
      if (y >= 1 && w < g.W) u = crack_v(g, DV, prow - g.W + w, w, perm);
      const u32 l = (r << 1) | (rp >> 31);
      // bit-sliced degree of the 32 vertices of this word
      const u32 s1 = r ^ l, c1 = r & l, s2 = d ^ u, c2 = d & u;
      const u32 sum0 = s1 ^ s2, carry = s1 & s2;
      const u32 two = c1 ^ c2 ^ carry, four = c1 & c2;
      (void)two;
      u32 n = sum0 | four;                                   // degree 1, 3, 4
      // (R,D)-only corners stay nodes unless their horizontal run to the right ends (inside this word) at a vertex
      // with an up edge (then the component reaches a higher row and the corner cannot be its minimum vertex)
      u32 cm = r & d & ~l & ~u;
      const u32 stop = ~r | d | u;
      while (cm) {
        const u32 b = __ffs(cm) - 1;
        cm &= cm - 1;
        const u32 sm = b == 31 ? 0u : (stop & ~((2u << b) - 1u));
        if (sm == 0 || !((u >> (__ffs(sm) - 1)) & 1u)) n |= 1u << b;
      }
      // adjacency of the 32 vertices, 8 vertices per u32 as four byte planes: byte 0 right, 1 left, 2 down, 3 up (three
      // byte permutes per word; a walker gathers one vertex's nibble with a shift, a mask and a multiply).  Degree-2 nodes
      // (the corner candidates) are stored as 0 so that "popcount != 2" is the walkers' only node test.
      const u32 pass = ~(n & ~(sum0 | four));                // clear the corner nodes (degree exactly 2)
      const u32 rr = r & pass, dd = d & pass;
      uint4 nib;
      nib.x = __byte_perm(__byte_perm(rr, l, 0x0040), __byte_perm(dd, u, 0x0040), 0x5410);
      nib.y = __byte_perm(__byte_perm(rr, l, 0x0051), __byte_perm(dd, u, 0x0051), 0x5410);
      nib.z = __byte_perm(__byte_perm(rr, l, 0x0062), __byte_perm(dd, u, 0x0062), 0x5410);
      nib.w = __byte_perm(__byte_perm(rr, l, 0x0073), __byte_perm(dd, u, 0x0073), 0x5410);
      VW[i] = nib;
      NM[i] = n;
      cnt[i] = __popc(n);
      e = __popc(r) + __popc(d);
      s = __popc(n & r) + __popc(n & l) + __popc(n & d) + __popc(n & u);
      c = __popc(n & ~l & ~u);
    }
    e = __reduce_add_sync(FULL_MASK, e);
    s = __reduce_add_sync(FULL_MASK, s);
    c = __reduce_add_sync(FULL_MASK, c);
    if ((threadIdx.x & 31) == 0) {
      if (e) atomicAdd(&sb[0], e);
      if (s) atomicAdd(&sb[1], s);
      if (c) atomicAdd(&sb[2], c);
    }
    __syncthreads();
    if (threadIdx.x < 3 && sb[threadIdx.x]) atomicAdd(bounds + (u64)z * 4 + threadIdx.x, sb[threadIdx.x]);
    __syncthreads();
  }
}

// 2a. per vertex row: counts -> exclusive prefix (in place) and row totals.  One warp per row.
__global__ void __launch_bounds__(256) k_node_prefix(VGeom vg, u32* __restrict__ cnt, u32* __restrict__ rowNodes) {
  const u32 lane = threadIdx.x & 31;
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  for (u64 row = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < vg.rowsAll; row += nwarps) {
    u32 carry = 0;
    for (u32 w0 = 0; w0 < vg.Wv; w0 += 32) {
      const u32 w = w0 + lane;
      const u32 c = w < vg.Wv ? cnt[row * vg.Wv + w] : 0;
      const u32 inc = warp_incl_scan(c);
      if (w < vg.Wv) cnt[row * vg.Wv + w] = carry + inc - c;
      carry += __shfl_sync(FULL_MASK, inc, 31);
    }
    if (lane == 0) rowNodes[row] = carry;
  }
}
// 2b. per slice: row totals -> row bases, slice totals
__global__ void __launch_bounds__(256) k_node_rows(u32 sz, u32 sye, const u32* __restrict__ rowNodes, u32* __restrict__ rowBase,
                                                    u32* __restrict__ sliceNodes) {
  __shared__ u32 sm[33];
  for (u32 z = blockIdx.x; z < sz; z += gridDim.x) {
    u32 carry = 0;
    for (u32 y0 = 0; y0 < sye; y0 += blockDim.x) {
      const u32 y = y0 + threadIdx.x;
      const u32 v = y < sye ? rowNodes[(u64)z * sye + y] : 0;
      u32 tot;
      const u32 ex = block_excl_scan(v, sm, tot);
      if (y < sye) rowBase[(u64)z * sye + y] = carry + ex;
      carry += tot;
    }
    if (threadIdx.x == 0) sliceNodes[z] = carry;
  }
}

// caps[z] = {evCap, stackCap, chainCap, cpCap}
__global__ void k_trace_caps(u32 sz, const u32* __restrict__ bounds, const u32* __restrict__ sliceNodes, u32* __restrict__ caps, ull* scal) {
  const u32 z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= sz) return;
  const u32 E = bounds[(u64)z * 4], S = bounds[(u64)z * 4 + 1], C = bounds[(u64)z * 4 + 2], N = sliceNodes[z];
  const u32 B = S - N;                               // >= number of 'b' events: sum over nodes of (degree - 1)
  caps[(u64)z * 4 + 0] = S / 2 + 2 * B + C + 4;      // events: super-edges + b + t (t = b + chains)
  caps[(u64)z * 4 + 1] = B + 4;                      // revisit stack depth
  caps[(u64)z * 4 + 2] = C + 2;                      // chains
  caps[(u64)z * 4 + 3] = (E + 4 * B + 2 * C + 16 + 3u) & ~3u;   // codepoints: moves + 2b + 2t; slice bases stay 4-byte aligned
  atomicMax(&scal[SC_MAXNODES], (ull)N);
}

void launch_trace_prepare(const Geom& g, const u32* DV, const u32* DH, int permissible, TraceBufs& T, ull* scal, cudaStream_t st) {
  const VGeom vg = vgeom(g);
  T.VW.ensure(vg.wordsAll * 16);
  T.NM.ensure(vg.wordsAll * 4);
  T.nodePrefix.ensure(vg.wordsAll * 4);
  T.rowNodes.ensure(vg.rowsAll * 4);
  T.rowBase.ensure(vg.rowsAll * 4);
  T.sliceNodes.ensure((u64)g.sz * 4);
  T.nodeBase.ensure(((u64)g.sz + 1) * 8);
  T.bounds.ensure((u64)g.sz * 4 * 4 * 2);          // bounds (4 x u32) + caps (4 x u32) per slice
  T.offs.ensure(((u64)g.sz + 1) * 8 * 4);
  T.sliceInfo.ensure((u64)g.sz * 4 * 4);
  T.codeOff.ensure(((u64)g.sz + 1) * 8);
  u32* bounds = T.bounds.as<u32>();
  u32* caps = bounds + (u64)g.sz * 4;
  CUDA_CHECK(cudaMemsetAsync(bounds, 0, (u64)g.sz * 4 * 4, st));
  if ((u64)vg.Wv * vg.sye > 0xFFFFFFFFull) throw CklError(CKL_ERR_ARG, "crackle_b200: slice too large");
  const dim3 vgrid((vg.Wv * vg.sye + 255) / 256, g.sz < 65535u ? g.sz : 65535u);
  if (permissible) k_vw_build<1><<<vgrid, 256, 0, st>>>(g, vg, DV, DH, T.VW.as<uint4>(), T.NM.as<u32>(), T.nodePrefix.as<u32>(), bounds);
  else k_vw_build<0><<<vgrid, 256, 0, st>>>(g, vg, DV, DH, T.VW.as<uint4>(), T.NM.as<u32>(), T.nodePrefix.as<u32>(), bounds);
  LAUNCH_CHECK();
  k_node_prefix<<<grid_cap(vg.rowsAll, 8, 8), 256, 0, st>>>(vg, T.nodePrefix.as<u32>(), T.rowNodes.as<u32>());
  LAUNCH_CHECK();
  k_node_rows<<<grid_cap(g.sz, 1, 8), 256, 0, st>>>(g.sz, vg.sye, T.rowNodes.as<u32>(), T.rowBase.as<u32>(), T.sliceNodes.as<u32>());
  LAUNCH_CHECK();
  launch_exscan_u32_u64(T.sliceNodes.as<u32>(), g.sz, 1, T.nodeBase.as<u64>(), &scal[SC_NODES], 0, st);
  k_trace_caps<<<(g.sz + 255) / 256, 256, 0, st>>>(g.sz, bounds, T.sliceNodes.as<u32>(), caps, scal);
  LAUNCH_CHECK();
  u64* offs = T.offs.as<u64>();
  const u64 n1 = (u64)g.sz + 1;
  launch_exscan_u32_u64(caps + 0, g.sz, 4, offs + 0 * n1, &scal[SC_SYMCAP], 0, st);
  launch_exscan_u32_u64(caps + 1, g.sz, 4, offs + 1 * n1, &scal[SC_STACKCAP], 0, st);
  launch_exscan_u32_u64(caps + 2, g.sz, 4, offs + 2 * n1, &scal[SC_CHAINCAP], 0, st);
  launch_exscan_u32_u64(caps + 3, g.sz, 4, offs + 3 * n1, &scal[SC_CPCAP], 0, st);
}

// ---------------------------------------------------------------------------------------------------------
struct TraceParams {
  Geom g;
  VGeom vg;
  const uint4* VW;         // adjacency nibbles, 32 vertices per uint4
  const u32* NM;           // node mask per 32-vertex word
  u32* nodeP;              // per node: padded linear vertex index y * S + x
  const u32* nodePrefix;   // exclusive node count inside the vertex row, per vertex word
  const u32* rowBase;      // per vertex row: first (slice-local) node id
  const u32* sliceNodes;
  const u64* nodeBase;     // per slice: first global node index
  u32* nodeVertex;         // per node: vx + sxe * vy
  u32* seFar;              // per slot (4 per node): (far local node << 2) | arrival direction at the far node; NONE32 = no edge
  u32* seLen;              // per slot: moves along the super-edge
  u8* nodeAdj;             // per node: remaining-edge nibble (global-memory replay only)
  const u64* offs;         // 4 arrays of (sz+1): events, stack, chain, cp
  const u32* caps;         // per slice 4 x u32
  u32* ev;                 // events
  u32 pwChunk, exChunk;    // slots / events per warp chunk of the path walkers
  uint4* evRec;            // per event: walk task {output offset, padded vertex index, -, length | direction << 29 | flip << 31}
  u32* evCp;               // per event: exclusive codepoint offset inside the slice (creation order); one extra slot per slice
  uint2* stack;
  ChainRec* chain;
  u8* cp;
  u32* sliceInfo;          // per slice: nev|ncp, nchains, bocBytes, codeBytes
  ull* scal;
};

__device__ __forceinline__ u32 slice_of(const u64* __restrict__ base, u32 sz, u64 idx) {   // last z with base[z] <= idx
  u32 lo = 0, hi = sz;
  while (hi - lo > 1) { const u32 m = (lo + hi) >> 1; if (base[m] <= idx) lo = m; else hi = m; }
  return lo;
}

// 3a. node -> vertex, padded vertex index and static adjacency nibble
__global__ void __launch_bounds__(256) k_node_init(TraceParams P) {      // grid = (vertex words of one slice / 256, slices)
  const VGeom vg = P.vg;
  const u32 perSlice = vg.Wv * vg.sye;
  const u32 il = blockIdx.x * blockDim.x + threadIdx.x;
  if (il >= perSlice) return;
  const u32 y = il / vg.Wv, w = il - y * vg.Wv;
  for (u32 z = blockIdx.y; z < P.g.sz; z += gridDim.y) {
    const u64 i = (u64)z * perSlice + il;
    u32 n = P.NM[i];
    if (!n) continue;
    u64 id = P.nodeBase[z] + P.rowBase[(u64)z * vg.sye + y] + P.nodePrefix[i];
    const uint4 word = P.VW[i];
    while (n) {
      const u32 b = __ffs(n) - 1;
      n &= n - 1;
      const u32 q = b >> 3;
      const u32 c = q == 0 ? word.x : (q == 1 ? word.y : (q == 2 ? word.z : word.w));
      const u32 nib = vertex_nibble(c, b & 7u);
      P.nodeVertex[id] = y * vg.sxe + w * 32 + b;
      P.nodeP[id] = y * vg.S + w * 32 + b;
      P.nodeAdj[id] = (u8)(nib ? nib : 5u);          // a corner node is stored as 0; its edges are right and down
      id++;
    }
  }
}

// one step along a super-edge: move in direction kk (0 right, 1 left, 2 down, 3 up: the walk priority), then pick the
// exit of the vertex reached.  Returns true when that vertex is a node (its nibble does not have exactly two edges).
__device__ __forceinline__ u32 bfind_u32(u32 x) { u32 r; asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x)); return r; }   // position of the highest set bit
__device__ __forceinline__ bool nib_step(const u32* __restrict__ vn, u32 S, u32& p, u32& kk) {
  p += ((kk & 2u) ? S : 1u) * (1u - ((kk & 1u) << 1));
  const u32 nib = vertex_nibble(__ldg(vn + (p >> 3)), p & 7u);
  if ((0xE997u >> nib) & 1u) return true;            // popcount(nib) != 2
  const u32 a = nib ^ (1u << (kk ^ 1u));             // not back along the edge we came by: one bit left
  kk = bfind_u32(a);                                 // one-hot {1,2,4,8} -> {0,1,2,3}
  return false;
}
// the same step where the vertex reached is known to be a pass-through vertex (inside a super-edge of known length)
__device__ __forceinline__ void nib_step_inner(const u32* __restrict__ vn, u32 S, u32& p, u32& kk) {
  p += ((kk & 2u) ? S : 1u) * (1u - ((kk & 1u) << 1));
  const u32 nib = vertex_nibble(__ldg(vn + (p >> 3)), p & 7u);
  kk = bfind_u32(nib ^ (1u << (kk ^ 1u))) & 3u;
}
// predicated global store / OR-reduction of one word (no branch around them in the walkers' step)
__device__ __forceinline__ void st_u32_if(u32* a, u32 v, bool p) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.u32 [%0], %1; }" ::"l"(a), "r"(v), "r"((u32)p) : "memory");
}
__device__ __forceinline__ void red_or_u32_if(u32* a, u32 v, bool p) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q red.global.or.b32 [%0], %1; }" ::"l"(a), "r"(v), "r"((u32)p) : "memory");
}
#define WALK_QUAD 4             // steps per control round of the walkers (ballot / refill / exit test)
#define WALK_BLOCK 128          // threads per block of the walkers: a block waits for its slowest warp, so small blocks

// 3b. super-edges: a (node, direction) slot with an edge follows degree-2 vertices to the far node and records
// the result at BOTH ends.  Pass 0 walks the right / down slots; pass 1 walks the left / up slots that pass 0 did
// not already fill from the other end (most super-edges leave one node rightwards or downwards and arrive at the
// other from the left or from above, so almost every path is walked once instead of twice).
// grid = (chunks, slices); a lane takes the next slot of its warp's chunk as soon as its current path reaches a node
// (the refill reads the node's static adjacency nibble, so slots without an edge cost one byte load).
#define SE_UNSET 0xFEFEFEFEu
#define WALK_REFILL 8           // idle lanes that trigger a refill of the walkers (k_path_walk, k_expand)
// Small chunks = many warps per slice = few slices in flight at once: the vertex words the walkers chase (541 KB per
// 1024^2 slice) then stay L2-resident instead of streaming from DRAM once per step.
#define PW_CHUNK 128u
#define PW_PASS1_MULT 4u
template <int PASS>
__global__ void __launch_bounds__(WALK_BLOCK) k_path_walk(TraceParams P) {
  const VGeom vg = P.vg;
  const u32 lane = threadIdx.x & 31;
  const u32 ltmask = (1u << lane) - 1u;
  const u32 limit = 2u * vg.sxe * vg.sye + 8u;
  for (u32 z = blockIdx.y; z < P.g.sz; z += gridDim.y) {
    const u32 N = P.sliceNodes[z];
    if (!N) continue;
    const u64 nb = P.nodeBase[z];
    const u64 rowz = (u64)z * vg.sye;
    const u32* vn = reinterpret_cast<const u32*>(P.VW + rowz * vg.Wv);
    const u8* adj = P.nodeAdj + nb;
    const u32* nodeP = P.nodeP + nb;
    u32* seFar = P.seFar + nb * 4;
    u32* seLen = P.seLen + nb * 4;
    const u32 nitems = 2u * N;
    const u32 PWC = P.pwChunk;
    const u32 nchunks = (nitems + PWC - 1) / PWC;
    for (u32 chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); chunk < nchunks; chunk += gridDim.x * (blockDim.x >> 5)) {
      u32 next = chunk * PWC;
      const u32 end = min(nitems, next + PWC);
      bool active = false;
      u32 pos = 0, kk = 0, len = 0, slot = 0;
      for (;;) {
        const u32 idle = __ballot_sync(FULL_MASK, !active);
        if ((__popc(idle) >= WALK_REFILL && next < end) || idle == FULL_MASK) {      // refill in batches: the refill code is as long as the step itself
          const u32 i = next + __popc(idle & ltmask);
          if (!active && i < end) {
            const u32 node = i >> 1, k0 = (i & 1u) * 2u + PASS;      // pass 0: right, down; pass 1: left, up
            const u32 s = node * 4 + k0;
            if (!((adj[node] >> k0) & 1u)) { seFar[s] = NONE32; seLen[s] = 0; }
            else if (PASS == 0 || seFar[s] == SE_UNSET) {            // pass 1: not filled from the other end
              pos = nodeP[node];
              kk = k0; len = 0; slot = s;
              active = true;
            }
          }
          next = min(end, next + __popc(idle));
          if (!__any_sync(FULL_MASK, active)) {
            if (next >= end) break;
            continue;
          }
        }
        if (active) {
          // WALK_QUAD steps per control round (ballot, refill and exit tests cost as much as a step); a lane that reaches
          // its far node sits out the rest of the round
          bool done = false;
#pragma unroll
          for (int q = 0; q < WALK_QUAD; q++)
            if (!done) { len++; done = nib_step(vn, vg.S, pos, kk); }
          if (done) {
            const u32 y = pos / vg.S, x = pos - y * vg.S;
            const u64 row = rowz + y;
            const u64 wi = row * vg.Wv + (x >> 5);
            const u32 far = P.rowBase[row] + P.nodePrefix[wi] + __popc(P.NM[wi] & ((1u << (x & 31)) - 1u));
            const u32 fk = kk ^ 1u;
            seFar[slot] = (far << 2) | fk;
            seLen[slot] = len;
            const u32 twin = far * 4 + fk;
            seFar[twin] = slot;                                      // (this node << 2) | departure direction
            seLen[twin] = len;
            active = false;
          } else if (len > limit) {                                  // cannot happen on a consistent vertex plane: every cycle holds a node
            atomicExch(&P.scal[SC_ERROR], 4ull);
            seFar[slot] = NONE32; seLen[slot] = 0;
            active = false;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// 4. the replay.  Node records: SMEM mode = one u64 per node holding four u16 entries (far << 2 | arrival dir),
// 0xFFFF = no edge left; GLOBAL mode = a remaining-edge nibble per node in global memory + the read-only seFar.
// Revisit stack: the walk is a depth-first search whose stack runs thousands of entries deep on a 1024^2 slice, and
// almost every pop happens far from the bottom.  Every push goes to global memory (fire and forget) AND to a shared-memory
// ring holding the top REPLAY_STACK entries; a pop below the ring's valid range refills REPLAY_REFILL entries with
// independent loads.  Measured on the bench volume: ~5300 pops per slice, ~80 refills.
#define REPLAY_STACK 256      // ring entries (power of two)
#define REPLAY_REFILL 32

// A node record travels in registers between steps: after consuming edge k the far node's record is loaded,
// cleared of the arrival edge and becomes the current record, so the dependent chain of one step is a single
// shared-memory load plus a few ALU ops.  Three stores:
//   MODE 0  shared memory, slices with <= 8191 nodes: uint2 per node = four 15-bit entries (far << 2 | arrival dir)
//           at bits 0/15 of each word and the remaining-edge nibble in bits 30..31 of the two words
//   MODE 1  shared memory, <= 16382 nodes: four u16 entries, 0xFFFF = edge consumed
//   MODE 2  global memory: remaining-edge nibble per node + the read-only super-edge table
#define REPLAY_CAP0 8191u
#define REPLAY_CAP1 16382u
template <int MODE> struct NodeStore;

// Shared memory through explicit 32-bit shared-space addresses: with generic pointers the compiler re-derives the shared
// window base (S2R SR_CgaCtaId + LEA) inside the serial loop, ~25 cycles on the dependent chain of every step.
__device__ __forceinline__ u32 smem_addr(const void* p) {
  u32 a = (u32)__cvta_generic_to_shared(p);
  asm volatile("" : "+r"(a));          // opaque: keep the address in a register instead of rematerialising it
  return a;
}
__device__ __forceinline__ uint2 lds64(u32 a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ u32 lds32(u32 a) {
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(u32 a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

template <> struct NodeStore<0> {
  typedef uint2 Rec;      // 64-bit record (x = low word): entry k at bits [15k, 15k+15), remaining-edge nibble at bits 60..63
  uint2* rec; u8* adj; const u32* far;
  u32 ra;                 // shared-space address of rec[0]
  __device__ __forceinline__ void bind() { ra = smem_addr(rec); }
  __device__ __forceinline__ void init(u32 i, uint4 f) {
    u64 v = 0;
    if (f.x != NONE32) v |= (u64)(f.x & 0x7FFFu) | (1ull << 60);
    if (f.y != NONE32) v |= ((u64)(f.y & 0x7FFFu) << 15) | (2ull << 60);
    if (f.z != NONE32) v |= ((u64)(f.z & 0x7FFFu) << 30) | (4ull << 60);
    if (f.w != NONE32) v |= ((u64)(f.w & 0x7FFFu) << 45) | (8ull << 60);
    rec[i] = make_uint2((u32)v, (u32)(v >> 32));
  }
  __device__ __forceinline__ Rec load(u32 node) const { return lds64(ra + node * 8u); }
  __device__ __forceinline__ u32 adjacency(Rec w) const { return w.y >> 28; }
  __device__ __forceinline__ void take(u32& node, u32 k, Rec& w) {
    const u32 e = (u32)((((u64)w.y << 32) | w.x) >> (15u * k)) & 0x7FFFu;
    sts32(ra + node * 8u + 4u, w.y & ~(0x10000000u << k));              // only the nibble word changes
    const u32 f = e >> 2, fk = e & 3u;
    w = lds64(ra + f * 8u);                                              // after the store: a self-loop sees its own update
    w.y &= ~(0x10000000u << fk);
    sts32(ra + f * 8u + 4u, w.y);
    node = f;
  }
  __device__ __forceinline__ bool has_edges(u32 node) const { return (lds32(ra + node * 8u + 4u) >> 28) != 0; }
};

template <> struct NodeStore<1> {
  typedef u64 Rec;
  u64* rec; u8* adj; const u32* far;
  __device__ __forceinline__ void bind() {}
  __device__ __forceinline__ void init(u32 i, uint4 f) {
    const u64 e0 = f.x == NONE32 ? 0xFFFFull : (u64)(f.x & 0xFFFFu), e1 = f.y == NONE32 ? 0xFFFFull : (u64)(f.y & 0xFFFFu);
    const u64 e2 = f.z == NONE32 ? 0xFFFFull : (u64)(f.z & 0xFFFFu), e3 = f.w == NONE32 ? 0xFFFFull : (u64)(f.w & 0xFFFFu);
    rec[i] = e0 | (e1 << 16) | (e2 << 32) | (e3 << 48);
  }
  __device__ __forceinline__ Rec load(u32 node) const { return rec[node]; }
  __device__ __forceinline__ u32 adjacency(Rec w) const {
    const u32 lo = (u32)w, hi = (u32)(w >> 32);
    return ((lo & 0xFFFFu) != 0xFFFFu ? 1u : 0u) | ((lo >> 16) != 0xFFFFu ? 2u : 0u) | ((hi & 0xFFFFu) != 0xFFFFu ? 4u : 0u) |
           ((hi >> 16) != 0xFFFFu ? 8u : 0u);
  }
  __device__ __forceinline__ void take(u32& node, u32 k, Rec& w) {
    const u32 e = (u32)(w >> (16 * k)) & 0xFFFFu;
    w |= 0xFFFFull << (16 * k);
    rec[node] = w;
    const u32 f = e >> 2, fk = e & 3u;
    if (f != node) w = rec[f];
    w |= 0xFFFFull << (16 * fk);
    rec[f] = w;
    node = f;
  }
  __device__ __forceinline__ bool has_edges(u32 node) const { return rec[node] != ~0ull; }
};

template <> struct NodeStore<2> {
  typedef u32 Rec;
  u64* rec; u8* adj; const u32* far;
  __device__ __forceinline__ void bind() {}
  __device__ __forceinline__ void init(u32 i, uint4 f) {
    adj[i] = (u8)((f.x != NONE32 ? 1u : 0u) | (f.y != NONE32 ? 2u : 0u) | (f.z != NONE32 ? 4u : 0u) | (f.w != NONE32 ? 8u : 0u));
  }
  __device__ __forceinline__ Rec load(u32 node) const { return adj[node]; }
  __device__ __forceinline__ u32 adjacency(Rec w) const { return w; }
  __device__ __forceinline__ void take(u32& node, u32 k, Rec& w) {
    const u32 e = far[(u64)node * 4 + k];
    adj[node] = (u8)(w & ~(1u << k));
    const u32 f = e >> 2, fk = e & 3u;
    w = adj[f] & ~(1u << fk) & 0xFu;
    adj[f] = (u8)w;
    node = f;
  }
  __device__ __forceinline__ bool has_edges(u32 node) const { return adj[node] != 0; }
};

__device__ __forceinline__ int replay_mode(u32 N) { return N <= REPLAY_CAP0 ? 0 : (N <= REPLAY_CAP1 ? 1 : 2); }

__device__ __forceinline__ void st_ev(u32* base, u32 idx, u32 v) {     // event store: st.global through an opaque 64-bit base
  asm volatile("st.global.u32 [%0], %1;" ::"l"(base + idx), "r"(v));
}

template <int MODE>
__global__ void __launch_bounds__(32) k_replay(TraceParams P) {
  extern __shared__ u64 smem64[];
  const u32 z = blockIdx.x, lane = threadIdx.x;
  const Geom g = P.g;
  const u32 N = P.sliceNodes[z];
  if (replay_mode(N) != MODE) return;                 // another instantiation handles this slice
  const u64 n1 = (u64)g.sz + 1;
  const u64 nb = P.nodeBase[z];
  uint2* sstack = reinterpret_cast<uint2*>(smem64);   // REPLAY_STACK entries
  NodeStore<MODE> S;
  S.rec = reinterpret_cast<decltype(S.rec)>(smem64 + REPLAY_STACK);
  S.adj = P.nodeAdj + nb;
  S.far = P.seFar + nb * 4;
  S.bind();
  const u32 ssa = smem_addr(sstack);
  for (u32 i = lane; i < N; i += 32) S.init(i, reinterpret_cast<const uint4*>(P.seFar + nb * 4)[i]);
  __syncwarp();
  u32* ev = P.ev + P.offs[0 * n1 + z];
  uint2* gstack = P.stack + P.offs[1 * n1 + z];
  asm volatile("" : "+l"(ev));                        // opaque base: one IMAD.WIDE per event address in the serial loop
  ChainRec* chains = P.chain + P.offs[2 * n1 + z];
  const u32 evCap = P.caps[(u64)z * 4 + 0], stackCap = P.caps[(u64)z * 4 + 1], chainCap = P.caps[(u64)z * 4 + 2];
  const u32* nodeVertex = P.nodeVertex + nb;
  u32 nev = 0, nch = 0, cursor = 0;
  bool ok = true;
  while (ok) {
    // next_cluster: smallest node id >= cursor with an edge left (all earlier nodes are exhausted)
    u32 start = NONE32;
    for (u32 base = cursor; base < N; base += 32) {
      const u32 i = base + lane;
      const u32 m = __ballot_sync(FULL_MASK, i < N && S.has_edges(i));
      if (m) { start = base + (__ffs(m) - 1); break; }
    }
    if (start == NONE32) break;
    if (lane == 0) {
      if (nch >= chainCap) { atomicExch(&P.scal[SC_ERROR], 3ull); ok = false; }
      else {
        u32 node = start, sp = 0, low = 0, ne = nev;          // ne: running event index of the slice; low: lowest stack index valid in the ring
        const u32 begin = nev;
        bool firstT = true;
        u32 t2f = 0, adjStart = nodeVertex[start];
        typename NodeStore<MODE>::Rec w = S.load(node);
        const u32 a0 = S.adjacency(w);
        const bool firstIsB = (a0 & (a0 - 1)) != 0;           // the chain opens with a 'b'
        // The serial loop only records what happened: moves, 'b's, and 't's carrying the event index of the 'b' they
        // return to.  Spurious branches (remove_spurious_branches) are marked afterwards, in parallel, by k_event_post.
        // Event / stack capacities are exact upper bounds (k_trace_caps), so the loop carries no capacity checks.
        for (;;) {
          const u32 a = S.adjacency(w);
          if (a == 0) {                                        // dead end: a 't'
            if (__builtin_expect(firstT, 0)) {
              firstT = false;      // no pop yet, so the stack depth is the number of 'b's so far
              if (sp == 1 && firstIsB) { t2f = ne - begin; adjStart = nodeVertex[node]; }   // remove_initial_branch applies
            }
            if (sp == 0) { st_ev(ev, ne++, ((u32)EV_T << 30) | 0x3FFFFFFFu); break; }
            --sp;
            if (__builtin_expect(sp < low, 0)) {               // below the ring: refill from the global copy
              const u32 lo2 = sp + 1 >= REPLAY_REFILL ? sp + 1 - REPLAY_REFILL : 0u;
              for (u32 base = lo2; base <= sp; base += 16) {
                uint2 t[16];
#pragma unroll
                for (u32 j = 0; j < 16; j++) t[j] = gstack[min(base + j, sp)];
#pragma unroll
                for (u32 j = 0; j < 16; j++) sts64(ssa + (min(base + j, sp) & (REPLAY_STACK - 1)) * 8u, t[j]);
              }
              low = lo2;
            }
            const uint2 e = lds64(ssa + (sp & (REPLAY_STACK - 1)) * 8u);
            st_ev(ev, ne++, ((u32)EV_T << 30) | e.y);          // payload: the 'b' this 't' returns to
            node = e.x;
            w = S.load(node);
            continue;
          }
          // branch point (popcount > 1): push, branch-free -- the slot above the top of the stack is scratch, and the
          // 'b' event written here is overwritten by the move below when nothing was pushed
          const u32 push = (a & (a - 1)) ? 1u : 0u;
          const uint2 e = make_uint2(node, ne);
          sts64(ssa + (sp & (REPLAY_STACK - 1)) * 8u, e);
          gstack[sp] = e;
          st_ev(ev, ne, (u32)EV_B << 30);
          low = max(low, max(sp + 1u, (u32)REPLAY_STACK) - REPLAY_STACK);   // the write above (pushed or scratch) replaced entry sp - REPLAY_STACK
          sp += push; ne += push;
          const u32 k = (0x12131210u >> (2u * a)) & 3u;        // ctz of the nibble by table (priority: right, left, down, up)
          st_ev(ev, ne++, ((u32)EV_E << 30) | (node * 4 + k));
          S.take(node, k, w);
        }
        if (ne > evCap || sp > stackCap) { atomicExch(&P.scal[SC_ERROR], 1ull); ok = false; }
        const bool t2 = t2f != 0;
        nev = ne;
        ChainRec& rec = chains[nch];
        rec.adjStart = adjStart;
        rec.symBegin = begin;
        rec.symEnd = nev;
        rec.t2f = t2 ? t2f : 0;
        nch++;
      }
    }
    nev = __shfl_sync(FULL_MASK, nev, 0);
    nch = __shfl_sync(FULL_MASK, nch, 0);
    ok = __shfl_sync(FULL_MASK, (u32)ok, 0) != 0;
    __syncwarp();                                       // lane 0's record stores before the other lanes' has_edges() loads
    cursor = start;
  }
  if (lane == 0) { P.sliceInfo[(u64)z * 4 + 0] = nev; P.sliceInfo[(u64)z * 4 + 1] = nch; }
}

// ---------------------------------------------------------------------------------------------------------
// 5. post-passes over the event list
__device__ __forceinline__ u32 chain_of(const ChainRec* chains, u32 nch, u32 i) {
  u32 lo = 0, hi = nch;
  while (hi - lo > 1) { const u32 m = (lo + hi) >> 1; if (chains[m].symBegin <= i) lo = m; else hi = m; }
  return lo;
}

__global__ void __launch_bounds__(256) k_event_post(TraceParams P) {
  __shared__ u32 sm[33];
  const Geom g = P.g;
  const u64 n1 = (u64)g.sz + 1;
  const u32 sxe = g.sx + 1;
  const int xw = ckl_byte_width(g.sx + 1), yw = ckl_byte_width(g.sy + 1);
  for (u32 z = blockIdx.x; z < g.sz; z += gridDim.x) {
    u32* ev = P.ev + P.offs[0 * n1 + z];
    u32* pre = P.evCp + P.offs[0 * n1 + z] + z;            // one extra slot per slice
    ChainRec* chains = P.chain + P.offs[2 * n1 + z];
    const u32* seLen = P.seLen + P.nodeBase[z] * 4;
    const u32 nev = P.sliceInfo[(u64)z * 4 + 0], nch = P.sliceInfo[(u64)z * 4 + 1];
    // (0) remove_spurious_branches (crackcodes.hpp:250-281): a 't' directly after a 't' closes a branch that emitted no
    // symbol -- that 't' and the 'b' the previous 't' returned to (its payload) disappear, unless that 'b' is the chain's
    // first symbol and the initial branch is being removed.  Types change to 's' with the payload kept, so a neighbour
    // reading a converted 't' still finds its payload; a 'b' is always followed by a move, never by a 't'.
    for (u32 i = threadIdx.x; i < nev; i += blockDim.x) {
      const u32 e = ev[i];
      if ((e >> 30) != EV_T || i == 0) continue;
      const u32 q = ev[i - 1], qt = q >> 30;
      if (qt != EV_T && qt != EV_S) continue;
      const ChainRec& c = chains[chain_of(chains, nch, i)];
      if (i <= c.symBegin) continue;
      const u32 pb = q & 0x3FFFFFFFu;
      if (c.t2f && pb == c.symBegin) continue;
      ev[pb] = (u32)EV_S << 30;
      ev[i] = ((u32)EV_S << 30) | (e & 0x3FFFFFFFu);
    }
    __syncthreads();
    // (a) remove_initial_branch: the leading 'b' and the first 't' of the chain disappear
    for (u32 c = threadIdx.x; c < nch; c += blockDim.x)
      if (chains[c].t2f) { ev[chains[c].symBegin] = (u32)EV_S << 30; ev[chains[c].symBegin + chains[c].t2f] = (u32)EV_S << 30; }
    __syncthreads();
    // (b) codepoints per event -> exclusive prefix
    u32 carry = 0;
    constexpr u32 PER = 4;                               // consecutive events per thread: four independent loads per block scan
    for (u32 i0 = 0; i0 < nev; i0 += blockDim.x * PER) {
      const u32 i = i0 + threadIdx.x * PER;
      u32 cnt[PER], sum = 0;
#pragma unroll
      for (u32 j = 0; j < PER; j++) {
        cnt[j] = 0;
        if (i + j < nev) {
          const u32 e = ev[i + j], t = e >> 30;
          cnt[j] = t == EV_E ? seLen[e & 0x3FFFFFFFu] : (t == EV_S ? 0u : 2u);
        }
        sum += cnt[j];
      }
      u32 tot;
      u32 run = carry + block_excl_scan(sum, sm, tot);
#pragma unroll
      for (u32 j = 0; j < PER; j++) {
        if (i + j < nev) pre[i + j] = run;
        run += cnt[j];
      }
      carry += tot;
    }
    if (threadIdx.x == 0) pre[nev] = carry;
    __syncthreads();
    const u32 ncp = carry;
    // (c) per-chain codepoint counts, rank by adjusted start vertex (chains of a slice are vertex-disjoint)
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
    if (threadIdx.x == 0) {
      const u32 boc = 4 + yw + nrows * (yw + xw) + nch * xw;
      P.sliceInfo[(u64)z * 4 + 0] = ncp;
      P.sliceInfo[(u64)z * 4 + 2] = boc;
      P.sliceInfo[(u64)z * 4 + 3] = boc + (ncp + 3) / 4;
    }
    __syncthreads();
  }
}

// 6. expansion: absolute codepoints (symbols_to_codepoints), written at their final position.
// direction index -> codepoint: right 1, left 3, down 2, up 0  (crackcodes.hpp:20-26)
__device__ __forceinline__ u8 dir_code(u32 kk) { return (u8)((0x0231u >> (4 * kk)) & 0xFu); }

// 6a. per event (fully parallel): the walk task of a super-edge event -- output position, start vertex, direction,
// length, reversed/flipped inside a removed initial branch -- as one 16-byte record; 'b' / 't' escape pairs are
// written here directly (they depend only on the previous kept symbol; kept 't's in between alternate).
#define EX_CHUNK 256u
#define EX_LEN_MASK 0x1FFFFFFFu        // record.w = length | direction << 29 | flip << 31
__global__ void __launch_bounds__(256) k_event_setup(TraceParams P) {
  const Geom g = P.g;
  const VGeom vg = P.vg;
  const u64 n1 = (u64)g.sz + 1;
  for (u32 z = blockIdx.y; z < g.sz; z += gridDim.y) {
    const u32 nch = P.sliceInfo[(u64)z * 4 + 1];
    if (!nch) continue;
    const ChainRec* chains = P.chain + P.offs[2 * n1 + z];
    const u32 nev = chains[nch - 1].symEnd;            // sliceInfo[0] holds ncp by now; events end with the last chain
    const u32* ev = P.ev + P.offs[z];
    const u32* pre = P.evCp + P.offs[z] + z;
    uint4* recs = P.evRec + P.offs[z];
    u8* cpz = P.cp + P.offs[3 * n1 + z];
    const u64 nb = P.nodeBase[z];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nev; i += gridDim.x * blockDim.x) {
      const u32 e = ev[i], t = e >> 30;
      uint4 rec = make_uint4(0, 0, 0, 0);
      if (t != EV_S) {
        const ChainRec& c = chains[chain_of(chains, nch, i)];
        const u32 rel = pre[i] - pre[c.symBegin];
        if (t == EV_E) {
          const u32 slot = e & 0x3FFFFFFFu;
          const u32 len = P.seLen[nb * 4 + slot];
          rec.y = P.nodeP[nb + (slot >> 2)];
          u32 flip = 0;
          if (c.t2f && i < c.symBegin + c.t2f) {
            // inside a removed initial branch: the move run is reversed and every move flipped
            const u32 fcp = pre[c.symBegin + c.t2f] - pre[c.symBegin];
            rec.x = c.outBase + (fcp - 1 - rel); flip = 1;
          } else rec.x = c.outBase + rel;
          if (len > EX_LEN_MASK) atomicExch(&P.scal[SC_ERROR], 4ull);
          rec.w = (len & EX_LEN_MASK) | ((slot & 3u) << 29) | (flip << 31);
        } else {
          u32 tcount = 0, j = i;
          int prev = -1;                                        // previous kept move (direction index), -1 = none
          while (j > c.symBegin) {
            j--;
            const u32 q = ev[j], qt = q >> 30;
            if (qt == EV_S) continue;
            if (qt == EV_T) { tcount++; continue; }
            if (qt == EV_B) break;                              // cannot happen (a 'b' is always followed by a move)
            if (c.t2f && j < c.symBegin + c.t2f) prev = (int)((ev[c.symBegin + 1] & 3u) ^ 1u);   // reversed run ends with flip(first move)
            else prev = (int)((P.seFar[nb * 4 + (q & 0x3FFFFFFFu)] & 3u) ^ 1u);                  // last move = opposite of arrival dir
            break;
          }
          u8* ob = cpz + c.outBase + rel;
          if (t == EV_B) {
            // (UP,DOWN) unless first symbol of the chain or the previous codepoint is DOWN -> (LEFT,RIGHT)
            const bool alt = (i == c.symBegin) || (tcount == 0 && prev == 2);
            ob[0] = alt ? 3 : 0;
            ob[1] = alt ? 1 : 2;
          } else {
            // (DOWN,UP) unless the previous codepoint is UP -> (RIGHT,LEFT); consecutive 't's alternate
            const bool alt = ((prev == 3) ? 1u : 0u) ^ (tcount & 1u);
            ob[0] = alt ? 1 : 2;
            ob[1] = alt ? 3 : 0;
          }
        }
      }
      recs[i] = rec;
    }
  }
}

// 6b. the walk: grid = (chunks, slices); a lane takes the next task record of its warp's chunk as soon as its current
// super-edge is written out, so the refill is one 16-byte load and the loop body is the step itself.  Codepoints are
// gathered into the 32-bit word of `cp` they belong to and leave with ONE store per word: a plain store for a word the
// super-edge covers entirely, an OR-reduction for the partial words at its two ends (neighbouring events own the other bytes;
// `cp` is zeroed before k_event_setup writes the escape pairs).  The body is branch-free apart from the reversed initial
// branch (rare, written back to front byte by byte).  The slice bases of `cp` are multiples of 4 (k_trace_caps).
__global__ void __launch_bounds__(WALK_BLOCK) k_expand(TraceParams P) {
  const Geom g = P.g;
  const VGeom vg = P.vg;
  const u64 n1 = (u64)g.sz + 1;
  const u32 lane = threadIdx.x & 31;
  const u32 ltmask = (1u << lane) - 1u;
  for (u32 z = blockIdx.y; z < g.sz; z += gridDim.y) {
    const u32 nch = P.sliceInfo[(u64)z * 4 + 1];
    if (!nch) continue;
    const ChainRec* chains = P.chain + P.offs[2 * n1 + z];
    const u32 nev = chains[nch - 1].symEnd;
    const uint4* recs = P.evRec + P.offs[z];
    u8* cpz = P.cp + P.offs[3 * n1 + z];
    u32* cpw = reinterpret_cast<u32*>(cpz);
    const u32* vn = reinterpret_cast<const u32*>(P.VW + (u64)z * vg.sye * vg.Wv);
    const u32 EXC = P.exChunk;
    const u32 nchunks = (nev + EXC - 1) / EXC;
    for (u32 chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); chunk < nchunks; chunk += gridDim.x * (blockDim.x >> 5)) {
      u32 next = chunk * EXC;
      const u32 end = min(nev, next + EXC);
      u32 pos = 0, kk = 0, left = 0, flip = 0, off = 0, acc = 0, full = 0;
      for (;;) {
        const u32 idle = __ballot_sync(FULL_MASK, left == 0);
        if ((__popc(idle) >= WALK_REFILL && next < end) || idle == FULL_MASK) {
          const u32 i = next + __popc(idle & ltmask);
          if (left == 0 && i < end) {
            const uint4 r = __ldg(recs + i);
            left = r.w & EX_LEN_MASK;
            kk = (r.w >> 29) & 3u; flip = r.w >> 31;
            pos = r.y;
            off = r.x; acc = 0;
            full = (off & 3u) == 0 ? 1u : 0u;                      // the word being gathered started at its first byte
          }
          next = min(end, next + __popc(idle));
          if (!__any_sync(FULL_MASK, left != 0)) {
            if (next >= end) break;
            continue;
          }
        }
#pragma unroll
        for (int q = 0; q < WALK_QUAD; q++) {
          if (left) {
            const u32 code = dir_code(kk ^ flip);
            --left;
            if (flip) { cpz[off] = (u8)code; off--; }
            else {
              acc |= code << (8u * (off & 3u));
              u32* wa = cpw + (off >> 2);
              off++;
              const bool wordEnd = (off & 3u) == 0;
              st_u32_if(wa, acc, wordEnd && full);                         // all four bytes are this super-edge's
              red_or_u32_if(wa, acc, (wordEnd && !full) || (!wordEnd && left == 0));   // partial word at either end
              if (wordEnd) { acc = 0; full = 1u; }
            }
            if (left) nib_step_inner(vn, vg.S, pos, kk);
          }
        }
      }
    }
  }
}

static TraceParams make_params(const Geom& g, TraceBufs& T, ull* scal);

// node numbering -> vertices; returns false when the shard has no crack-graph nodes at all
bool launch_trace_nodes(const Geom& g, TraceBufs& T, ull* scal, u64 total_nodes, cudaStream_t st) {
  TraceParams P = make_params(g, T, scal);
  const VGeom vg = P.vg;
  k_node_init<<<dim3((vg.Wv * vg.sye + 255) / 256, g.sz < 65535u ? g.sz : 65535u), 256, 0, st>>>(P);
  LAUNCH_CHECK();
  if (!total_nodes) {
    CUDA_CHECK(cudaMemsetAsync(T.sliceInfo.p, 0, (u64)g.sz * 16, st));
    return false;
  }
  CUDA_CHECK(cudaMemsetAsync(T.seFar.p, 0xFE, total_nodes * 16, st));     // SE_UNSET
  return true;
}
// super-edges between nodes
void launch_trace_paths(const Geom& g, TraceBufs& T, ull* scal, u32 max_nodes, cudaStream_t st) {
  TraceParams P = make_params(g, T, scal);
  const u32 gy = g.sz < 65535u ? g.sz : 65535u;
  const u32 bs = WALK_BLOCK, wpb = bs / 32;
  u32 gx = (2u * max_nodes + wpb * P.pwChunk - 1) / (wpb * P.pwChunk);
  if (gx > 256) gx = 256;
  if (gx < 1) gx = 1;
  k_path_walk<0><<<dim3(gx, gy), bs, 0, st>>>(P);
  LAUNCH_CHECK();
  // pass 1 finds most of its slots already filled from the other end: four times the slots per warp keep its lanes busy
  P.pwChunk *= PW_PASS1_MULT;
  gx = (2u * max_nodes + wpb * P.pwChunk - 1) / (wpb * P.pwChunk);
  if (gx < 1) gx = 1;
  k_path_walk<1><<<dim3(gx, gy), bs, 0, st>>>(P);
  LAUNCH_CHECK();
}
// the serial chain replay: slices are replayed by the instantiation matching their node count (the others return at once)
void launch_trace_replay(const Geom& g, TraceBufs& T, ull* scal, u32 max_nodes, cudaStream_t st) {
  TraceParams P = make_params(g, T, scal);
  const size_t stack_bytes = (size_t)REPLAY_STACK * 8;
  {
    const u32 n0 = max_nodes < REPLAY_CAP0 ? max_nodes : REPLAY_CAP0;
    const size_t smem = stack_bytes + (size_t)n0 * 8;
    if (smem > 48 * 1024)   // per device, cheap: set on every launch that needs it
      CUDA_CHECK(cudaFuncSetAttribute(k_replay<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(stack_bytes + (size_t)REPLAY_CAP0 * 8)));
    k_replay<0><<<g.sz, 32, smem, st>>>(P);
    LAUNCH_CHECK();
  }
  if (max_nodes > REPLAY_CAP0) {
    const u32 n1 = max_nodes < REPLAY_CAP1 ? max_nodes : REPLAY_CAP1;
    const size_t smem = stack_bytes + (size_t)n1 * 8;
    CUDA_CHECK(cudaFuncSetAttribute(k_replay<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(stack_bytes + (size_t)REPLAY_CAP1 * 8)));
    k_replay<1><<<g.sz, 32, smem, st>>>(P);
    LAUNCH_CHECK();
  }
  if (max_nodes > REPLAY_CAP1) {
    k_replay<2><<<g.sz, 32, stack_bytes, st>>>(P);
    LAUNCH_CHECK();
  }
}
void launch_trace_post(const Geom& g, TraceBufs& T, ull* scal, u64 total_ev_cap, cudaStream_t st) {
  TraceParams P = make_params(g, T, scal);
  k_event_post<<<grid_cap(g.sz, 1, 8), 256, 0, st>>>(P);
  LAUNCH_CHECK();
  if (total_ev_cap) {
    const u64 per_slice = (total_ev_cap + g.sz - 1) / g.sz;
    const u32 bs = WALK_BLOCK, wpb = bs / 32;
    u32 gx = (u32)((per_slice + wpb * P.exChunk - 1) / (wpb * P.exChunk));
    if (gx > 256) gx = 256;
    if (gx < 1) gx = 1;
    const u32 gy = g.sz < 65535u ? g.sz : 65535u;
    u32 gs = (u32)((per_slice + 255) / 256);
    if (gs > 64) gs = 64;
    if (gs < 1) gs = 1;
    k_event_setup<<<dim3(gs, gy), 256, 0, st>>>(P);
    LAUNCH_CHECK();
    k_expand<<<dim3(gx, gy), bs, 0, st>>>(P);
    LAUNCH_CHECK();
  }
  // total codepoints (for the "all slices empty" rule, crackle.hpp:107-118)
  launch_exscan_u32_u64(T.sliceInfo.as<u32>(), g.sz, 4, T.codeOff.as<u64>(), &scal[SC_CODEPOINTS], 0, st);
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
  P.vg = vgeom(g);
  P.VW = T.VW.as<uint4>(); P.NM = T.NM.as<u32>(); P.nodeP = T.nodeP.as<u32>(); P.nodePrefix = T.nodePrefix.as<u32>(); P.rowBase = T.rowBase.as<u32>();
  P.sliceNodes = T.sliceNodes.as<u32>(); P.nodeBase = T.nodeBase.as<u64>(); P.nodeVertex = T.nodeVertex.as<u32>();
  P.seFar = T.seFar.as<u32>(); P.seLen = T.seLen.as<u32>(); P.nodeAdj = T.nodeAdj.as<u8>();
  P.offs = T.offs.as<u64>();
  P.caps = T.bounds.as<u32>() + (u64)g.sz * 4;
  P.ev = T.ev.as<u32>(); P.evRec = T.evRec.as<uint4>(); P.evCp = T.evCp.as<u32>(); P.stack = T.stack.as<uint2>(); P.chain = T.chain.as<ChainRec>();
  P.cp = T.cp.as<u8>(); P.sliceInfo = T.sliceInfo.as<u32>();
  P.scal = scal;
  P.pwChunk = PW_CHUNK; P.exChunk = EX_CHUNK;
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
