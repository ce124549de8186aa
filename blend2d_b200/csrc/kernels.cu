// kernels.cu - the sm_100a kernels of libb2dgpu and their launchers.
//
//   K1  k_build_edges<count|write>   one thread per path segment: transform -> clip -> monotone split -> flatten
//                                    (FP64, no FMA) -> 24.8 integer edges; curves are flattened by the whole warp (one
//                                    node of the subdivision tree per lane).  Two passes (count, exclusive scan, write)
//                                    give every command a contiguous edge range.
//   K1b k_scan_*                     exclusive prefix sum of the per-segment edge counts.
//   K1c k_analytic_bbox / k_finalize_commands   per-command pixel bounding boxes used for tile culling.
//   K1d k_bin_*                      per-band ordered command lists + per (band, command) column extents (tile culling)
//                                    and the list of the command's edges that cross the band (k_bin_edges).
//   K2+K3 k_tile_render<BPP,TH>      one CTA of TH warps per 128 x TH destination tile, a warp per block of 32 columns x
//                                    4 rows.  The tile's pixels are loaded ONCE into registers (16-byte vector loads),
//                                    every command that touches the tile is replayed in submission order - phase 1: a
//                                    warp per command classifies its edges against the tile and rasterizes the (edge,
//                                    row) crossings into per-row entry lists in shared memory; phase 2: every warp builds
//                                    the masks of its block from the backdrop and the entries, fetches, composites - and
//                                    the tile is stored ONCE.  This is the GPU form of the reference's per-band command
//                                    replay (blend2d/raster/workerproc.cpp:166-299) fused with its FillBoxA / FillMask /
//                                    FillAnalytic pipelines (pipeline/reference/fillgeneric_p.h:22-388).
//   K3s k_box_stream / k_stream_solid  persistent streaming compositors for batches of a few large box fills.
//
// No tensor cores: nothing on this path is a dense contraction.  Compile with -fmad=false (see dev_flatten.cuh).
#include "kernels.h"
#include <stdlib.h>
#include "dev_pixel.cuh"
#include "dev_raster.cuh"
#include "dev_flatten.cuh"
#include "dev_fetch.cuh"
#include "dev_tile.cuh"
#include "dev_glyph.cuh"

#include <cuda_runtime.h>
#include <limits.h>

namespace b2d {

// =================================================================================================================
// K0 - glyph instancing: TrueType outline + instance matrix -> vertices + path segments
// =================================================================================================================
struct GlyphVertexPut {
  double2* dst;
  __device__ __forceinline__ void operator()(uint32_t index, double x, double y) { dst[index] = make_double2(x, y); }
};

__global__ void __launch_bounds__(128) k_glyph_instances(GlyphParams P) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.instance_count) return;
  const b2dgpu_glyph_instance inst = P.instances[i];
  const GlyphBlobView g = glyph_blob_view(P.cache + inst.blob_offset);
  GlyphVertexPut put{ reinterpret_cast<double2*>(P.vertices) + inst.vertex_base };
  const uint32_t n = glyph_emit(g, inst.m, put);
  if (n != g.vertices) atomicOr(P.error_flag, 2u);                 // the host checked every entry: cannot happen
  const uint32_t* sw = g.segment_words();
  b2dgpu_segment* seg = P.segments + inst.segment_base;
  for (uint32_t k = 0; k < g.segments; k++) {
    b2dgpu_segment o;
    o.p0 = inst.vertex_base + sw[2 * k];
    o.p1_kind = ((inst.vertex_base + (sw[2 * k + 1] >> 2)) << 2) | (sw[2 * k + 1] & 3u);
    o.command = inst.command;
    seg[k] = o;
  }
}

// =================================================================================================================
// K0b - gradient table interpolation
//
// interpolate_prgb32 as the reference executes it on AVX2 hosts (pixelops/interpolation_avx2.cpp:18-205; called by
// GradientInternal::ensure_lut32, core/gradient.cpp:289-330): the stops are walked in order, a span of i + 1 table
// entries between two stops is a linear ramp of 8-bit channels in 9.23 fixed point with the per-channel step
// trunc((c1 - c0) * (2^23 / i)) computed in double, each entry is premultiplied afterwards, and a later span overwrites
// the entry it shares with the previous one.  The ramp is a closed form in the entry index (u32 wrap-around adds), so
// the lanes of a warp take the entries of a span in parallel; the spans stay sequential.
// =================================================================================================================
__device__ __forceinline__ uint32_t lut_premul_top8(uint64_t c) {
  const uint32_t a = uint32_t(c >> 56), r = uint32_t((c >> 40) & 0xFFu), g = uint32_t((c >> 24) & 0xFFu), b = uint32_t((c >> 8) & 0xFFu);
  return (a << 24) | (udiv255(r * a) << 16) | (udiv255(g * a) << 8) | udiv255(b * a);
}

__global__ void __launch_bounds__(128) k_build_luts(LutParams P) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= P.request_count) return;
  const b2dgpu_lut_request rq = P.requests[w];
  const b2dgpu_gradient_stop* stops = P.stops + rq.stop_offset;
  uint32_t* d = P.tables + P.table_offsets[w];
  const uint32_t size = rq.lut_size, n = rq.stop_count;

  uint64_t c0 = stops[0].rgba64, c1 = c0;
  uint32_t u0 = 0;
  uint32_t si = (stops[0].offset == 0.0 && n > 1) ? 1u : 0u;
  const uint32_t d_size = size - 1u;
  const double f_width = double(int32_t(d_size << 8));
  uint32_t pos = 0;                                       // `dp - d` of the reference after the last span

  do {
    c1 = stops[si].rgba64;
    const uint32_t u1 = uint32_t(__double2int_rn(stops[si].offset * f_width));
    pos = u0 >> 8;
    const uint32_t i = (u1 >> 8) - (u0 >> 8);
    u0 = u1;
    if (i <= 1u) {
      if (lane == 0) {
        d[pos] = lut_premul_top8(c0);
        if (i == 1u) d[pos + 1u] = lut_premul_top8(c1);
      }
      pos += i == 0u ? 1u : 2u;
    }
    else {
      const uint32_t cnt = i + 1u;
      const double scale = double(1 << 23) / double(int(i));
      uint32_t cx[4]; uint32_t dx[4];
      #pragma unroll
      for (int ch = 0; ch < 4; ch++) {
        const int sh = 8 + ch * 16;                       // top byte of each 16-bit channel: b, g, r, a
        const int32_t a0 = int32_t((c0 >> sh) & 0xFFu), a1 = int32_t((c1 >> sh) & 0xFFu);
        dx[ch] = uint32_t(__double2int_rz(double(a1 - a0) * scale));          // cvttpd2dq
        cx[ch] = (uint32_t(a0) << 23) + (1u << 22);
      }
      for (uint32_t k = lane; k < cnt; k += 32) {
        const uint32_t b = (cx[0] + k * dx[0]) >> 23, g = (cx[1] + k * dx[1]) >> 23, r = (cx[2] + k * dx[2]) >> 23, a = (cx[3] + k * dx[3]) >> 23;
        d[pos + k] = (a << 24) | (udiv255(r * a) << 16) | (udiv255(g * a) << 8) | udiv255(b * a);
      }
      pos += cnt;
    }
    c0 = c1;
    __syncwarp();                                         // the next span may overwrite this span's last entry
  } while (++si < n);

  const uint32_t last = lut_premul_top8(c0);
  for (uint32_t k = pos + lane; k < size; k += 32) d[k] = last;
  __syncwarp();
  if (lane == 0) d[0] = lut_premul_top8(stops[0].rgba64);
}

// =================================================================================================================
// K1 - edge builder
// =================================================================================================================

struct CountOut {
  enum : bool { kOrdered = false };
  uint32_t n;
  __device__ __forceinline__ void edge(int, int, int, int) { n++; }
};

// Write pass: the edges of ONE segment go to consecutive slots starting at `dst`.  What the reference emits once per
// curve or per piece (borders, the vertical-line case) takes its slot from a counter the warp shares; the lines of the
// node walks are written in CURVE ORDER (lane_dst: the lane's own run of slots, from a counting walk + warp scan), so
// that consecutive edges stay neighbours on the canvas - the tile compositor classifies edges 32 at a time and is
// measurably slower (14.5 vs 13.9 ms per config-1 step) when a curve's edges are shuffled.
struct WriteOut {
  enum : bool { kOrdered = true };
  b2dgpu_edge* dst;
  uint32_t* cursor;                  // shared memory (a local variable cannot be the target of an atomic)
  b2dgpu_edge* lane_dst;
  uint32_t n;
  int min_x, min_y, max_x, max_y;
  __device__ __forceinline__ void edge(int x0, int y0, int x1, int y1) {
    b2dgpu_edge e; e.x0 = x0; e.y0 = y0; e.x1 = x1; e.y1 = y1;
    if (lane_dst) lane_dst[n++] = e;
    else dst[atomicAdd(cursor, 1u)] = e;
    min_x = min(min_x, min(x0, x1)); max_x = max(max_x, max(x0, x1));
    min_y = min(min_y, min(y0, y1)); max_y = max(max_y, max(y0, y1));
  }
};

// A line segment is written by its own lane: private slot counter.
struct WriteOutLane {
  b2dgpu_edge* dst;
  uint32_t n;
  int min_x, min_y, max_x, max_y;
  __device__ __forceinline__ void edge(int x0, int y0, int x1, int y1) {
    b2dgpu_edge e; e.x0 = x0; e.y0 = y0; e.x1 = x1; e.y1 = y1;
    dst[n++] = e;
    min_x = min(min_x, min(x0, x1)); max_x = max(max_x, max(x0, x1));
    min_y = min(min_y, min(y0, y1)); max_y = max(max_y, max(y0, y1));
  }
};

enum : int { kBigCurve = 48 };        // control polygon wider / taller than this many pixels: flattened by the whole warp

// Per-warp scratch of the edge builder.
struct K1Warp {
  P2 spline[8 * 3 + 1];              // monotone pieces of the curve being flattened
  P2 node[2][kNodeCap][4];           // breadth-first frontier of its subdivision tree (double buffered)
  uint32_t meta[2][kNodeCap];
  uint32_t cursor;                   // write pass: edges of the current segment written so far
};

struct SegmentInput {
  P2 p[4];
  ClipBox cb;
  double tol_sq;
  uint32_t kind;
};

__device__ __forceinline__ SegmentInput load_segment(const BuildParams& P, uint32_t seg_index) {
  const b2dgpu_segment seg = P.segments[seg_index];
  const b2dgpu_command& cmd = P.commands[seg.command];
  const b2dgpu_geometry_state& gs = P.states[cmd.state_index];

  GeomXform xf;
  xf.m00 = gs.m[0]; xf.m01 = gs.m[1]; xf.m10 = gs.m[2]; xf.m11 = gs.m[3]; xf.m20 = gs.m[4]; xf.m21 = gs.m[5];
  xf.affine = gs.transform_type > 2u;

  SegmentInput in;
  in.cb.x0 = gs.clip[0]; in.cb.y0 = gs.clip[1]; in.cb.x1 = gs.clip[2]; in.cb.y1 = gs.clip[3];
  in.cb.ix0 = trunc_i(in.cb.x0); in.cb.ix1 = trunc_i(in.cb.x1);
  in.tol_sq = gs.tolerance_sq;

  const double2* v = reinterpret_cast<const double2*>(P.vertices);
  in.kind = seg.p1_kind & 3u;
  const uint32_t i1 = seg.p1_kind >> 2;
  const double2 a = v[seg.p0], b = v[i1];
  in.p[0] = xform(xf, mk(a.x, a.y));
  in.p[1] = xform(xf, mk(b.x, b.y));
  in.p[2] = in.p[1]; in.p[3] = in.p[1];
  if (in.kind == B2DGPU_SEG_CUBIC) {
    const double2 c = v[i1 + 1], d = v[i1 + 2];
    in.p[2] = xform(xf, mk(c.x, c.y));
    in.p[3] = xform(xf, mk(d.x, d.y));
  }
  else if (in.kind != B2DGPU_SEG_LINE) {
    // Quad, or conic: the reference flattens a conic through its quad machinery using vertex[+2] as the end point
    // (EdgeSourcePath::next_conic_to, edgebuilder_p.h:267-272; FlattenMonoConic :666-768).
    const double2 c = v[i1 + (in.kind == B2DGPU_SEG_CONIC ? 2u : 1u)];
    in.p[2] = xform(xf, mk(c.x, c.y));
  }
  return in;
}

__device__ __forceinline__ void set_lane_run(CountOut&, uint32_t) {}
__device__ __forceinline__ void set_lane_run(WriteOut& out, uint32_t offset) { out.lane_dst = out.dst + *out.cursor + offset; out.n = 0; }

// A warp flattens ONE curve (N = 3 quad / conic, 4 cubic).  Lane 0 does what the reference does once per curve (reject,
// monotone split); the lanes then take the pieces, expand them breadth first into at most kNodeCap nodes of the
// subdivision tree (node_split: the very halves the reference's depth-first walk would visit) and walk one node each
// (node_walk).  A canvas-sized curve of bl_bench / the tester flattens into ~80 lines: one thread walking it alone was
// what kept this kernel at 1.3 active lanes per instruction in round 1.
template<int N, typename Out>
__device__ __forceinline__ void flatten_curve_warp(const SegmentInput& in, K1Warp& W, int lane, Out& out) {
  int pieces = 0;
  uint32_t any = 0;
  if (lane == 0) {
    pieces = N == 3 ? prepare_quad(in.p[0], in.p[1], in.p[2], in.cb, out, W.spline, any)
                    : prepare_cubic(in.p[0], in.p[1], in.p[2], in.p[3], in.cb, out, W.spline, any);
  }
  pieces = __shfl_sync(0xFFFFFFFFu, pieces, 0);
  any = __shfl_sync(0xFFFFFFFFu, any, 0);
  if (!pieces) return;
  __syncwarp();

  MonoCurve<N> mc;
  mc.tol_sq = in.tol_sq;

  // roots: lane k sets up piece k
  bool have = false;
  uint32_t meta = 0;
  if (lane < pieces) have = piece_root<N>(mc, W.spline + lane * (N - 1), any != 0u, in.cb, out, meta);
  uint32_t mask = __ballot_sync(0xFFFFFFFFu, have);
  uint32_t count = __popc(mask);
  if (!count) return;
  int cur = 0;
  if (have) {
    const uint32_t pos = __popc(mask & ((1u << lane) - 1u));
    #pragma unroll
    for (int k = 0; k < N; k++) W.node[0][pos][k] = mc.p[k];
    W.meta[0][pos] = meta;
  }
  __syncwarp();

  // breadth-first expansion while the frontier can still double
  while (count * 2u <= uint32_t(kNodeCap)) {
    P2 first[4], second[4];
    bool split = false;
    if (uint32_t(lane) < count) {
      meta = W.meta[cur][lane];
      mc.begin_at(W.node[cur][lane], 0);
      split = node_split<N>(mc, meta & kNodePendingMask, first, second);
    }
    const uint32_t smask = __ballot_sync(0xFFFFFFFFu, split);
    if (!smask) break;
    if (uint32_t(lane) < count) {
      const uint32_t pos = uint32_t(lane) + __popc(smask & ((1u << lane) - 1u));
      if (split) {
        #pragma unroll
        for (int k = 0; k < N; k++) { W.node[cur ^ 1][pos][k] = first[k]; W.node[cur ^ 1][pos + 1][k] = second[k]; }
        W.meta[cur ^ 1][pos] = meta + 1u;
        W.meta[cur ^ 1][pos + 1] = meta;
      }
      else {
        #pragma unroll
        for (int k = 0; k < N; k++) W.node[cur ^ 1][pos][k] = mc.p[k];
        W.meta[cur ^ 1][pos] = meta;
      }
    }
    count += __popc(smask);
    cur ^= 1;
    __syncwarp();
  }

  if (Out::kOrdered) {
    CountOut cnt; cnt.n = 0;
    if (uint32_t(lane) < count) node_walk<N>(mc, W.node[cur][lane], W.meta[cur][lane], in.cb, cnt);
    uint32_t inc = cnt.n;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncwarp();
    set_lane_run(out, inc - cnt.n);                  // after what prepare / the roots emitted through the shared counter
  }
  if (uint32_t(lane) < count) node_walk<N>(mc, W.node[cur][lane], W.meta[cur][lane], in.cb, out);
  __syncwarp();
}

// K1: one thread per path segment, a warp per batch of 32 consecutive segments.  Lines are built by their own lane;
// the curves of the batch are then flattened one after the other by the whole warp.
//   WRITE == false: seg_counts[i] = edges of segment i.
//   WRITE == true : the edges are written at seg_offsets[i] .. seg_offsets[i + 1]; the command's bounding box grows.
template<bool WRITE>
__global__ void __launch_bounds__(128, 5) k_build_edges(BuildParams P) {
  __shared__ K1Warp s_warp[4];
  const int lane = threadIdx.x & 31;
  K1Warp& W = s_warp[threadIdx.x >> 5];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < P.segment_count;

  uint32_t kind = B2DGPU_SEG_LINE;
  uint32_t begin = 0, end = 0;
  bool writable = false;
  if (valid) {
    kind = P.segments[i].p1_kind & 3u;
    if (WRITE) {
      begin = P.seg_offsets[i]; end = P.seg_offsets[i + 1];
      writable = begin != end;
      if (writable && end > P.edge_capacity - P.edge_base) {       // capacity is checked on the host too; never write past it
        atomicOr(P.error_flag, 1u);
        writable = false;
      }
    }
  }
  const bool active = valid && (!WRITE || writable);

  // ---- lines and small curves: one per lane (a glyph-sized curve flattens into a handful of lines; sharing it over
  //      a warp would cost more than it saves) ----
  bool big = false;
  if (active) {
    const SegmentInput in = load_segment(P, i);
    if (kind != B2DGPU_SEG_LINE) {
      // size of the control polygon against kBigCurve pixels (24.8 units): NaNs compare false -> handled here
      const double lim = double(kBigCurve) * 256.0;
      #pragma unroll
      for (int k = 1; k < 4; k++) {
        const double dx = in.p[k].x - in.p[0].x, dy = in.p[k].y - in.p[0].y;
        big = big || dx > lim || dx < -lim || dy > lim || dy < -lim;
      }
    }
    if (!big) {
      if (!WRITE) {
        CountOut out; out.n = 0;
        if (kind == B2DGPU_SEG_LINE) build_line(in.p[0], in.p[1], in.cb, out);
        else if (kind == B2DGPU_SEG_CUBIC) build_cubic(in.p[0], in.p[1], in.p[2], in.p[3], in.cb, in.tol_sq, out);
        else build_quad(in.p[0], in.p[1], in.p[2], in.cb, in.tol_sq, out);
        P.seg_counts[i] = out.n;
      }
      else {
        WriteOutLane out;
        out.dst = P.edges + P.edge_base + begin; out.n = 0;
        out.min_x = out.min_y = INT_MAX; out.max_x = out.max_y = INT_MIN;
        if (kind == B2DGPU_SEG_LINE) build_line(in.p[0], in.p[1], in.cb, out);
        else if (kind == B2DGPU_SEG_CUBIC) build_cubic(in.p[0], in.p[1], in.p[2], in.p[3], in.cb, in.tol_sq, out);
        else build_quad(in.p[0], in.p[1], in.p[2], in.cb, in.tol_sq, out);
        if (out.n) {
          int4* bb = P.cmd_bbox_fixed + P.segments[i].command;
          atomicMin(&bb->x, out.min_x); atomicMin(&bb->y, out.min_y);
          atomicMax(&bb->z, out.max_x); atomicMax(&bb->w, out.max_y);
        }
      }
    }
  }

  // ---- big curves: the warp takes them one at a time ----
  uint32_t todo = __ballot_sync(0xFFFFFFFFu, big);
  if (!WRITE) {
    // segments that are skipped still need a count
    if (valid && !active) P.seg_counts[i] = 0;
  }
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t seg = blockIdx.x * blockDim.x + (threadIdx.x & ~31u) + uint32_t(src);
    const SegmentInput in = load_segment(P, seg);                  // every lane reads the same addresses: broadcast loads
    if (!WRITE) {
      CountOut out; out.n = 0;
      if (in.kind == B2DGPU_SEG_CUBIC) flatten_curve_warp<4>(in, W, lane, out);
      else flatten_curve_warp<3>(in, W, lane, out);
      const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, out.n);
      if (lane == 0) P.seg_counts[seg] = total;
    }
    else {
      if (lane == 0) W.cursor = 0;
      __syncwarp();
      WriteOut out;
      out.dst = P.edges + P.edge_base + __shfl_sync(0xFFFFFFFFu, begin, src); out.cursor = &W.cursor;
      out.lane_dst = nullptr; out.n = 0;
      out.min_x = out.min_y = INT_MAX; out.max_x = out.max_y = INT_MIN;
      if (in.kind == B2DGPU_SEG_CUBIC) flatten_curve_warp<4>(in, W, lane, out);
      else flatten_curve_warp<3>(in, W, lane, out);
      const int mnx = __reduce_min_sync(0xFFFFFFFFu, out.min_x), mny = __reduce_min_sync(0xFFFFFFFFu, out.min_y);
      const int mxx = __reduce_max_sync(0xFFFFFFFFu, out.max_x), mxy = __reduce_max_sync(0xFFFFFFFFu, out.max_y);
      if (lane == 0 && mnx <= mxx) {
        int4* bb = P.cmd_bbox_fixed + P.segments[seg].command;
        atomicMin(&bb->x, mnx); atomicMin(&bb->y, mny);
        atomicMax(&bb->z, mxx); atomicMax(&bb->w, mxy);
      }
    }
    __syncwarp();
  }
}

// =================================================================================================================
// K1b - exclusive scan (u32), 2048 items per block.
// =================================================================================================================
enum { kScanThreads = 256, kScanItems = 8, kScanBlock = kScanThreads * kScanItems };

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_warp, uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (kScanThreads / 32) ? s_warp[lane] : 0u;
    uint32_t winc = w;
    #pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (kScanThreads / 32)) s_warp[lane] = winc - w;
    if (lane == (kScanThreads / 32) - 1) s_warp[8] = winc;
  }
  __syncthreads();
  total = s_warp[8];
  return s_warp[warp] + inc - v;
}

// out[i] = exclusive prefix of in[0..n); block_sums[b] = sum of block b.  `out` has n + 1 entries (k_scan_add writes the total).
__global__ void __launch_bounds__(kScanThreads) k_scan_blocks(const uint32_t* in, uint32_t* out, uint32_t* block_sums, uint32_t n, const uint32_t* n_dev) {
  __shared__ uint32_t s_warp[9];
  if (n_dev) n = min(n, *n_dev);                       // the item count is only known on the device (k_bin_*)
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
  #pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    sum += v[k];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan_256(sum, s_warp, total);
  #pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < n) out[base + k] = off;
    off += v[k];
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_add(uint32_t* out, const uint32_t* block_offsets, uint32_t n, uint32_t* total_out, const uint32_t* n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  const uint32_t add = block_offsets[blockIdx.x];
  #pragma unroll
  for (int k = 0; k < kScanItems; k++)
    if (base + k < n) out[base + k] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // block_offsets has one extra entry holding the grand total (written by the level above).
    out[n] = block_offsets[gridDim.x];
    if (total_out) *total_out = block_offsets[gridDim.x];
  }
}

// Single-block scan for <= kScanBlock items; writes out[n] = total as well.
__global__ void __launch_bounds__(kScanThreads) k_scan_small(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total_out, const uint32_t* n_dev) {
  __shared__ uint32_t s_warp[9];
  if (n_dev) n = min(n, *n_dev);
  const uint32_t base = threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
  #pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    sum += v[k];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan_256(sum, s_warp, total);
  #pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < n) out[base + k] = off;
    off += v[k];
  }
  if (threadIdx.x == 0) {
    out[n] = total;
    if (total_out) *total_out = total;
  }
}

// =================================================================================================================
// K1c - command bounding boxes
// =================================================================================================================
__global__ void k_init_bbox(int4* bbox, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bbox[i] = make_int4(INT_MAX, INT_MAX, INT_MIN, INT_MIN);
}

// One warp per command with caller-supplied edges.
__global__ void __launch_bounds__(256) k_analytic_bbox(const b2dgpu_command* cmds, uint32_t ncmd, const b2dgpu_edge* edges, int4* bbox) {
  uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (c >= ncmd) return;
  const b2dgpu_command& cmd = cmds[c];
  if (cmd.type != B2DGPU_CMD_FILL_ANALYTIC) return;
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
  for (uint32_t i = lane; i < cmd.data_count; i += 32) {
    b2dgpu_edge e = edges[cmd.data_offset + i];
    mnx = min(mnx, min(e.x0, e.x1)); mxx = max(mxx, max(e.x0, e.x1));
    mny = min(mny, min(e.y0, e.y1)); mxy = max(mxy, max(e.y0, e.y1));
  }
  mnx = __reduce_min_sync(0xFFFFFFFFu, mnx); mny = __reduce_min_sync(0xFFFFFFFFu, mny);
  mxx = __reduce_max_sync(0xFFFFFFFFu, mxx); mxy = __reduce_max_sync(0xFFFFFFFFu, mxy);
  if (lane == 0) bbox[c] = make_int4(mnx, mny, mxx, mxy);
}

__global__ void k_finalize_commands(FinalizeParams P) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.command_count) return;
  const b2dgpu_command cmd = P.commands[c];
  uint32_t e_begin = 0, e_count = 0;
  if (cmd.type == B2DGPU_CMD_FILL_ANALYTIC) {
    e_begin = cmd.data_offset; e_count = cmd.data_count;
  }
  else if (cmd.type == B2DGPU_CMD_FILL_GEOMETRY) {
    uint32_t o0 = P.seg_offsets[cmd.data_offset];
    uint32_t o1 = P.seg_offsets[cmd.data_offset + cmd.data_count];
    e_begin = P.edge_base + o0; e_count = o1 - o0;
  }
  int4 bb = P.cmd_bbox_fixed[c];
  CmdBox box = command_pixel_box(cmd, e_count, bb.x, bb.y, bb.z, bb.w, P.width, P.y_begin, P.y_end);
  int x0 = box.x0, y0 = box.y0, x1 = box.x1, y1 = box.y1;
  P.cmd_bbox_px[c] = make_int4(x0, y0, x1, y1);
  P.cmd_edges[c] = make_uint2(e_begin, e_count);
}

// =================================================================================================================
// K1d - binning: per-band command lists with x-extents
//
// A band is one tile row (tile_h scanlines).  Band b gets the ordered list of the commands whose pixel box covers it,
// and every (band, command) cell records the columns the command can touch inside that band: a box fill its box, a
// shape the union of its edges' extents there.  The compositor culls with it: a tile left of the extent sees nothing, a
// tile right of it sees every edge of the band on its left, and the covers of a closed outline sum to zero on every
// scanline (EdgeBuilder closes figures and keeps clipped parts as border lines, edgebuilder_p.h:1058-1062, 2546-2622)
// - so both are skipped.  This is what the reference's per-band edge lists (edgestorage_p.h:38-178) and its band-by-
// band command walk (workerproc.cpp:166-255) achieve on the CPU: a tile looks at the commands of its band, not at the
// whole batch, and work follows the shape, not its bounding box.
//
//   k_bin_count    thread per command: bands covered -> cm_count, band_count (atomics)
//   scans          cm_base, band_off; k_bin_check compares the total with the buffer capacity
//   k_bin_fill     CTA per band: order-preserving compaction of the band's commands into its cells
//   k_bin_extents  warp per command: (edge, band) extents into the cells (atomic min / max)
// =================================================================================================================
__device__ __forceinline__ NormEdge load_edge(const int4* __restrict__ edges, uint32_t index) {
  int4 ev = __ldg(edges + index);
  b2dgpu_edge raw; raw.x0 = ev.x; raw.y0 = ev.y; raw.x1 = ev.z; raw.y1 = ev.w;
  return normalize_edge(raw);
}

__global__ void __launch_bounds__(256) k_bin_count(BinParams B) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B.command_count) return;
  const int4 bb = B.cmd_bbox_px[c];
  uint32_t nb = 0;
  if (bb.x < bb.z && bb.y < bb.w) {
    const int band0 = (bb.y - B.y_begin) / B.tile_h, band1 = (bb.w - 1 - B.y_begin) / B.tile_h;
    nb = uint32_t(band1 - band0 + 1);
    // difference array: +1 where the command's bands start, -1 after they end; its prefix sum is the number of commands
    // per band (two atomics per command whatever it spans)
    atomicAdd(B.band_count + band0, 1u);
    atomicAdd(B.band_count + band1 + 1, 0xFFFFFFFFu);
  }
  B.cm_count[c] = nb;
}

__global__ void k_bin_check(BinParams B) {
  const uint32_t total = B.cm_base[B.command_count];
  B.state[0] = total;
  B.state[1] = total <= B.capacity ? 1u : 0u;
  B.state[2] = 1u; B.state[3] = 0u;                     // overwritten by k_bin_check_edges when edge lists are built
}

__global__ void __launch_bounds__(1024) k_bin_fill(BinParams B) {
  if (!B.state[1]) return;
  __shared__ uint32_t s_w[32];
  const int b = blockIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int by0 = B.y_begin + b * B.tile_h, by1 = by0 + B.tile_h;
  uint32_t run = B.band_off[b];
  if (B.band_off[b + 1] == run) return;                  // nothing covers this band
  for (uint32_t base = 0; base < B.command_count; base += 1024) {
    const uint32_t c = base + threadIdx.x;
    bool hit = false;
    int4 bb = make_int4(0, 0, 0, 0);
    if (c < B.command_count) {
      bb = B.cmd_bbox_px[c];
      hit = bb.x < bb.z && bb.y < by1 && bb.w > by0;
    }
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
    if (lane == 0) s_w[warp] = __popc(ballot);
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    #pragma unroll
    for (uint32_t w = 0; w < 32; w++) { const uint32_t n = s_w[w]; if (w < warp) wbase += n; total += n; }
    if (hit) {
      const uint32_t pos = run + wbase + __popc(ballot & ((1u << lane) - 1u));
      B.cell_cmd[pos] = c;
      // a box's pixel box is exact; a shape starts empty and k_bin_extents widens it
      B.cell_ext[pos] = command_has_edges(B.commands[c].type) ? make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu) : make_uint2(uint32_t(bb.x), ~uint32_t(bb.z - 1));
      B.cm_index[B.cm_base[c] + uint32_t(b - (bb.y - B.y_begin) / B.tile_h)] = pos;
      if (B.band_edges) B.cell_edge_cnt[pos] = 0;
    }
    run += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_bin_extents(BinParams B) {
  if (!B.state[1]) return;
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (c >= B.command_count) return;
  const int4 bb = B.cmd_bbox_px[c];
  if (bb.x >= bb.z || bb.y >= bb.w) return;
  if (!command_has_edges(B.commands[c].type)) return;
  const int band0 = (bb.y - B.y_begin) / B.tile_h;
  const uint32_t* __restrict__ index = B.cm_index + B.cm_base[c];
  const uint2 er = B.cmd_edges[c];
  const int4* __restrict__ edges = reinterpret_cast<const int4*>(B.edges);
  for (uint32_t e = lane; e < er.y; e += 32) {
    const NormEdge ne = load_edge(edges, er.x + e);
    if (ne.y0 == ne.y1) continue;
    const int row_first = max(ne.y0 >> 8, bb.y), row_last = min((ne.y1 - 1) >> 8, bb.w - 1);
    if (row_first > row_last) continue;
    for (int b = (row_first - B.y_begin) / B.tile_h; b <= (row_last - B.y_begin) / B.tile_h; b++) {
      int lo, hi;
      band_edge_extent(ne, B.y_begin + b * B.tile_h, lo, hi, B.tile_h);
      const uint32_t ci = index[b - band0];
      uint32_t* cell = reinterpret_cast<uint32_t*>(B.cell_ext + ci);
      atomicMin(cell, uint32_t(lo));
      atomicMin(cell + 1, ~uint32_t(hi));
      if (B.band_edges) atomicAdd(B.cell_edge_cnt + ci, 1u);
    }
  }
}

// After the scan of the per-cell pair counts: do the edge lists fit?
__global__ void k_bin_check_edges(BinParams B) {
  if (!B.state[1]) { B.state[2] = 0; B.state[3] = 0; return; }
  const uint32_t pairs = B.cell_edge_off[B.state[0]];
  B.state[3] = pairs;
  B.state[2] = pairs <= B.edge_list_capacity ? 1u : 0u;
}

// Second pass over the (edge, band) pairs: every pair drops its edge index into its cell's list.  The counters run back
// down to zero, the order inside a list is irrelevant (u32 cover / area adds commute).
__global__ void __launch_bounds__(256) k_bin_edges(BinParams B) {
  if (!B.state[1] || !B.state[2]) return;
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (c >= B.command_count) return;
  const int4 bb = B.cmd_bbox_px[c];
  if (bb.x >= bb.z || bb.y >= bb.w) return;
  if (!command_has_edges(B.commands[c].type)) return;
  const int band0 = (bb.y - B.y_begin) / B.tile_h;
  const uint32_t* __restrict__ index = B.cm_index + B.cm_base[c];
  const uint2 er = B.cmd_edges[c];
  const int4* __restrict__ edges = reinterpret_cast<const int4*>(B.edges);
  for (uint32_t e = lane; e < er.y; e += 32) {
    const NormEdge ne = load_edge(edges, er.x + e);
    if (ne.y0 == ne.y1) continue;
    const int row_first = max(ne.y0 >> 8, bb.y), row_last = min((ne.y1 - 1) >> 8, bb.w - 1);
    if (row_first > row_last) continue;
    for (int b = (row_first - B.y_begin) / B.tile_h; b <= (row_last - B.y_begin) / B.tile_h; b++) {
      const uint32_t ci = index[b - band0];
      const uint32_t k = atomicSub(B.cell_edge_cnt + ci, 1u) - 1u;
      B.band_edges[B.cell_edge_off[ci] + k] = er.x + e;
    }
  }
}

// =================================================================================================================
// K2 + K3 - tile compositor
//
// One CTA of TH warps per 128 x TH destination tile.  A WARP owns a block of 32 columns x 4 rows of the tile (8 lanes
// per row, 4 consecutive pixels = one 16-byte vector per lane): warp w -> row group w % (TH / 4) (rows 4g .. 4g + 3),
// column block w / (TH / 4).  The block's pixels are loaded ONCE into registers, every command that touches the tile is replayed in
// submission order and the block is stored ONCE.  Blocks are what the work is decided on: phase 1 leaves, per command,
// one bit per warp ("this block can receive coverage"), the backdrop of every row at the start of each block, and one
// bit per (row, block) that says whether an edge crosses it - so a warp whose block lies outside a shape skips the
// command in a handful of instructions, a warp whose block lies inside takes the constant-mask path, and only the warps
// that really contain a crossing walk the row's entry list (the reference's FillAnalytic does the same with its bit
// vector of 4-pixel groups and its CMask / VMask spans, fillgeneric_p.h:174-388).
// =================================================================================================================

// Shared-memory cell row used by the slow path (u32 wrap-around adds: order independent).
struct SmemRowStore {
  uint32_t* cells;      // kTileW cells of ONE row (warp private while it is being rasterized)
  uint32_t* carry;      // that row's backdrop accumulator
  __device__ __forceinline__ void add_cell(int, int rel, uint32_t v) { atomicAdd(cells + rel, v); }
  __device__ __forceinline__ void add_carry(int, uint32_t v) { atomicAdd(carry, v); }
};

// Result of the classification / rasterization phase for one (command, tile) pair, kept in shared memory.
enum : int { kEntCap = 2, kRing = 2048, kBlockW = 32, kBlocks = kTileW / kBlockW, kBlockRows = 4 };
static_assert(kBlocks == 4, "a warp is addressed as (row group, one of four column blocks)");
// Per tile height TH (= warps per CTA): commands per phase-1 round, chained-entry pool size.
template<int TH> struct TileCfg { enum : int { kThreads = 32 * TH, kSub = TH == 32 ? 84 : 3 * TH, kPool = 128 * TH }; };
enum : uint32_t { kPreStraddle = 1u, kPreOverflow = 2u, kPreClipRight = 4u };   // ClipRight: the clipped box ends inside the tile
enum : uint32_t { kDenseItemsPerRow = 8u };        // (edge, row) crossings per tile row beyond which phase 1 gives up

template<int TH>
struct PreCmdT {
  uint32_t carry_left[TH];      // backdrop per tile row from the edges entirely left of the tile
  uint4 carry4[TH];             // .x: backdrop from straddling edges' cells left of the tile; .y/.z/.w: what the entries
                                // that lie entirely left of block 1 / 2 / 3 add to the backdrop from that block on
  uint32_t nent[TH];            // number of cell entries appended per row (beyond kEntCap: chained in the pool)
  uint32_t ovf_head[TH];        // 1 + pool index of the row's last chained entry, 0 = none
  uint32_t blk_has[(TH + 7) / 8]; // 4 bits per row: block b of the row holds an entry (or the spill of its left neighbour)
  uint32_t flags;               // phase 1 only (atomically updated); the replay reads hdr
  uint32_t warp_mask;           // bit w: warp w's block can receive coverage from this command (0: skipped by the replay)
  uint32_t wrec4[TH / 4];       // one byte per warp: the mask of EVERY pixel of the warp's block when it is the same for all
                                // of them (interior of a shape, inside of a box), 0 when the block needs the general path
  uint32_t pad_[(4 - ((TH + 7) / 8 + 2 + TH / 4) % 4) % 4];
  // What the replay needs of the command, decoded once by phase 1 (two 16-byte shared-memory loads per iteration):
  //   [0] type | flags << 8   [1] alpha   [2] fill rule mask   [3] solid colour
  //   [4] fetch type | comp op << 8 | source format << 16   [5] right end of the clipped box   [6,7] FetchData pointer
  uint32_t hdr[8];
  uint2 ent[TH][kEntCap];       // edge crossings: (cell relative to the tile | area << 8, (cover << 9) - area)
  // Per-row state of the gradient fetchers (fetch_row_ctx: 64-bit row origin of a linear gradient, the three floats of
  // a radial / conic row), computed once per (command, row) by the lane that owns the row in phase 1.
  uint32_t rowctx[TH][3];
  // The command itself (64 B): box coordinates and the mask index of the box fills.
  uint32_t cmd_words[sizeof(b2dgpu_command) / 4];
};
static_assert(sizeof(PreCmdT<8>) % 16 == 0 && sizeof(PreCmdT<16>) % 16 == 0 && sizeof(PreCmdT<32>) % 16 == 0, "PreCmd must keep 16-byte alignment of the staged blocks");

// Coverage sink of phase 1: the few cells a straddling edge touches in a row are appended to that row's entry list.
// One entry per edge crossing: cell `rel` gets v0 = (cover << 9) - area and cell rel + 1 gets `area`
// (cell_merge, analyticrasterizer_p.h:1202-1210), packed as x = rel | area << 8 (|area| <= 2^17), y = v0.
// A row with more crossings than its inline list holds (bl_bench's random polygons, map-like paths) chains the rest
// through a pool shared by the sub-chunk; s_pool_link[i] = 1 + index of the previous chained entry of the same row.
template<int TH>
struct EntrySink {
  PreCmdT<TH>* pre;
  uint32_t pool_cap;
  uint2* pool;
  uint16_t* pool_link;
  uint32_t* pool_next;
  int tx0;
  int row;
  __device__ __forceinline__ void append(int rel, uint32_t v0, uint32_t area) {
    const uint2 en = make_uint2(uint32_t(rel) | (area << 8), v0);
    uint32_t idx = atomicAdd(&pre->nent[row], 1u);
    if (idx < uint32_t(kEntCap)) pre->ent[row][idx] = en;
    else {
      const uint32_t pi = atomicAdd(pool_next, 1u);
      if (pi < pool_cap) {
        const uint32_t prev = atomicExch(&pre->ovf_head[row], pi + 1u);
        pool[pi] = en;
        pool_link[pi] = uint16_t(prev);
      }
      else atomicOr(&pre->flags, kPreOverflow);         // pool exhausted: the row re-rasterizes itself (slow path)
    }
    // Block bookkeeping.  The entry changes the coverage from pixel `rel` on (v0) and from pixel rel + 1 on (area): the
    // block that holds `rel` - and the next one when rel is a block's last pixel - applies it pixel by pixel, every
    // block further right just sees v0 + area = cover << 9 more backdrop.
    const int bi = rel >> 5;
    const bool spill = (rel & 31) == 31;
    uint32_t bits = 1u << bi;
    if (spill && bi < kBlocks - 1) bits |= 2u << bi;
    atomicOr(&pre->blk_has[row >> 3], bits << ((row & 7) * 4));
    const int first_full = bi + 1 + int(spill);
    if (first_full < kBlocks) atomicAdd(reinterpret_cast<uint32_t*>(&pre->carry4[row]) + first_full, v0 + area);
  }
  // window interface of edge_step_scanline<true> (dev_raster.cuh)
  __device__ __forceinline__ int win_lo() const { return tx0; }
  __device__ __forceinline__ int win_hi() const { return tx0 + kTileW; }
  __device__ __forceinline__ void add_left(uint32_t v) { if (v) atomicAdd(&pre->carry4[row].x, v); }
  __device__ __forceinline__ void merge(int x, uint32_t cover, uint32_t area) {
    const uint32_t v0 = (cover << 9) - area;
    const int rel = x - tx0;
    if (rel >= kTileW || (v0 | area) == 0u) return;
    if (rel >= 0) append(rel, v0, area);
    else {
      // cell x is left of the tile: it only feeds the row's backdrop; cell x + 1 may be the tile's first column
      if (rel == -1) { if (v0) atomicAdd(&pre->carry4[row].x, v0); if (area) append(0, area, 0u); }
      else atomicAdd(&pre->carry4[row].x, v0 + area);
    }
  }
};

// Masks of a lane's four pixels for a FillBoxMaskA command (image masks are not in the bench's configurations: cold).
__device__ __noinline__ uint4 box_mask_a_row(const b2dgpu_command& cmd, const b2dgpu_pattern_source& ms, int x, int y) {
  return make_uint4(box_mask_a(cmd, ms, x, y), box_mask_a(cmd, ms, x + 1, y), box_mask_a(cmd, ms, x + 2, y), box_mask_a(cmd, ms, x + 3, y));
}

// Slow path of the replay (a command with more crossings in this tile than the entry lists hold: dense polygons,
// nearly horizontal edges).  Called by the four warps of a row group together: warp `b` rasterizes row 4g + b of the
// tile into that row's shared-memory cells and turns them, in place, into the running coverage of the reference's
// scanline walk (fillgeneric_p.h:285-297); after the group's barrier every warp reads its own block of all four rows.
// `elist` != nullptr: the edges are elist[0 .. er.y) (the band's list of the command), otherwise er.x .. er.x + er.y.
// Out of line: it is rare and large.
__device__ __noinline__ void slow_group_rows(const int4* __restrict__ edges, uint2 er, const uint32_t* __restrict__ elist, int tx0, int ty0, int row, int lane,
                                             uint32_t* cells_row, uint32_t* carry_row, uint32_t base, int tile_h, int group) {
  *reinterpret_cast<uint4*>(cells_row + lane * 4) = make_uint4(0, 0, 0, 0);
  if (lane == 0) *carry_row = 0;
  __syncwarp();
  SmemRowStore store{ cells_row, carry_row };
  TileSink<SmemRowStore> sink(store, tx0);
  sink.row = row;
  const int py = ty0 + row;
  for (uint32_t e = lane; e < er.y; e += 32) {
    NormEdge ne = load_edge(edges, elist ? __ldg(elist + e) : er.x + e);
    if (py >= (ne.y0 >> 8) && py <= ((ne.y1 - 1) >> 8) && tile_edge_class(ne, tx0, ty0, tile_h) == kEdgeStraddle)
      tile_rasterize_edge_row(ne, py, sink);
  }
  __syncwarp();
  uint4 c = *reinterpret_cast<uint4*>(cells_row + lane * 4);
  c.y += c.x; c.z += c.y; c.w += c.z;
  uint32_t inc = c.w;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  const uint32_t add = base + *carry_row + (inc - c.w);
  *reinterpret_cast<uint4*>(cells_row + lane * 4) = make_uint4(c.x + add, c.y + add, c.z + add, c.w + add);
  asm volatile("bar.sync %0, 128;" :: "r"(group + 1) : "memory");
}

#ifdef B2D_PHASE_TIMING
// Experiment build only (make EXTRA=-DB2D_PHASE_TIMING): cycles the warps of k_tile_render spend in / wait for each phase.
//   [0] phase-1 busy (sum over warps)  [1] wait at the barrier behind phase 1  [2] phase-2 busy  [3] cull + wait before phase 1
//   [4] phase-1 length seen by warp 0 (barrier to barrier)  [5] sub-chunks  [6] commands replayed (per CTA)  [7] longest phase-1 command (sum over sub-chunks)
__device__ unsigned long long g_phase_cycles[16];   // [8] cycles in item rounds [9] rounds [10] cycles classifying chunks [11] chunks [12] finalize cycles [13] prologue cycles [14] cycles inside commands (all by lane 0 of every warp)
extern "C" __attribute__((visibility("default"))) int b2dgpu_debug_phase_cycles(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(g_phase_cycles));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
  return 0;
}
__device__ __forceinline__ long long pt_now() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
#define PT(...) __VA_ARGS__
#else
#define PT(...)
#endif

template<int BPP, int TH>
__global__ void __launch_bounds__(32 * TH, 32 / TH) k_tile_render(TileParams P) {
  using PreCmd = PreCmdT<TH>;
  constexpr int kThreads = TileCfg<TH>::kThreads, kSub = TileCfg<TH>::kSub, kPool = TileCfg<TH>::kPool;
  __shared__ __align__(16) uint32_t s_cells[TH][kTileW];     // slow path only
  __shared__ uint32_t s_carry[TH];
  __shared__ uint32_t s_list[kRing];
  extern __shared__ __align__(16) uint8_t s_dynamic[];             // kSub PreCmd records (dynamic: > 48 KB)
  PreCmd* const s_pre = reinterpret_cast<PreCmd*>(s_dynamic);
  uint2* const s_pool = reinterpret_cast<uint2*>(s_dynamic + sizeof(PreCmd) * kSub);      // kPool chained entries
  uint16_t* const s_pool_link = reinterpret_cast<uint16_t*>(s_pool + kPool);
  __shared__ uint32_t s_wmask[kSub];                               // copy of PreCmd::warp_mask, densely packed
  __shared__ uint32_t s_wcount[TH];
  __shared__ uint32_t s_next;
  __shared__ uint32_t s_pool_next;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int tile_x = blockIdx.x % P.tiles_x;
  const int tile_y = blockIdx.x / P.tiles_x;
  const int tx0 = tile_x * kTileW;
  const int ty0 = P.y_begin + tile_y * TH;          // absolute y of the tile's first row
  // Replay geometry of this lane: block (grp, blk) of the tile, row `row` of the tile, pixels px .. px + 3.
  // Warp w -> row group w % G, column block w / G (G = TH / 4 groups): the warp schedulers take warps round robin
  // (w % 4), so each of them gets blocks from all four columns of the tile and from row groups that lie apart -
  // whatever part of the tile a shape covers, its work is spread over the four schedulers.
  constexpr int kGroups = TH / kBlockRows;
  const int grp = warp % kGroups, blk = warp / kGroups;
  const int row = grp * kBlockRows + (lane >> 3);
  const int bx0 = tx0 + blk * kBlockW;              // first column of the warp's block
  const int px = bx0 + (lane & 7) * 4;
  const int py = ty0 + row;
  const int4* __restrict__ edges = reinterpret_cast<const int4*>(P.edges);

  // Load the destination once: 4 consecutive pixels per lane, kept in registers for the whole command list.
  uint8_t* dst_row = P.dst + size_t(py - P.y_begin) * P.dst_stride;
  uint32_t d[4];
  if (BPP == 4) {
    uint4 v0 = *reinterpret_cast<const uint4*>(dst_row + size_t(px) * 4);
    d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w;
  }
  else {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(dst_row + px);
    #pragma unroll
    for (int i = 0; i < 4; i++) d[i] = ((v >> (8 * i)) & 0xFFu) * 0x01010101u;
  }
  bool dirty = false;
  uint32_t px_written = 0;
  PT(long long pt_busy1 = 0; long long pt_wait1 = 0; long long pt_busy2 = 0; long long pt_wait2 = 0; long long pt_len1 = 0; long long pt_sub = 0; long long pt_cmds = 0; long long pt_maxcmd = 0; long long pt_last = pt_now(); long long pt_round = 0, pt_rounds = 0, pt_cls = 0, pt_chunks = 0, pt_fin = 0, pt_pro = 0, pt_in = 0; __shared__ unsigned long long s_pt_max; __shared__ unsigned long long s_pt_maxround;)
  const bool count_pixels = P.pixel_counter != nullptr;

  // Commands that touch the tile are appended, in order, to a ring in shared memory; whenever kSub of them are
  // pending (or the command list ends) they are classified (phase 1) and replayed (phase 2).
  uint32_t ring_head = 0, ring_tail = 0;                // block-uniform
  // The commands to look at: the band's list (k_bin_*), or - when the lists were not built - every command.
  const bool binned = P.bin_state != nullptr && P.bin_state[1] != 0u;
  const uint32_t list_begin = binned ? P.band_off[tile_y] : 0u;
  const uint32_t list_count = binned ? P.band_off[tile_y + 1] - list_begin : P.command_count;
  const bool edge_lists = binned && P.band_edges != nullptr && P.bin_state[2] != 0u;
  for (uint32_t base = 0; base < list_count || ring_head != ring_tail; base += kThreads) {
    // ---- cull: which of the next kThreads candidates touch this tile? (order preserving compaction) ----
    if (base < list_count) {
      uint32_t c = 0;
      bool hit = false;
      if (base + tid < list_count) {
        if (binned) {
          const uint2 ex = P.cell_ext[list_begin + base + tid];
          hit = uint32_t(tx0 + kTileW) > ex.x && uint32_t(tx0) <= ~ex.y;
          c = list_begin + base + tid;                  // the ring holds the CELL; the command is cell_cmd[cell]
        }
        else {
          c = base + tid;
          int4 bb = P.cmd_bbox_px[c];
          hit = bb.x < tx0 + kTileW && bb.z > tx0 && bb.y < ty0 + TH && bb.w > ty0;
        }
      }
      uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
      __syncthreads();                                  // everybody is done reading s_wcount of the previous chunk
      if (lane == 0) s_wcount[warp] = __popc(ballot);
      __syncthreads();
      uint32_t wbase = 0, total = 0;
      #pragma unroll
      for (int w = 0; w < TH; w++) {
        uint32_t cnt = s_wcount[w];
        if (w < warp) wbase += cnt;
        total += cnt;
      }
      if (hit) s_list[(ring_tail + wbase + __popc(ballot & ((1u << lane) - 1u))) & (kRing - 1)] = c;
      ring_tail += total;
    }
    const bool last = base + kThreads >= list_count;

    while (ring_tail - ring_head >= uint32_t(kSub) || (last && ring_tail != ring_head)) {
      const uint32_t sub = ring_head;
      const uint32_t sub_n = min(uint32_t(kSub), ring_tail - ring_head);
      ring_head += sub_n;
      if (tid == 0) { s_next = TH; s_pool_next = 0; } // commands beyond the first TH are handed out dynamically
      __syncthreads();                                  // ring entries written / previous sub-chunk's s_pre consumed
      PT(const long long pt_a = pt_now(); pt_wait2 += pt_a - pt_last; pt_sub++; pt_cmds += sub_n; if (tid == 0) { s_pt_max = 0; s_pt_maxround = 0; })

      // ---- phase 1 (K2): one warp per command - classify its edges against the tile and rasterize the few that
      //      straddle it, one (edge, row) item per lane, into the command's per-row entry lists.  No block barrier.
      for (uint32_t k = warp; k < sub_n; ) {
        PT(const long long pt_c0 = pt_now();)
        const uint32_t cell = s_list[(sub + k) & (kRing - 1)];
        const uint32_t ci = binned ? __ldg(P.cell_cmd + cell) : cell;
        PreCmd* pre = &s_pre[k];
        if (lane < TH) { pre->carry4[lane] = make_uint4(0, 0, 0, 0); pre->nent[lane] = 0; pre->ovf_head[lane] = 0; }
        if (lane < (TH + 7) / 8) pre->blk_has[lane] = 0;
        const int4 bbp = P.cmd_bbox_px[ci];
        if (lane == 0) pre->flags = bbp.z < tx0 + kTileW ? kPreClipRight : 0u;
        {
          // stage the command (one coalesced 64-byte load)
          const uint32_t* src = reinterpret_cast<const uint32_t*>(P.commands + ci);
          uint32_t w = lane < int(sizeof(b2dgpu_command) / 4) ? __ldg(src + lane) : 0u;
          if (lane < int(sizeof(b2dgpu_command) / 4)) pre->cmd_words[lane] = w;
        }
        __syncwarp();

        uint32_t left_acc = 0;                        // lane r < TH: backdrop of tile row r from the edges left of the tile
        uint32_t nstr = 0, items = 0;                 // straddling edges / their (edge, row) crossings in this tile
        const bool is_box = !command_has_edges(pre->cmd_words[0]);

        if (!is_box) {
          // The edges to look at: the cell's list (the edges of the command that cross this band) or, without lists,
          // every edge of the command.
          uint2 er = P.cmd_edges[ci];
          const uint32_t* __restrict__ elist = nullptr;
          if (edge_lists) {
            const uint32_t o0 = __ldg(P.cell_edge_off + cell), o1 = __ldg(P.cell_edge_off + cell + 1);
            elist = P.band_edges + o0;
            er.y = o1 - o0;
          }
          EntrySink<TH> sink; sink.pre = pre; sink.pool_cap = uint32_t(kPool); sink.pool = s_pool; sink.pool_link = s_pool_link; sink.pool_next = &s_pool_next; sink.tx0 = tx0; sink.row = 0;
          PT(pt_pro += pt_now() - pt_c0;)
          for (uint32_t e0 = 0; e0 < er.y; e0 += 32) {
            PT(const long long pt_k0 = pt_now(); pt_chunks++;)
            const uint32_t e = e0 + lane;
            int cls = kEdgeNone;
            uint32_t rows_crossed = 0, first_row = 0;
            uint32_t extra_cells = 0;                   // cells beyond one per row that a shallow edge will touch in the tile (estimate)
            int ey0 = 0, ey1 = 0;
            uint32_t esign = 0;
            uint32_t eidx = 0;
            if (e < er.y) {
              eidx = elist ? __ldg(elist + e) : er.x + e;
              NormEdge ne = load_edge(edges, eidx);
              cls = tile_edge_class(ne, tx0, ty0, TH);
              ey0 = ne.y0; ey1 = ne.y1; esign = ne.sign_bit;
              if (cls == kEdgeStraddle) {
                first_row = uint32_t(max(ne.y0 >> 8, ty0) - ty0);
                rows_crossed = uint32_t(min((ne.y1 - 1) >> 8, ty0 + TH - 1) - ty0) - first_row + 1u;
                // a nearly horizontal edge is one entry per CELL, walked by one lane: count the columns it spans inside the
                // tile's rows so that such a command takes the cell-row path of the replay instead (kDenseItemsPerRow)
                const uint32_t edge_rows = uint32_t(((ne.y1 - 1) >> 8) - (ne.y0 >> 8)) + 1u;
                const uint32_t cols = min(uint32_t(kTileW), (uint32_t(abs(ne.x1 - ne.x0)) >> 8) * rows_crossed / edge_rows);
                extra_cells = cols > rows_crossed ? cols - rows_crossed : 0u;
              }
            }
            nstr += __popc(__ballot_sync(0xFFFFFFFFu, cls == kEdgeStraddle));
            // Edges entirely left of the tile: every scanline's cells of an edge sum to (cover << 9), cover = its signed
            // y-extent inside the row.  One edge at a time is broadcast and every lane adds the cover of ITS row.
            uint32_t lb = __ballot_sync(0xFFFFFFFFu, cls == kEdgeLeft);
            while (lb) {
              const int src = __ffs(lb) - 1;
              lb &= lb - 1;
              const int y0 = __shfl_sync(0xFFFFFFFFu, ey0, src), y1 = __shfl_sync(0xFFFFFFFFu, ey1, src);
              const uint32_t sg = __shfl_sync(0xFFFFFFFFu, esign, src);
              const int yt = (ty0 + lane) << 8;
              const int cov = min(y1, yt + 256) - max(y0, yt);
              if (cov > 0 && lane < TH) left_acc += uint32_t(sg ? -cov : cov) << 9;
            }
            // (edge, row) items of this chunk, packed densely over the lanes: inclusive scan of the rows each straddling
            // edge crosses; item i belongs to the first edge whose inclusive count exceeds i.
            uint32_t inc = rows_crossed;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
              if (lane >= o) inc += t;
            }
            const uint32_t chunk_items = __shfl_sync(0xFFFFFFFFu, inc, 31);
            items += chunk_items + __reduce_add_sync(0xFFFFFFFFu, extra_cells);
            // So many crossings that the entry lists and the pool would overflow anyway: do not rasterize here, every
            // row of the replay is rasterized as a whole (slow_group_rows).
            PT(pt_cls += pt_now() - pt_k0;)
            if (items > kDenseItemsPerRow * uint32_t(TH)) continue;
            for (uint32_t base_i = 0; base_i < chunk_items; base_i += 32) {
              PT(const long long pt_r0 = pt_now(); pt_rounds++;)
              const uint32_t i = base_i + lane;
              uint32_t j = 0;
              #pragma unroll
              for (uint32_t step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, inc, int(j + step - 1u));
                if (v <= i) j += step;
              }
              const uint32_t j_inc = __shfl_sync(0xFFFFFFFFu, inc, int(j & 31u));
              const uint32_t j_rows = __shfl_sync(0xFFFFFFFFu, rows_crossed, int(j & 31u));
              const uint32_t j_first = __shfl_sync(0xFFFFFFFFu, first_row, int(j & 31u));
              const uint32_t j_eidx = __shfl_sync(0xFFFFFFFFu, eidx, int(j & 31u));
              if (i < chunk_items) {
                const int r = int(j_first + (i - (j_inc - j_rows)));
                NormEdge ne = load_edge(edges, j_eidx);
                sink.row = r;
                tile_rasterize_edge_row(ne, ty0 + r, sink);
              }
              PT(__syncwarp(); { const long long dr = pt_now() - pt_r0; pt_round += dr; if (lane == 0) atomicMax(&s_pt_maxround, (unsigned long long)dr); })
            }
          }
        }
        PT(const long long pt_f0 = pt_now();)
        if (lane < TH) pre->carry_left[lane] = left_acc;
        const bool dense = nstr && items > kDenseItemsPerRow * uint32_t(TH);
        if (lane == 0 && nstr) atomicOr(&pre->flags, dense ? (kPreStraddle | kPreOverflow) : kPreStraddle);
        __syncwarp();
        {
          const b2dgpu_command& c = *reinterpret_cast<const b2dgpu_command*>(pre->cmd_words);
          const uint32_t sig = c.signature;
          const uint32_t ft = B2DGPU_SIG_FETCH_TYPE(sig);
          const b2dgpu_fetch_data* fd = P.fetch_data + c.fetch_index;
          if (lane == 0) {
            pre->hdr[0] = c.type | (pre->flags << 8); pre->hdr[1] = c.alpha; pre->hdr[2] = c.fill_rule_mask; pre->hdr[3] = c.solid_prgb32;
            pre->hdr[4] = ft | (B2DGPU_SIG_COMP_OP(sig) << 8) | (B2DGPU_SIG_SRC_FORMAT(sig) << 16);
            pre->hdr[5] = uint32_t(bbp.z);
            pre->hdr[6] = uint32_t(uintptr_t(fd)); pre->hdr[7] = uint32_t(uintptr_t(fd) >> 32);
          }
          if (lane < TH && ft >= B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD) {
            const RowCtx3 rc = fetch_row_ctx(ft, *fd, uint32_t(ty0 + lane));
            pre->rowctx[lane][0] = rc.a; pre->rowctx[lane][1] = rc.b; pre->rowctx[lane][2] = rc.c;
          }
        }

        // Which warps of the replay does this command concern, and how?  Lane w answers for warp w = block (w / 4, w % 4):
        // skipped (nothing to composite), uniform (every pixel of the block has the same mask: the inside of a shape or
        // of a box - the replay then needs neither the backdrop nor the entries), or general.
        {
          bool any = false;
          uint32_t uniform_mask = 0;
          if (lane < TH) {
            const b2dgpu_command& c = *reinterpret_cast<const b2dgpu_command*>(pre->cmd_words);
            const int g = lane % kGroups, b = lane / kGroups;
            const int r0 = g * kBlockRows;
            const int bxl = tx0 + b * kBlockW;
            const bool rows_in = bbp.y < ty0 + r0 + kBlockRows && bbp.w > ty0 + r0;
            if (is_box) {
              any = rows_in && bbp.x < bxl + kBlockW && bbp.z > bxl;
              if (any && c.type == B2DGPU_CMD_FILL_BOX_A && c.box[0] <= bxl && c.box[2] >= bxl + kBlockW &&
                  c.box[1] <= ty0 + r0 && c.box[3] >= ty0 + r0 + kBlockRows) uniform_mask = c.alpha;
            }
            else if (pre->flags & kPreOverflow) any = rows_in;          // the four warps of a group meet at a barrier: same answer for all
            else if (rows_in && bbp.z > bxl) {
              bool uni = bbp.z >= bxl + kBlockW;                        // the clipped box does not end inside the block
              uint32_t m0 = 0;
              #pragma unroll
              for (int r = r0; r < r0 + kBlockRows; r++) {
                const uint4 c4 = pre->carry4[r];
                const uint32_t carry = pre->carry_left[r] + c4.x + (b >= 1 ? c4.y : 0u) + (b >= 2 ? c4.z : 0u) + (b >= 3 ? c4.w : 0u);
                const bool has = (pre->blk_has[r >> 3] >> ((r & 7) * 4 + b)) & 1u;
                const uint32_t mr = calc_mask((256u << 9) + carry, c.fill_rule_mask, c.alpha);
                any = any || has || mr != 0u;
                uni = uni && !has && (r == r0 || mr == m0);
                if (r == r0) m0 = mr;
              }
              if (any && uni) uniform_mask = m0;
            }
          }
          const uint32_t wm = __ballot_sync(0xFFFFFFFFu, any);
          if (lane == 0) { pre->warp_mask = wm; s_wmask[k] = wm; }
          if (lane < TH) reinterpret_cast<uint8_t*>(pre->wrec4)[lane] = uint8_t(uniform_mask);
        }
        PT(pt_fin += pt_now() - pt_f0; pt_in += pt_now() - pt_c0;)
        PT(if (lane == 0) atomicMax(&s_pt_max, (unsigned long long)(pt_now() - pt_c0));)
        // next command: whichever warp is free takes it (edge counts differ a lot between commands)
        uint32_t nk = 0;
        if (lane == 0) nk = atomicAdd(&s_next, 1u);
        k = __shfl_sync(0xFFFFFFFFu, nk, 0);
      }
      PT(const long long pt_b = pt_now();)
      __syncthreads();
      PT(const long long pt_c = pt_now(); pt_busy1 += pt_b - pt_a; pt_wait1 += pt_c - pt_b; pt_len1 += pt_c - pt_a; pt_maxcmd += (long long)s_pt_max; if (tid == 0) atomicAdd(&g_phase_cycles[15], s_pt_maxround);)

      // ---- phase 2 (K3): every warp replays, in order, the commands that concern ITS block; warps never wait for each
      //      other (except the four warps of a row group inside the slow path) ----
      #pragma unroll 1
      for (uint32_t kb = 0; kb < sub_n; kb += 32) {
      uint32_t act = __ballot_sync(0xFFFFFFFFu, kb + uint32_t(lane) < sub_n && ((s_wmask[min(kb + uint32_t(lane), uint32_t(kSub - 1))] >> warp) & 1u));
      dirty = dirty || act != 0u;                       // warp uniform: the block is stored if any command concerned it
      while (act) {
        const uint32_t k = kb + uint32_t(__ffs(act) - 1);
        act &= act - 1;
        const PreCmd& pre = s_pre[k];
        const b2dgpu_command& cmd = *reinterpret_cast<const b2dgpu_command*>(pre.cmd_words);
        const uint4 h0 = *reinterpret_cast<const uint4*>(&pre.hdr[0]);
        const uint4 h1 = *reinterpret_cast<const uint4*>(&pre.hdr[4]);
        const uint32_t type = h0.x & 0xFFu;
        const uint32_t alpha = h0.y;
        uint32_t m[4] = { 0, 0, 0, 0 };
        bool opaque;
        const uint32_t uniform_mask = reinterpret_cast<const uint8_t*>(pre.wrec4)[warp];

        if (uniform_mask) {
          // The whole block lies inside the shape (FillAnalytic's CMask spans, fillgeneric_p.h:300-330) or the box.
          m[0] = m[1] = m[2] = m[3] = uniform_mask;
          opaque = uniform_mask == 255u;
        }
        else {
        if (type == B2DGPU_CMD_FILL_BOX_A) {
          // FillBoxA_Base (fillgeneric_p.h:22-65): constant mask inside the box.
          if (py >= cmd.box[1] && py < cmd.box[3]) {
            #pragma unroll
            for (int i = 0; i < 4; i++) m[i] = (px + i >= cmd.box[0] && px + i < cmd.box[2]) ? alpha : 0u;
          }
        }
        else if (type == B2DGPU_CMD_FILL_BOX_U) {
          BoxUParams bu = box_u_setup(cmd.box, alpha);
          #pragma unroll
          for (int i = 0; i < 4; i++) m[i] = box_u_mask(bu, px + i, py);
        }
        else if (type == B2DGPU_CMD_FILL_BOX_MASK_A) {
          const uint4 mk = box_mask_a_row(cmd, P.fetch_data[cmd.reserved[0]].pattern.src, px, py);
          m[0] = mk.x; m[1] = mk.y; m[2] = mk.z; m[3] = mk.w;
        }
        else {
          const uint32_t flags = h0.x >> 8;
          const uint32_t rule = h0.z;
          if (!(flags & kPreOverflow)) {
            // Backdrop of the lane's row at the first pixel of the block.
            const uint4 c4 = pre.carry4[row];
            const uint32_t carry = (256u << 9) + pre.carry_left[row] + c4.x + (blk >= 1 ? c4.y : 0u) + (blk >= 2 ? c4.z : 0u) + (blk >= 3 ? c4.w : 0u);
            const bool has = (pre.blk_has[row >> 3] >> ((row & 7) * 4 + blk)) & 1u;
            if (!__any_sync(0xFFFFFFFFu, has)) {
              // No edge inside the block: coverage is constant along each of its rows.
              const uint32_t mm = calc_mask(carry, rule, alpha);
              m[0] = mm; m[1] = mm; m[2] = mm; m[3] = mm;
            }
            else {
              // The running sum of fillgeneric_p.h:285-297 at pixel x is the backdrop plus every entry at or left of x
              // (u32 adds commute), so no prefix scan is needed: each entry of the row that belongs to this block is
              // added to the pixels from its cell onwards; the entries left of the block are in `carry` already.
              uint32_t cov0 = carry, cov1 = carry, cov2 = carry, cov3 = carry;
              const uint32_t n = has ? pre.nent[row] : 0u;
              const uint32_t n_inline = min(n, uint32_t(kEntCap));
              uint32_t chain = n > uint32_t(kEntCap) ? pre.ovf_head[row] : 0u;
              const int lo = blk * kBlockW - 1;                       // cells from here on are applied pixel by pixel
              const int lane_x = blk * kBlockW + (lane & 7) * 4;      // the lane's first pixel, relative to the tile
              for (uint32_t j = 0; __any_sync(0xFFFFFFFFu, j < n_inline || chain != 0u); j++) {
                uint2 en = make_uint2(0u, 0u);
                bool valid = false;
                if (j < n_inline) { en = pre.ent[row][j]; valid = true; }
                else if (chain) { en = s_pool[chain - 1u]; chain = s_pool_link[chain - 1u]; valid = true; }
                const int rel = int(en.x & 0xFFu);
                valid = valid && rel >= lo;
                const uint32_t v0 = valid ? en.y : 0u;
                const uint32_t area = valid ? uint32_t(int32_t(en.x) >> 8) : 0u;
                const int first = rel - lane_x;                       // first pixel of the lane that the crossing reaches
                // pixel i gets v0 from cell `first` on and `area` from cell `first + 1` on: i > first <=> i - 1 >= first,
                // so five comparisons serve the eight selects
                const bool pm = first < 0, p0 = first <= 0, p1 = first <= 1, p2 = first <= 2, p3 = first <= 3;
                cov0 += (p0 ? v0 : 0u) + (pm ? area : 0u);
                cov1 += (p1 ? v0 : 0u) + (p0 ? area : 0u);
                cov2 += (p2 ? v0 : 0u) + (p1 ? area : 0u);
                cov3 += (p3 ? v0 : 0u) + (p2 ? area : 0u);
              }
              m[0] = calc_mask(cov0, rule, alpha); m[1] = calc_mask(cov1, rule, alpha);
              m[2] = calc_mask(cov2, rule, alpha); m[3] = calc_mask(cov3, rule, alpha);
            }
          }
          else {
            // Slow path: the group's four warps rasterize one row each, then read their blocks of all four rows.
            const int my_row = grp * kBlockRows + blk;
            // The edges to rasterize: the band's list of the command (k_bin_edges) or, without lists, all its edges.
            const uint32_t ring_entry = s_list[(sub + k) & (kRing - 1)];
            uint2 er = P.cmd_edges[binned ? __ldg(P.cell_cmd + ring_entry) : ring_entry];
            const uint32_t* elist = nullptr;
            if (edge_lists) {
              const uint32_t o0 = __ldg(P.cell_edge_off + ring_entry);
              elist = P.band_edges + o0;
              er.y = __ldg(P.cell_edge_off + ring_entry + 1) - o0;
            }
            slow_group_rows(edges, er, elist, tx0, ty0, my_row, lane, &s_cells[my_row][0], &s_carry[my_row], (256u << 9) + pre.carry_left[my_row], TH, grp);
            const uint4 c = *reinterpret_cast<const uint4*>(&s_cells[row][blk * kBlockW + (lane & 7) * 4]);
            asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory");      // the cells may be overwritten after this
            m[0] = calc_mask(c.x, rule, alpha); m[1] = calc_mask(c.y, rule, alpha);
            m[2] = calc_mask(c.z, rule, alpha); m[3] = calc_mask(c.w, rule, alpha);
          }
          // Pixels outside the command's clipped box never composite (FillData::Analytic::box clamps x1 to the width).
          if (flags & kPreClipRight) {
            const int bx1 = int(h1.y);
            #pragma unroll
            for (int i = 0; i < 4; i++) if (px + i >= bx1) m[i] = 0;
          }
        }
        // The votes keep every branch warp-uniform, so the lanes stay converged for the next iteration.
        if (!__any_sync(0xFFFFFFFFu, (m[0] | m[1] | m[2] | m[3]) != 0u)) continue;
        const uint32_t not_opaque = ((m[0] + 1u) | (m[1] + 1u) | (m[2] + 1u) | (m[3] + 1u)) & 0xFEu;
        opaque = __all_sync(0xFFFFFFFFu, not_opaque == 0u);
        }

        // ---- fetch + composite ----
        FetchEnv env;
        env.fetch_type = h1.x & 0xFFu;
        env.src_format = h1.x >> 16;
        env.solid = h0.w;
        env.fd = reinterpret_cast<const b2dgpu_fetch_data*>(uintptr_t(h1.z) | (uintptr_t(h1.w) << 32));
        env.bayer = P.bayer;
        env.origin_x = P.origin_x; env.origin_y = P.origin_y;
        const uint32_t comp_op = (h1.x >> 8) & 0xFFu;
        uint32_t s[4] = { 0, 0, 0, 0 };
        RowCtx3 rc3;
        rc3.a = pre.rowctx[row][0]; rc3.b = pre.rowctx[row][1]; rc3.c = pre.rowctx[row][2];
        fetch4(env, uint32_t(px), uint32_t(py), m, s, &rc3);
        if (BPP == 1) {
          #pragma unroll
          for (int i = 0; i < 4; i++) s[i] = (s[i] >> 24) * 0x01010101u;
        }
        composite4(comp_op, d, s, m, opaque);
        if (count_pixels) px_written += (m[0] != 0) + (m[1] != 0) + (m[2] != 0) + (m[3] != 0);
      }
      }
      PT(pt_last = pt_now(); pt_busy2 += pt_last - pt_c;)
    }
  }
  PT(if (lane == 0) { atomicAdd(&g_phase_cycles[8], (unsigned long long)pt_round); atomicAdd(&g_phase_cycles[9], (unsigned long long)pt_rounds); atomicAdd(&g_phase_cycles[10], (unsigned long long)pt_cls); atomicAdd(&g_phase_cycles[11], (unsigned long long)pt_chunks); atomicAdd(&g_phase_cycles[12], (unsigned long long)pt_fin); atomicAdd(&g_phase_cycles[13], (unsigned long long)pt_pro); atomicAdd(&g_phase_cycles[14], (unsigned long long)pt_in); })
  PT(if (lane == 0) { atomicAdd(&g_phase_cycles[0], (unsigned long long)pt_busy1); atomicAdd(&g_phase_cycles[1], (unsigned long long)pt_wait1);
                      atomicAdd(&g_phase_cycles[2], (unsigned long long)pt_busy2); atomicAdd(&g_phase_cycles[3], (unsigned long long)pt_wait2);
                      if (warp == 0) { atomicAdd(&g_phase_cycles[4], (unsigned long long)pt_len1); atomicAdd(&g_phase_cycles[5], (unsigned long long)pt_sub);
                                       atomicAdd(&g_phase_cycles[6], (unsigned long long)pt_cmds); atomicAdd(&g_phase_cycles[7], (unsigned long long)pt_maxcmd); } })

  if (dirty) {
    if (BPP == 4) *reinterpret_cast<uint4*>(dst_row + size_t(px) * 4) = make_uint4(d[0], d[1], d[2], d[3]);
    else *reinterpret_cast<uint32_t*>(dst_row + px) = (d[0] >> 24) | ((d[1] >> 24) << 8) | ((d[2] >> 24) << 16) | ((d[3] >> 24) << 24);
  }

  if (count_pixels) {
    px_written = __reduce_add_sync(0xFFFFFFFFu, px_written);
    if (lane == 0 && px_written) atomicAdd(P.pixel_counter, (unsigned long long)px_written);
  }
}

// =================================================================================================================
// K3s - streaming compositor for batches that only hold a few box fills (fill_all / clear_all / large rectangles).
//
// Such batches are pure HBM streaming (8 B per pixel for SrcOver): there is no coverage to accumulate and
// nothing to order across tiles, so instead of one short-lived CTA per tile a persistent grid walks the canvas in
// 16-byte chunks with four independent vector loads in flight per thread before any of them is consumed.
// =================================================================================================================
enum : int { kStreamMaxCmds = 8, kStreamUnroll = 2 };

// What a streaming thread needs of one command, decoded once per CTA.
struct StreamCmd {
  uint32_t type, alpha, comp_op, pad_;
  int box[4];                 // BOX_A: pixels
  BoxUParams bu;              // BOX_U: closed form of the mask-command program
  FetchEnv env;
};

template<int BPP>
__device__ __forceinline__ void stream_chunk(const StreamCmd* cmds, int ncmd, int x, int y, uint32_t* d, uint32_t& written) {
  for (int k = 0; k < ncmd; k++) {
    const StreamCmd& cmd = cmds[k];
    uint32_t m[4];
    bool opaque;
    if (cmd.type == B2DGPU_CMD_FILL_BOX_A) {
      const bool in_y = y >= cmd.box[1] && y < cmd.box[3];
      #pragma unroll
      for (int i = 0; i < 4; i++) m[i] = (in_y && x + i >= cmd.box[0] && x + i < cmd.box[2]) ? cmd.alpha : 0u;
      opaque = cmd.alpha == 255u;
    }
    else {
      #pragma unroll
      for (int i = 0; i < 4; i++) m[i] = box_u_mask(cmd.bu, x + i, y);
      opaque = false;
    }
    if ((m[0] | m[1] | m[2] | m[3]) == 0) continue;
    uint32_t s[4] = { 0, 0, 0, 0 };
    fetch4(cmd.env, uint32_t(x), uint32_t(y), m, s);
    if (BPP == 1) {
      #pragma unroll
      for (int i = 0; i < 4; i++) s[i] = (s[i] >> 24) * 0x01010101u;
    }
    composite4(cmd.comp_op, d, s, m, opaque);
    written += (m[0] != 0) + (m[1] != 0) + (m[2] != 0) + (m[3] != 0);
  }
}

template<int BPP>
__global__ void __launch_bounds__(256) k_box_stream(TileParams P, int rows, int y0r, int x0c, int chunks_per_row, int step_r, int step_c) {
  __shared__ StreamCmd s_cmds[kStreamMaxCmds];
  const int ncmd = int(P.command_count);
  if (threadIdx.x < ncmd) {
    const b2dgpu_command& c = P.commands[threadIdx.x];
    StreamCmd& o = s_cmds[threadIdx.x];
    o.type = c.type; o.alpha = c.alpha; o.comp_op = B2DGPU_SIG_COMP_OP(c.signature); o.pad_ = 0;
    o.box[0] = c.box[0]; o.box[1] = c.box[1]; o.box[2] = c.box[2]; o.box[3] = c.box[3];
    o.bu = box_u_setup(c.box, c.alpha);
    o.env.fetch_type = B2DGPU_SIG_FETCH_TYPE(c.signature);
    o.env.src_format = B2DGPU_SIG_SRC_FORMAT(c.signature);
    o.env.solid = c.solid_prgb32;
    o.env.fd = P.fetch_data + c.fetch_index;
    o.env.bayer = P.bayer;
    o.env.origin_x = P.origin_x; o.env.origin_y = P.origin_y;
  }
  __syncthreads();

  // The dirty region in 4-pixel chunks.  The linear chunk index advances by a fixed stride; (row, column) follow it
  // without a division: stride = step_r rows + step_c columns (computed by the launcher).
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int r = int(tid / chunks_per_row);
  int c = int(tid - (long long)r * chunks_per_row);
  uint32_t written = 0;
  while (r < rows) {
    uint4 v[kStreamUnroll];
    int xs[kStreamUnroll], ys[kStreamUnroll];
    uint8_t* ptr[kStreamUnroll];
    #pragma unroll
    for (int u = 0; u < kStreamUnroll; u++) {
      ptr[u] = nullptr;
      if (r < rows) {
        ys[u] = y0r + r;
        xs[u] = (x0c + c) * 4;
        ptr[u] = P.dst + size_t(ys[u] - P.y_begin) * P.dst_stride + size_t(xs[u]) * BPP;
        if (BPP == 4) v[u] = *reinterpret_cast<const uint4*>(ptr[u]);
        else { uint32_t b = *reinterpret_cast<const uint32_t*>(ptr[u]); v[u] = make_uint4((b & 0xFFu) * 0x01010101u, ((b >> 8) & 0xFFu) * 0x01010101u, ((b >> 16) & 0xFFu) * 0x01010101u, (b >> 24) * 0x01010101u); }
      }
      c += step_c; r += step_r;
      if (c >= chunks_per_row) { c -= chunks_per_row; r++; }
    }
    #pragma unroll
    for (int u = 0; u < kStreamUnroll; u++) {
      if (!ptr[u]) continue;
      uint32_t d[4] = { v[u].x, v[u].y, v[u].z, v[u].w };
      stream_chunk<BPP>(s_cmds, ncmd, xs[u], ys[u], d, written);
      if (BPP == 4) *reinterpret_cast<uint4*>(ptr[u]) = make_uint4(d[0], d[1], d[2], d[3]);
      else *reinterpret_cast<uint32_t*>(ptr[u]) = (d[0] >> 24) | ((d[1] >> 24) << 8) | ((d[2] >> 24) << 16) | ((d[3] >> 24) << 24);
    }
  }
  if (P.pixel_counter) {
    written = __reduce_add_sync(0xFFFFFFFFu, written);
    if ((threadIdx.x & 31) == 0 && written) atomicAdd(P.pixel_counter, (unsigned long long)written);
  }
}

// =================================================================================================================
// K3f - one solid box fill over a large region: FillBoxA + FetchSolid + SrcOver / SrcCopy (compositeCSpan of
// compopgeneric_p.h:104-122 with a constant source and a constant mask).
//
// Everything that does not depend on the destination is folded on the host side of the launch into two constants:
//   SrcOver  d' = sm + div255(d * ia)          sm = div255(s * m), ia = 255 - sm.a
//   SrcCopy  d' = div255(d * (255 - m) + s*m)  (m == 255: d' = s, the destination is not even read)
// which leaves ~11 integer instructions per pixel, few registers (full occupancy) and four 16-byte loads in flight
// per thread: the kernel is bound by HBM (8 B per pixel, 4 B for the store-only case).  An A8 target is processed as
// 32-bit words of four pixels: the per-byte arithmetic of SrcOver / SrcCopy is the same for a channel and for an A8
// pixel (pixelgeneric_p.h:85-200 vs :292-405), and no byte ever overflows into its neighbour.
// =================================================================================================================
enum : int { kSolidUnroll = 4 };

template<int MODE>
__device__ __forceinline__ uint32_t solid_word(uint32_t d, uint32_t k0, uint32_t k1, uint32_t k2) {
  if (MODE == 0) return k0 + pack_div255(mul(unpack(d), k1));                              // k0 = sm, k1 = ia
  return pack_div255(Lanes2{ (d & 0x00FF00FFu) * k2 + k0, lanes_hi(d) * k2 + k1 });        // k0/k1 = s*m lanes, k2 = 255 - m
}

template<int MODE>
__global__ void __launch_bounds__(256) k_stream_solid(SolidStreamParams P, int chunks_per_row, int c0, int step_r, int step_c) {
  // Constants of the fill.
  uint32_t k0, k1, k2 = 0;
  if (MODE == 0) { k0 = pack_div255(mul(unpack(P.src), P.mask)); k1 = (k0 >> 24) ^ 0xFFu; }
  else { Lanes2 sm = mul(unpack(P.src), P.mask); k0 = sm.rb; k1 = sm.ag; k2 = P.mask ^ 0xFFu; }
  if (MODE == 2) k0 = P.src;

  // Chunk = four 32-bit words.  The linear chunk index advances by a fixed stride; (row, column) follow it without
  // a division: stride = step_r rows + step_c columns (computed by the launcher).
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int r = int(tid / chunks_per_row);
  int c = int(tid - (long long)r * chunks_per_row);
  const int rows = P.y1 - P.y0;

  while (r < rows) {
    uint4 v[kSolidUnroll];
    uint4* ptr[kSolidUnroll];
    int cw[kSolidUnroll];
    #pragma unroll
    for (int u = 0; u < kSolidUnroll; u++) {
      ptr[u] = nullptr;
      if (r < rows) {
        cw[u] = (c0 + c) * 4;
        ptr[u] = reinterpret_cast<uint4*>(P.dst + size_t(P.y0 + r) * P.dst_stride) + (c0 + c);
        if (MODE != 2 || cw[u] < P.x0w || cw[u] + 4 > P.x1w) v[u] = *ptr[u];
      }
      c += step_c; r += step_r;
      if (c >= chunks_per_row) { c -= chunks_per_row; r++; }
    }
    #pragma unroll
    for (int u = 0; u < kSolidUnroll; u++) {
      if (!ptr[u]) continue;
      uint4 o;
      if (cw[u] >= P.x0w && cw[u] + 4 <= P.x1w) {
        if (MODE == 2) o = make_uint4(k0, k0, k0, k0);
        else o = make_uint4(solid_word<MODE>(v[u].x, k0, k1, k2), solid_word<MODE>(v[u].y, k0, k1, k2),
                            solid_word<MODE>(v[u].z, k0, k1, k2), solid_word<MODE>(v[u].w, k0, k1, k2));
      }
      else {
        // chunk cut by the left / right end of the box
        o = v[u];
        if (cw[u] + 0 >= P.x0w && cw[u] + 0 < P.x1w) o.x = MODE == 2 ? k0 : solid_word<MODE>(o.x, k0, k1, k2);
        if (cw[u] + 1 >= P.x0w && cw[u] + 1 < P.x1w) o.y = MODE == 2 ? k0 : solid_word<MODE>(o.y, k0, k1, k2);
        if (cw[u] + 2 >= P.x0w && cw[u] + 2 < P.x1w) o.z = MODE == 2 ? k0 : solid_word<MODE>(o.z, k0, k1, k2);
        if (cw[u] + 3 >= P.x0w && cw[u] + 3 < P.x1w) o.w = MODE == 2 ? k0 : solid_word<MODE>(o.w, k0, k1, k2);
      }
      *ptr[u] = o;
    }
  }
  if (tid == 0 && P.pixel_counter) atomicAdd(P.pixel_counter, P.pixels);
}

// =================================================================================================================
// Launchers (host)
// =================================================================================================================
static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

int launch_build_luts(const LutParams& P, cudaStream_t s) {
  if (!P.request_count) return 0;
  k_build_luts<<<div_up(P.request_count * 32, 128), 128, 0, s>>>(P);
  return 1;
}

int launch_glyph_instances(const GlyphParams& P, cudaStream_t s) {
  if (!P.instance_count) return 0;
  k_glyph_instances<<<div_up(P.instance_count, 128), 128, 0, s>>>(P);
  return 1;
}

int launch_count_edges(const BuildParams& P, cudaStream_t s) {
  if (!P.segment_count) return 0;
  k_build_edges<false><<<div_up(P.segment_count, 128), 128, 0, s>>>(P);
  return 1;
}

int launch_write_edges(const BuildParams& P, cudaStream_t s) {
  if (!P.segment_count) return 0;
  k_build_edges<true><<<div_up(P.segment_count, 128), 128, 0, s>>>(P);
  return 1;
}

// Exclusive scan of in[0..n) into out[0..n], out[n] = total, *total_out = total.  `scratch` needs
// scan_scratch_items(n) u32.  Returns the number of kernels launched.
size_t scan_scratch_items(uint32_t n) {
  size_t total = 0;
  uint32_t m = n;
  while (m > kScanBlock) {
    m = div_up(m, kScanBlock);
    total += size_t(m) + 1;
  }
  return total + 8;
}

int launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch, uint32_t* total_out, cudaStream_t s, const uint32_t* n_dev) {
  if (n <= kScanBlock) {
    k_scan_small<<<1, kScanThreads, 0, s>>>(in, out, n, total_out, n_dev);
    return 1;
  }
  uint32_t nb = div_up(n, kScanBlock);
  uint32_t* block_sums = scratch;               // nb + 1 entries; scanned in place
  k_scan_blocks<<<nb, kScanThreads, 0, s>>>(in, out, block_sums, n, n_dev);
  int launches = 1;
  // Scan the block sums in place (out == in is fine: every element is read before it is written by its own thread,
  // and blocks only touch their own range).  Blocks beyond a device-side count contribute zero.
  launches += launch_exclusive_scan(block_sums, block_sums, nb, scratch + nb + 1, nullptr, s, nullptr);
  k_scan_add<<<nb, kScanThreads, 0, s>>>(out, block_sums, n, total_out, n_dev);
  return launches + 1;
}

int launch_init_bbox(int4* bbox, uint32_t n, cudaStream_t s) {
  if (!n) return 0;
  k_init_bbox<<<div_up(n, 256), 256, 0, s>>>(bbox, n);
  return 1;
}

int launch_analytic_bbox(const b2dgpu_command* cmds, uint32_t ncmd, const b2dgpu_edge* edges, int4* bbox, cudaStream_t s) {
  if (!ncmd) return 0;
  k_analytic_bbox<<<div_up(ncmd * 32, 256), 256, 0, s>>>(cmds, ncmd, edges, bbox);
  return 1;
}

int launch_finalize_commands(const FinalizeParams& P, cudaStream_t s) {
  if (!P.command_count) return 0;
  k_finalize_commands<<<div_up(P.command_count, 256), 256, 0, s>>>(P);
  return 1;
}

// Streaming path: `box` = union of the commands' pixel boxes (already clipped to the target), in pixels.
int launch_box_stream(const TileParams& P, int bpp, const int* box, int sm_count, cudaStream_t s) {
  int x0c = box[0] / 4, x1c = (box[2] + 3) / 4;
  int rows = box[3] - box[1];
  int chunks_per_row = x1c - x0c;
  if (rows <= 0 || chunks_per_row <= 0) return 0;
  static int per_sm[2] = { 0, 0 };
  const int which = bpp == 4 ? 0 : 1;
  if (!per_sm[which]) {
    int n = 0;
    const void* fn = bpp == 4 ? (const void*)k_box_stream<4> : (const void*)k_box_stream<1>;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 256, 0) != cudaSuccess || n < 1) n = 2;
    per_sm[which] = n;
  }
  long long total = (long long)rows * chunks_per_row;
  long long want = (total + 256LL * kStreamUnroll - 1) / (256LL * kStreamUnroll);
  const long long cap = (long long)sm_count * per_sm[which];
  int grid = int(want < cap ? (want < 1 ? 1 : want) : cap);
  const long long stride = (long long)grid * 256;
  const int step_r = int(stride / chunks_per_row), step_c = int(stride - (long long)step_r * chunks_per_row);
  if (bpp == 4) k_box_stream<4><<<grid, 256, 0, s>>>(P, rows, box[1], x0c, chunks_per_row, step_r, step_c);
  else k_box_stream<1><<<grid, 256, 0, s>>>(P, rows, box[1], x0c, chunks_per_row, step_r, step_c);
  return 1;
}

int launch_stream_solid(const SolidStreamParams& P, int sm_count, cudaStream_t s) {
  const int c0 = P.x0w / 4, c1 = (P.x1w + 3) / 4;
  const int chunks_per_row = c1 - c0, rows = P.y1 - P.y0;
  if (chunks_per_row <= 0 || rows <= 0) return 0;
  const long long total = (long long)rows * chunks_per_row;
  // Persistent grid: exactly the CTAs that are resident at once (one wave, no tail).
  static int per_sm[3] = { 0, 0, 0 };
  if (!per_sm[P.mode]) {
    int n = 0;
    const void* fn = P.mode == 0 ? (const void*)k_stream_solid<0> : P.mode == 1 ? (const void*)k_stream_solid<1> : (const void*)k_stream_solid<2>;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 256, 0) != cudaSuccess || n < 1) n = 4;
    per_sm[P.mode] = n;
  }
  long long want = (total + 256LL * kSolidUnroll - 1) / (256LL * kSolidUnroll);
  const long long cap = (long long)sm_count * per_sm[P.mode];
  const int grid = int(want < cap ? (want < 1 ? 1 : want) : cap);
  const long long stride = (long long)grid * 256;
  const int step_r = int(stride / chunks_per_row), step_c = int(stride - (long long)step_r * chunks_per_row);
  if (P.mode == 0) k_stream_solid<0><<<grid, 256, 0, s>>>(P, chunks_per_row, c0, step_r, step_c);
  else if (P.mode == 1) k_stream_solid<1><<<grid, 256, 0, s>>>(P, chunks_per_row, c0, step_r, step_c);
  else k_stream_solid<2><<<grid, 256, 0, s>>>(P, chunks_per_row, c0, step_r, step_c);
  return 1;
}

size_t bin_scratch_items(uint32_t command_count, int tiles_y) {
  return scan_scratch_items(command_count) + scan_scratch_items(uint32_t(tiles_y) + 1u) + 16;
}

int launch_binning(const BinParams& B, cudaStream_t s) {
  if (!B.command_count || B.tiles_y <= 0) return 0;
  int launches = 0;
  // band_count doubles as the difference array: tiles_y + 2 entries (k_bin_count writes index band1 + 1 <= tiles_y)
  cudaMemsetAsync(B.band_count, 0, sizeof(uint32_t) * (size_t(B.tiles_y) + 2), s);
  k_bin_count<<<div_up(B.command_count, 256), 256, 0, s>>>(B);
  launches += 1;
  launches += launch_exclusive_scan(B.cm_count, B.cm_base, B.command_count, B.scan_scratch, nullptr, s);
  // exclusive scan E of the differences: commands per band b = E[b + 1]; band_off = exclusive scan of those
  uint32_t* scratch2 = B.scan_scratch + scan_scratch_items(B.command_count);
  launches += launch_exclusive_scan(B.band_count, B.band_prefix, uint32_t(B.tiles_y) + 1u, scratch2, nullptr, s);
  launches += launch_exclusive_scan(B.band_prefix + 1, B.band_off, uint32_t(B.tiles_y), scratch2, nullptr, s);
  k_bin_check<<<1, 1, 0, s>>>(B);
  k_bin_fill<<<B.tiles_y, 1024, 0, s>>>(B);
  k_bin_extents<<<div_up(B.command_count * 32, 256), 256, 0, s>>>(B);
  launches += 3;
  if (B.band_edges) {
    // per-cell edge lists: scan of the pair counts over the cells (their number lives in state[0]), then the fill pass
    launches += launch_exclusive_scan(B.cell_edge_cnt, B.cell_edge_off, B.capacity, B.scan_scratch2, nullptr, s, B.state);
    k_bin_check_edges<<<1, 1, 0, s>>>(B);
    k_bin_edges<<<div_up(B.command_count * 32, 256), 256, 0, s>>>(B);
    launches += 2;
  }
  return launches;
}

size_t bin_cell_scan_scratch_items(uint32_t capacity) { return scan_scratch_items(capacity); }

// Tile height for a target of `rows` rows and `tiles_x` tile columns: 32-row tiles (one 32-warp CTA per SM, the
// configuration the 4K bench runs) when they give every SM a couple of tiles, shorter tiles for small canvases.
int choose_tile_height(int tiles_x, int rows, int sm_count) {
  const int heights[3] = { 32, 16, 8 };
  if (const char* e = getenv("B2D_TILE_H")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) return v; }   // experiment knob
  for (int i = 0; i < 3; i++)
    if (tiles_x * ((rows + heights[i] - 1) / heights[i]) >= 2 * sm_count) return heights[i];
  return 8;
}

template<int BPP, int TH>
static void launch_tile_variant(const TileParams& P, uint32_t tiles, cudaStream_t s) {
  // Function attributes are per device: remember which devices of this process were configured.
  static bool configured[64] = {};
  const int dyn = int(sizeof(PreCmdT<TH>)) * TileCfg<TH>::kSub + TileCfg<TH>::kPool * int(sizeof(uint2) + sizeof(uint16_t));
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(k_tile_render<BPP, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  k_tile_render<BPP, TH><<<tiles, TileCfg<TH>::kThreads, dyn, s>>>(P);
}

int launch_tile_render(const TileParams& P, int bpp, int tile_h, cudaStream_t s) {
  if (!P.command_count) return 0;
  uint32_t tiles = uint32_t(P.tiles_x) * uint32_t(P.tiles_y);
  if (!tiles) return 0;
  if (bpp == 4) {
    if (tile_h == 32) launch_tile_variant<4, 32>(P, tiles, s);
    else if (tile_h == 16) launch_tile_variant<4, 16>(P, tiles, s);
    else launch_tile_variant<4, 8>(P, tiles, s);
  }
  else {
    if (tile_h == 32) launch_tile_variant<1, 32>(P, tiles, s);
    else if (tile_h == 16) launch_tile_variant<1, 16>(P, tiles, s);
    else launch_tile_variant<1, 8>(P, tiles, s);
  }
  return 1;
}

} // namespace b2d
