// hostsim.cpp - TEST-ONLY lockstep simulator of the GPU compositor's scalar device functions.
//
// This translation unit includes the very same dev_*.cuh headers that nvcc compiles into the sm_100a kernels and runs
// them sequentially on the CPU: per path segment -> build_line/quad/cubic, per tile and command -> tile_accumulate_edge,
// prefix sum, calc_mask, fetch_pixel, composite.  It exists so that the parts of the algorithm that do not depend on
// CUDA's execution model can be checked against the reference on a machine without a GPU (`pytest -m "not gpu"`).
// It is built from tests/ into tests/hostsim/_hostsim.so and is NOT part of libb2dgpu.so nor reachable from the
// blend2d_b200 package: the product has no CPU path.
#include "../../blend2d_b200/csrc/dev_pixel.cuh"
#include "../../blend2d_b200/csrc/dev_raster.cuh"
#include "../../blend2d_b200/csrc/dev_flatten.cuh"
#include "../../blend2d_b200/csrc/dev_fetch.cuh"
#include "../../blend2d_b200/csrc/dev_tile.cuh"

#include <limits.h>
#include <string.h>
#include <algorithm>
#include <utility>
#include <vector>

using namespace b2d;

namespace {

struct VecOut {
  std::vector<b2dgpu_edge>* v;
  void edge(int x0, int y0, int x1, int y1) { b2dgpu_edge e; e.x0 = x0; e.y0 = y0; e.x1 = x1; e.y1 = y1; v->push_back(e); }
};

void build_segment(const b2dgpu_batch_view* B, uint32_t si, VecOut& out) {
  const b2dgpu_segment seg = B->segments[si];
  const b2dgpu_command& cmd = B->commands[seg.command];
  const b2dgpu_geometry_state& gs = B->geometry_states[cmd.state_index];
  GeomXform xf;
  xf.m00 = gs.m[0]; xf.m01 = gs.m[1]; xf.m10 = gs.m[2]; xf.m11 = gs.m[3]; xf.m20 = gs.m[4]; xf.m21 = gs.m[5];
  xf.affine = gs.transform_type > 2u;
  ClipBox cb;
  cb.x0 = gs.clip[0]; cb.y0 = gs.clip[1]; cb.x1 = gs.clip[2]; cb.y1 = gs.clip[3];
  cb.ix0 = trunc_i(cb.x0); cb.ix1 = trunc_i(cb.x1);
  const double* v = B->vertices;
  uint32_t kind = seg.p1_kind & 3u, i1 = seg.p1_kind >> 2;
  P2 p0 = xform(xf, mk(v[seg.p0 * 2], v[seg.p0 * 2 + 1]));
  P2 p1 = xform(xf, mk(v[i1 * 2], v[i1 * 2 + 1]));
  if (kind == B2DGPU_SEG_LINE) build_line(p0, p1, cb, out);
  else if (kind == B2DGPU_SEG_CUBIC)
    build_cubic(p0, p1, xform(xf, mk(v[(i1 + 1) * 2], v[(i1 + 1) * 2 + 1])), xform(xf, mk(v[(i1 + 2) * 2], v[(i1 + 2) * 2 + 1])), cb, gs.tolerance_sq, out);
  else {
    uint32_t i2 = i1 + (kind == B2DGPU_SEG_CONIC ? 2u : 1u);
    build_quad(p0, p1, xform(xf, mk(v[i2 * 2], v[i2 * 2 + 1])), cb, gs.tolerance_sq, out);
  }
}

struct HostStore {
  uint32_t cells[kTileH][kTileW];
  uint32_t carry[kTileH];
  void add_cell(int row, int rel, uint32_t v) { cells[row][rel] += v; }
  void add_carry(int row, uint32_t v) { carry[row] += v; }
};

} // namespace

extern "C" {

// Flattens all FILL_GEOMETRY commands.  edges_out may be null to only count.  begins_out: command_count + 1 entries.
__attribute__((visibility("default")))
uint32_t hostsim_build_edges(const b2dgpu_batch_view* B, b2dgpu_edge* edges_out, uint32_t capacity, uint32_t* begins_out) {
  std::vector<b2dgpu_edge> all;
  VecOut out{ &all };
  for (uint32_t c = 0; c < B->command_count; c++) {
    const b2dgpu_command& cmd = B->commands[c];
    if (begins_out) begins_out[c] = uint32_t(all.size());
    if (cmd.type != B2DGPU_CMD_FILL_GEOMETRY) continue;
    for (uint32_t s = 0; s < cmd.data_count; s++) build_segment(B, cmd.data_offset + s, out);
  }
  if (begins_out) begins_out[B->command_count] = uint32_t(all.size());
  if (edges_out) memcpy(edges_out, all.data(), sizeof(b2dgpu_edge) * (all.size() < capacity ? all.size() : capacity));
  return uint32_t(all.size());
}

// Renders the batch into a host image, tile by tile, exactly in the order of operations of k_tile_render.
// `bayer` = 512-byte table (16 rows x 32).  Returns the number of composited pixels.
__attribute__((visibility("default")))
uint64_t hostsim_render(const b2dgpu_batch_view* B, uint8_t* pixels, intptr_t stride, int w, int h, uint32_t format, const uint8_t* bayer) {
  const int bpp = format == B2DGPU_FORMAT_A8 ? 1 : 4;
  uint64_t written = 0;

  // K1 for every geometry command.
  std::vector<b2dgpu_edge> edges(B->edges, B->edges + B->edge_count);
  std::vector<uint32_t> e_begin(B->command_count), e_count(B->command_count);
  std::vector<CmdBox> boxes(B->command_count);
  VecOut out{ &edges };
  for (uint32_t c = 0; c < B->command_count; c++) {
    const b2dgpu_command& cmd = B->commands[c];
    e_begin[c] = 0; e_count[c] = 0;
    if (cmd.type == B2DGPU_CMD_FILL_ANALYTIC) { e_begin[c] = cmd.data_offset; e_count[c] = cmd.data_count; }
    else if (cmd.type == B2DGPU_CMD_FILL_GEOMETRY) {
      e_begin[c] = uint32_t(edges.size());
      for (uint32_t s = 0; s < cmd.data_count; s++) build_segment(B, cmd.data_offset + s, out);
      e_count[c] = uint32_t(edges.size()) - e_begin[c];
    }
    int fx0 = INT_MAX, fy0 = INT_MAX, fx1 = INT_MIN, fy1 = INT_MIN;
    for (uint32_t i = 0; i < e_count[c]; i++) {
      const b2dgpu_edge& e = edges[e_begin[c] + i];
      fx0 = tmin(fx0, tmin(e.x0, e.x1)); fx1 = tmax(fx1, tmax(e.x0, e.x1));
      fy0 = tmin(fy0, tmin(e.y0, e.y1)); fy1 = tmax(fy1, tmax(e.y0, e.y1));
    }
    boxes[c] = command_pixel_box(cmd, e_count[c], fx0, fy0, fx1, fy1, w, 0, h);
  }

  const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;

  // K1d (k_band_extents): columns each command's edges can touch per band of kTileH rows.
  std::vector<int> ext_lo(size_t(tiles_y) * B->command_count, INT_MAX), ext_hi(size_t(tiles_y) * B->command_count, -1);
  for (uint32_t c = 0; c < B->command_count; c++) {
    const CmdBox& bb = boxes[c];
    if (bb.x0 >= bb.x1 || bb.y0 >= bb.y1) continue;
    if (!command_has_edges(B->commands[c].type)) {
      for (int b = bb.y0 / kTileH; b <= (bb.y1 - 1) / kTileH; b++) { ext_lo[size_t(b) * B->command_count + c] = 0; ext_hi[size_t(b) * B->command_count + c] = INT_MAX; }
      continue;
    }
    for (uint32_t e = 0; e < e_count[c]; e++) {
      NormEdge ne = normalize_edge(edges[e_begin[c] + e]);
      if (ne.y0 == ne.y1) continue;
      const int row_first = tmax(ne.y0 >> 8, bb.y0), row_last = tmin((ne.y1 - 1) >> 8, bb.y1 - 1);
      for (int b = row_first / kTileH; row_first <= row_last && b <= row_last / kTileH; b++) {
        int lo, hi;
        band_edge_extent(ne, b * kTileH, lo, hi);
        int& l = ext_lo[size_t(b) * B->command_count + c]; l = tmin(l, lo);
        int& hh = ext_hi[size_t(b) * B->command_count + c]; hh = tmax(hh, hi);
      }
    }
  }

  static HostStore store;
  for (int ty = 0; ty < tiles_y; ty++) for (int tx = 0; tx < tiles_x; tx++) {
    const int tx0 = tx * kTileW, ty0 = ty * kTileH;
    memset(&store, 0, sizeof(store));
    for (uint32_t c = 0; c < B->command_count; c++) {
      const CmdBox& bb = boxes[c];
      if (!(bb.x0 < tx0 + kTileW && bb.x1 > tx0 && bb.y0 < ty0 + kTileH && bb.y1 > ty0)) continue;
      if (!(tx0 + kTileW > ext_lo[size_t(ty) * B->command_count + c] && tx0 <= ext_hi[size_t(ty) * B->command_count + c])) continue;
      const b2dgpu_command& cmd = B->commands[c];
      const uint32_t sig = cmd.signature, alpha = cmd.alpha;
      uint32_t masks[kTileH][kTileW];
      memset(masks, 0, sizeof(masks));

      if (cmd.type == B2DGPU_CMD_FILL_BOX_A) {
        for (int r = 0; r < kTileH; r++) for (int x = 0; x < kTileW; x++) {
          int px = tx0 + x, py = ty0 + r;
          masks[r][x] = (py >= cmd.box[1] && py < cmd.box[3] && px >= cmd.box[0] && px < cmd.box[2]) ? alpha : 0u;
        }
      }
      else if (cmd.type == B2DGPU_CMD_FILL_BOX_MASK_A) {
        const b2dgpu_pattern_source& ms = B->fetch_data[cmd.reserved[0]].pattern.src;
        for (int r = 0; r < kTileH; r++) for (int x = 0; x < kTileW; x++) masks[r][x] = box_mask_a(cmd, ms, tx0 + x, ty0 + r);
      }
      else if (cmd.type == B2DGPU_CMD_FILL_BOX_U) {
        BoxUParams bu = box_u_setup(cmd.box, alpha);
        for (int r = 0; r < kTileH; r++) for (int x = 0; x < kTileW; x++) masks[r][x] = box_u_mask(bu, tx0 + x, ty0 + r);
      }
      else {
        // Phase 1 (classification): backdrop from the edges entirely left of the tile, and whether any edge straddles.
        uint32_t left_acc[kTileH];
        memset(left_acc, 0, sizeof(left_acc));
        bool touched = false;
        uint32_t straddlers = 0;
        for (uint32_t e = 0; e < e_count[c]; e++) {
          NormEdge ne = normalize_edge(edges[e_begin[c] + e]);
          int cls = tile_edge_class(ne, tx0, ty0);
          if (cls == kEdgeLeft) tile_left_cover(ne, ty0, left_acc);
          else if (cls == kEdgeStraddle) straddlers++;
        }
        for (int r = 0; r < kTileH; r++) if (left_acc[r]) { store.carry[r] += left_acc[r]; touched = true; }
        // Phase 2: only straddling edges go through the rasterizer.
        if (straddlers) {
          for (uint32_t e = 0; e < e_count[c]; e++) {
            NormEdge ne = normalize_edge(edges[e_begin[c] + e]);
            if (tile_edge_class(ne, tx0, ty0) != kEdgeStraddle) continue;
            // One (edge, row) item at a time, exactly like a GPU lane: prepare + advance_to_y + one step.
            TileSink<HostStore> sink(store, tx0);
            const int y_from = tmax(ne.y0 >> 8, ty0), y_to = tmin((ne.y1 - 1) >> 8, ty0 + kTileH - 1);
            for (int y = y_from; y <= y_to; y++) { sink.row = y - ty0; tile_rasterize_edge_row(ne, y, sink); }
            touched |= sink.touched != 0;
          }
        }
        if (!touched) continue;
        for (int r = 0; r < kTileH; r++) {
          uint32_t cov = (256u << 9) + store.carry[r];
          for (int x = 0; x < kTileW; x++) {
            cov += store.cells[r][x];
            uint32_t m = calc_mask(cov, cmd.fill_rule_mask, alpha);
            if (tx0 + x >= bb.x1) m = 0;
            masks[r][x] = m;
          }
        }
        memset(&store, 0, sizeof(store));
      }

      FetchEnv env;
      env.fetch_type = B2DGPU_SIG_FETCH_TYPE(sig);
      env.src_format = B2DGPU_SIG_SRC_FORMAT(sig);
      env.solid = cmd.solid_prgb32;
      env.fd = B->fetch_data ? B->fetch_data + cmd.fetch_index : nullptr;
      env.bayer = bayer;
      env.origin_x = B->pixel_origin_x; env.origin_y = B->pixel_origin_y;
      const uint32_t comp_op = B2DGPU_SIG_COMP_OP(sig);

      for (int r = 0; r < kTileH; r++) {
        int py = ty0 + r;
        if (py >= h) break;
        RowCtx rc;
        fetch_row_init(env, uint32_t(py), rc);
        for (int x = 0; x < kTileW; x++) {
          int px = tx0 + x;
          if (px >= w) break;
          uint32_t m = masks[r][x];
          if (!m) continue;
          uint8_t* p = pixels + intptr_t(py) * stride + intptr_t(px) * bpp;
          uint32_t d = bpp == 4 ? *reinterpret_cast<uint32_t*>(p) : uint32_t(*p) * 0x01010101u;
          uint32_t s = fetch_pixel(env, rc, uint32_t(px), uint32_t(py));
          if (bpp == 1) s = (s >> 24) * 0x01010101u;
          d = composite(comp_op, d, s, m);
          if (bpp == 4) *reinterpret_cast<uint32_t*>(p) = d; else *p = uint8_t(d >> 24);
          written++;
        }
      }
    }
  }
  return written;
}

// Cells of `n` lines (x0, y0, x1, y1 in 24.8, any direction) accumulated with the device's rasterizer (dev_raster.cuh):
// cells[y * (w + 2) + x].  mode 0: prepare once, one edge_step_scanline per row (the reference's walk); mode 1: every
// (edge, row) on its own - prepare, edge_advance_to_y, one step - which is what a GPU lane does
// (tile_rasterize_edge_row).  KAT against the reference's AnalyticRasterizer: tests/test_rasterizer_kat.py.
namespace {
struct CellImageSink {
  uint32_t* cells; int w, h, y;
  void merge(int x, uint32_t cover, uint32_t area) {
    if (y < 0 || y >= h || x < 0 || x > w) return;
    cells[size_t(y) * size_t(w + 2) + size_t(x)] += (cover << 9) - area;
    cells[size_t(y) * size_t(w + 2) + size_t(x) + 1] += area;
  }
};
}
__attribute__((visibility("default")))
void hostsim_rasterize_cells(const int32_t* lines, size_t n, int w, int h, int mode, uint32_t* cells) {
  for (size_t i = 0; i < n; i++) {
    b2dgpu_edge raw; raw.x0 = lines[i * 4]; raw.y0 = lines[i * 4 + 1]; raw.x1 = lines[i * 4 + 2]; raw.y1 = lines[i * 4 + 3];
    const b2d::NormEdge ne = b2d::normalize_edge(raw);
    if (ne.y0 == ne.y1) continue;
    const int y_first = ne.y0 >> 8, y_last = (ne.y1 - 1) >> 8;
    CellImageSink sink{ cells, w, h, 0 };
    if (mode == 0) {
      b2d::EdgeState st;
      if (!b2d::edge_prepare(st, ne.x0, ne.y0, ne.x1, ne.y1, ne.sign_bit)) continue;
      for (int y = y_first; y <= y_last; y++) { sink.y = y; if (b2d::edge_step_scanline(st, sink)) break; }
    }
    else {
      for (int y = y_first; y <= y_last; y++) {
        b2d::EdgeState st;
        if (!b2d::edge_prepare(st, ne.x0, ne.y0, ne.x1, ne.y1, ne.sign_bit)) break;
        b2d::edge_advance_to_y(st, y);
        sink.y = y;
        b2d::edge_step_scanline(st, sink);
      }
    }
  }
}

// err_multi_step(count) against `count` single err_step_u() calls on random DDA states (0 <= err < correction,
// 0 <= step < correction): the closed form the windowed one-row stepper (edge_step_scanline<true>) and edge_advance_to_y
// rest on.  Returns the number of mismatches.
__attribute__((visibility("default")))
uint32_t hostsim_err_multi_step_check(uint64_t seed, uint32_t cases) {
  uint64_t s = seed * 6364136223846793005ull + 1442695040888963407ull;
  auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return uint32_t(s >> 16); };
  uint32_t bad = 0;
  for (uint32_t i = 0; i < cases; i++) {
    const int corr = int(next() % ((i & 3) ? 1000000u : 16u)) + 1;      // dx (or dy): small and large
    const int step = int(next() % uint32_t(corr));
    const int err0 = int(next() % uint32_t(corr));
    const int count = int(next() % ((i & 1) ? 4096u : 40u));
    int acc_m = 0, err_m = err0;
    b2d::err_multi_step(acc_m, err_m, step, corr, count);
    uint32_t acc_s = 0; int err_s = err0;
    for (int k = 0; k < count; k++) b2d::err_step_u(acc_s, err_s, step, corr);
    if (uint32_t(acc_m) != acc_s || err_m != err_s) bad++;
  }
  return bad;
}

// dst[i] = composite(op, dst[i], src[i], mask[i]) with the device's operator code (dev_pixel.cuh) - swept against the C
// oracle's replay of the JIT sequences by tests/test_oracle.py.
__attribute__((visibility("default")))
void hostsim_composite_plane(uint32_t op, uint32_t* dst, const uint32_t* src, const uint8_t* mask, size_t n) {
  for (size_t i = 0; i < n; i++) if (mask[i]) dst[i] = b2d::composite(op, dst[i], src[i], mask[i]);
}

} // extern "C"

// Equivalence check for the device edge builder's node expansion (dev_flatten.cuh piece_root / node_split / node_walk):
// `count` random quads (kind 3) or cubics (kind 4) with a random clip box are flattened (a) by the sequential routine
// build_quad / build_cubic - the reference's walk - and (b) piece by piece through the breadth-first expansion into at
// most kNodeCap nodes followed by independent walks, exactly as k_build_edges does.  Lines must agree as multisets;
// vertical lines (borders are un-merged in (b)) must agree as signed coverage along y per x.  Returns the number of
// curves that differ.
template<int N>
static void expanded_build(const P2* pts, const ClipBox& cb, double tol_sq, VecOut& out) {
  P2 spline[25];
  uint32_t any = 0;
  const int pieces = N == 3 ? prepare_quad(pts[0], pts[1], pts[2], cb, out, spline, any)
                            : prepare_cubic(pts[0], pts[1], pts[2], pts[3], cb, out, spline, any);
  struct Node { P2 p[4]; uint32_t meta; };
  std::vector<Node> cur, nxt;
  MonoCurve<N> mc;
  mc.tol_sq = tol_sq;
  for (int i = 0; i < pieces; i++) {
    uint32_t meta = 0;
    if (!piece_root<N>(mc, spline + i * (N - 1), any != 0, cb, out, meta)) continue;
    Node n; for (int k = 0; k < N; k++) n.p[k] = mc.p[k]; n.meta = meta;
    cur.push_back(n);
  }
  while (!cur.empty() && cur.size() * 2 <= size_t(kNodeCap)) {
    bool split_any = false;
    nxt.clear();
    for (const Node& n : cur) {
      mc.begin_at(n.p, 0);
      Node a = n, b = n;
      if (node_split<N>(mc, n.meta & kNodePendingMask, a.p, b.p)) { a.meta = n.meta + 1u; nxt.push_back(a); nxt.push_back(b); split_any = true; }
      else nxt.push_back(n);
    }
    cur.swap(nxt);
    if (!split_any) break;
  }
  for (const Node& n : cur) node_walk<N>(mc, n.p, n.meta, cb, out);
}

static void canonical(const std::vector<b2dgpu_edge>& in, std::vector<b2dgpu_edge>& lines, std::vector<b2dgpu_edge>& vert) {
  lines.clear(); vert.clear();
  std::vector<std::pair<std::pair<int, int>, int>> ev;        // ((x, y), +-1)
  for (const b2dgpu_edge& e : in) {
    if (e.x0 != e.x1) { lines.push_back(e); continue; }
    const int sgn = e.y0 < e.y1 ? 1 : -1;
    ev.push_back({ { e.x0, e.y0 < e.y1 ? e.y0 : e.y1 }, sgn });
    ev.push_back({ { e.x0, e.y0 < e.y1 ? e.y1 : e.y0 }, -sgn });
  }
  auto less = [](const b2dgpu_edge& a, const b2dgpu_edge& b) { return memcmp(&a, &b, sizeof(a)) < 0; };
  std::sort(lines.begin(), lines.end(), less);
  std::sort(ev.begin(), ev.end());
  // running winding along y per x -> maximal runs of constant non-zero winding
  size_t i = 0;
  while (i < ev.size()) {
    const int x = ev[i].first.first;
    int w = 0, run_start = 0;
    while (i < ev.size() && ev[i].first.first == x) {
      const int y = ev[i].first.second;
      int d = 0;
      while (i < ev.size() && ev[i].first.first == x && ev[i].first.second == y) d += ev[i++].second;
      if (d == 0) continue;
      if (w != 0) { b2dgpu_edge e; e.x0 = x; e.x1 = w; e.y0 = run_start; e.y1 = y; vert.push_back(e); }
      w += d; run_start = y;
    }
  }
}

extern "C" {

__attribute__((visibility("default")))
uint32_t hostsim_flatten_equivalence(uint32_t kind, uint32_t count, uint64_t seed, uint32_t mode) {
  uint64_t st = seed * 6364136223846793005ull + 1442695040888963407ull;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return double(st >> 11) / 9007199254740992.0; };
  uint32_t bad = 0;
  std::vector<b2dgpu_edge> va, vb, la, lb, ca, cb2;
  for (uint32_t it = 0; it < count; it++) {
    // clip box in 24.8 units, like final_clip_box_fixed_d of a W x H canvas
    const double W = double(int(8 + rnd() * 4000)) * 256.0, H = double(int(8 + rnd() * 2500)) * 256.0;
    ClipBox cb; cb.x0 = 0.0; cb.y0 = 0.0; cb.x1 = W; cb.y1 = H; cb.ix0 = trunc_i(cb.x0); cb.ix1 = trunc_i(cb.x1);
    // mode 0: points a little outside the box (tester paths); 1: far outside; 2: all inside; 3: tiny curves (glyphs)
    const double m = mode == 0 ? 0.05 : mode == 1 ? 3.0 : 0.0;
    P2 pts[4];
    if (mode == 3) {
      const double ox = rnd() * W, oy = rnd() * H, sz = 256.0 * (1.0 + rnd() * 40.0);
      for (int k = 0; k < 4; k++) pts[k] = mk(ox + rnd() * sz, oy + rnd() * sz);
    }
    else for (int k = 0; k < 4; k++) pts[k] = mk((rnd() * (1.0 + 2.0 * m) - m) * W, (rnd() * (1.0 + 2.0 * m) - m) * H);
    if (it % 7 == 3) pts[1] = pts[0];                              // degenerate control points
    if (it % 11 == 5) pts[2].x = pts[1].x;
    if (it % 13 == 6) { pts[0].x = 0.0; pts[kind - 1].x = W; }     // end points exactly on the clip edges
    const double tol = 0.2 * 256.0;
    va.clear(); vb.clear();
    VecOut oa{ &va }, ob{ &vb };
    if (kind == 3) { build_quad(pts[0], pts[1], pts[2], cb, tol * tol, oa); expanded_build<3>(pts, cb, tol * tol, ob); }
    else { build_cubic(pts[0], pts[1], pts[2], pts[3], cb, tol * tol, oa); expanded_build<4>(pts, cb, tol * tol, ob); }
    canonical(va, la, ca); canonical(vb, lb, cb2);
    const bool same = la.size() == lb.size() && ca.size() == cb2.size() &&
                      (la.empty() || memcmp(la.data(), lb.data(), la.size() * sizeof(b2dgpu_edge)) == 0) &&
                      (ca.empty() || memcmp(ca.data(), cb2.data(), ca.size() * sizeof(b2dgpu_edge)) == 0);
    if (!same) bad++;
  }
  return bad;
}

} // extern "C"
