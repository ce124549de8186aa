// dev_tile.cuh - per-tile pieces of the compositor that are plain scalar code: the coverage sink, the classification of
// an edge against a tile, the axis-unaligned box mask and the command bounding box.  Shared by the CUDA kernel
// (kernels.cu) and the test-only host simulator (tests/hostsim).
#pragma once
#include "dev_common.cuh"
#include "dev_pixel.cuh"
#include "dev_raster.cuh"
#include "../../include/b2dgpu.h"

namespace b2d {

// Mask of an axis-unaligned box at pixel (x, y) - closed form of FillData::init_box_u_8bpc_24x8()'s mask-command
// program (pipeline/pipedefs_p.h:644-815) as interpreted by FillMask_Base (fillgeneric_p.h:67-162): per-column weight
// times per-row weight, with the reference's shifts.  Rows/columns whose mask is 0 are skipped there and here.
struct BoxUParams {
  int ax0, ay0, ax1, ay1;
  uint32_t wx_first, wx_last;      // weights of the first / last column (256 for interior columns)
  uint32_t fy0_a, fy1_a, alpha;    // row weights already multiplied by alpha
};

B2D_HD BoxUParams box_u_setup(const int32_t* box, uint32_t alpha) {
  BoxUParams b;
  int x0 = box[0], y0 = box[1], x1 = box[2], y1 = box[3];
  b.ax0 = int(uint32_t(x0) >> 8); b.ay0 = int(uint32_t(y0) >> 8);
  b.ax1 = int(uint32_t(x1 + 0xFF) >> 8); b.ay1 = int(uint32_t(y1 + 0xFF) >> 8);
  uint32_t fx0 = uint32_t(x0) & 0xFFu, fy0 = uint32_t(y0) & 0xFFu;
  uint32_t fx1 = (uint32_t(x1 - 1) & 0xFFu) + 1u, fy1 = (uint32_t(y1 - 1) & 0xFFu) + 1u;
  uint32_t w = uint32_t(b.ax1 - b.ax0), h = uint32_t(b.ay1 - b.ay0);
  fy0 = (h == 1 ? fy1 : 256u) - fy0;
  b.fy0_a = fy0 * alpha;
  b.fy1_a = fy1 * alpha;
  b.alpha = alpha;
  if (w == 1) { b.wx_first = fx1 - fx0; b.wx_last = b.wx_first; }
  else { b.wx_first = 256u - fx0; b.wx_last = fx1; }
  return b;
}

B2D_HD uint32_t box_u_mask(const BoxUParams& b, int x, int y) {
  if (x < b.ax0 || x >= b.ax1 || y < b.ay0 || y >= b.ay1) return 0u;
  uint32_t wx = (x == b.ax0) ? b.wx_first : (x == b.ax1 - 1) ? b.wx_last : 256u;
  if (y == b.ay0) return (wx * b.fy0_a) >> 16;
  if (y == b.ay1 - 1) return (wx * b.fy1_a) >> 16;
  return (wx * b.alpha) >> 8;
}

// Adapter from the rasterizer's merge(x, cover, area) to a tile: cells left of the tile go to the row's carry
// (backdrop), cells right of it are dropped.  `Store` provides add_cell(row, rel_x, v) and add_carry(row, v).
template<typename Store>
struct TileSink {
  Store& store;
  int tx0;
  int row;
  uint32_t touched;

  B2D_HD TileSink(Store& s, int tx0_) : store(s), tx0(tx0_), row(0), touched(0) {}
  B2D_HD void put(int x, uint32_t v) {
    if (!v) return;
    int rel = x - tx0;
    if (rel < 0) { store.add_carry(row, v); touched = 1; }
    else if (rel < kTileW) { store.add_cell(row, rel, v); touched = 1; }
  }
  B2D_HD void merge(int x, uint32_t cover, uint32_t area) {
    put(x, (cover << 9) - area);
    put(x + 1, area);
  }
  // window interface of edge_step_scanline<true>
  B2D_HD int win_lo() const { return tx0; }
  B2D_HD int win_hi() const { return tx0 + kTileW; }
  B2D_HD void add_left(uint32_t v) { if (v) { store.add_carry(row, v); touched = 1; } }
};

// Relation of an edge to the tile whose top-left pixel is (tx0, ty0).
enum : int {
  kEdgeNone = 0,        // no contribution: other rows, or entirely right of the tile
  kEdgeLeft = 1,        // entirely left: contributes only its signed y-extent per row to the backdrop
  kEdgeStraddle = 2     // touches the tile's columns (or the column before them): needs the rasterizer
};

struct NormEdge { int x0, y0, x1, y1; uint32_t sign_bit; };

// Edges are stored in their original direction; the rasterizer wants top -> bottom plus a sign bit.
B2D_HD NormEdge normalize_edge(b2dgpu_edge ed) {
  NormEdge n;
  if (ed.y0 > ed.y1) { n.x0 = ed.x1; n.y0 = ed.y1; n.x1 = ed.x0; n.y1 = ed.y0; n.sign_bit = 1; }
  else { n.x0 = ed.x0; n.y0 = ed.y0; n.x1 = ed.x1; n.y1 = ed.y1; n.sign_bit = 0; }
  return n;
}

// Columns an edge can touch inside the band of rows [band_y, band_y + kTileH): cells [min >> 8, (max >> 8) + 1]
// (cell_merge writes x and x + 1) of the edge's part inside the band, with one more column of slack on each side for
// the DDA's rounding.  Used by k_band_extents to cull (tile, command) pairs; must never be too tight.
B2D_HD void band_edge_extent(const NormEdge& ed, int band_y, int& lo, int& hi, int tile_h = kTileH) {
  // x at the band's first and last scanline boundary, by double-precision interpolation: |error| is far below the
  // one-pixel slack added on each side (the products are exact in a double, the quotient is off by < 1 ulp).
  const double slope = double(ed.x1 - ed.x0) / double(ed.y1 - ed.y0);
  const int ya = tmax(ed.y0, band_y << 8);
  const int yb = tmin(ed.y1, (band_y + tile_h) << 8);
  const int xa = ed.x0 + int(slope * double(ya - ed.y0));
  const int xb = ed.x0 + int(slope * double(yb - ed.y0));
  lo = tmax((tmin(xa, xb) >> 8) - 1, 0);
  hi = tmax((tmax(xa, xb) >> 8) + 2, 0);
}

B2D_HD int tile_edge_class(const NormEdge& ed, int tx0, int ty0, int tile_h = kTileH) {
  const int ey_first = ed.y0 >> 8, ey_last = (ed.y1 - 1) >> 8;
  if (ey_last < ty0 || ey_first >= ty0 + tile_h) return kEdgeNone;
  int cx_min = tmin(ed.x0, ed.x1) >> 8, cx_max = (tmax(ed.x0, ed.x1) >> 8) + 1;
  if (cx_min >= tx0 + kTileW) return kEdgeNone;
  if (cx_max < tx0) return kEdgeLeft;
  // A long edge (a closing chord can span the whole canvas) only touches a few columns inside this tile's rows:
  // classify by the part of the edge inside the band, conservatively rounded outwards (band_edge_extent).
  if (cx_max - cx_min >= 16) {
    band_edge_extent(ed, ty0, cx_min, cx_max, tile_h);
    if (cx_min >= tx0 + kTileW) return kEdgeNone;
    if (cx_max < tx0) return kEdgeLeft;
  }
  return kEdgeStraddle;
}

// Every scanline's cells of an edge sum to (cover << 9), cover = signed y-extent inside the row: that is all a tile
// to the right of the edge needs.  One accumulator per tile row.
B2D_HD void tile_left_cover(const NormEdge& ed, int ty0, uint32_t* left_acc) {
  #pragma unroll
  for (int r = 0; r < kTileH; r++) {
    int yt = (ty0 + r) << 8;
    int cov = tmin(ed.y1, yt + 256) - tmax(ed.y0, yt);
    if (cov > 0) left_acc[r] += uint32_t(ed.sign_bit ? -cov : cov) << 9;
  }
}

// Rasterizes the tile's rows of a straddling edge through `store`.  Returns true when anything was written.
template<typename Store>
B2D_HD bool tile_rasterize_edge(const NormEdge& ed, int tx0, int ty0, Store& store) {
  EdgeState st;
  if (!edge_prepare(st, ed.x0, ed.y0, ed.x1, ed.y1, ed.sign_bit)) return false;
  const int y_from = tmax(ed.y0 >> 8, ty0);
  const int y_to = tmin((ed.y1 - 1) >> 8, ty0 + kTileH - 1);
  edge_advance_to_y(st, y_from);
  TileSink<Store> sink(store, tx0);
  for (int y = y_from; y <= y_to; y++) {
    sink.row = y - ty0;
    if (edge_step_scanline(st, sink)) break;
  }
  return sink.touched != 0;
}

// One scanline `y` (absolute, inside the edge's y-range) of an edge: prepare, closed-form jump, one step.  This is the
// unit of work of a GPU lane - the reference's own unit test pins that jumping with advanceToY() equals stepping
// (raster/analyticrasterizer_test.cpp:34-157).
template<typename Sink>
B2D_HD void tile_rasterize_edge_row(const NormEdge& ed, int y, Sink& sink) {
  EdgeState st;
  if (!edge_prepare(st, ed.x0, ed.y0, ed.x1, ed.y1, ed.sign_bit)) return;
  edge_advance_to_y(st, y);
  edge_step_scanline<true>(st, sink);                  // windowed: a shallow edge's cells outside the tile are not walked
}

B2D_HD bool command_has_edges(uint32_t type) { return type == B2DGPU_CMD_FILL_ANALYTIC || type == B2DGPU_CMD_FILL_GEOMETRY; }

// FillBoxMaskA (rendercommandprocsync_p.h:65-86): one VMask span per row of the box.  alpha == 255 takes the mask byte
// as it is (kVMaskA8WithGA), otherwise m = udiv255(mask * alpha) (kVMaskA8WithoutGA, compopgeneric_p.h:175-183).
B2D_HD uint32_t box_mask_a(const b2dgpu_command& cmd, const b2dgpu_pattern_source& mask, int x, int y) {
  if (x < cmd.box[0] || x >= cmd.box[2] || y < cmd.box[1] || y >= cmd.box[3]) return 0u;
  const uint8_t* row = static_cast<const uint8_t*>(mask.pixel_data) + intptr_t(y - cmd.box[1]) * mask.stride;
  uint32_t m = row[x - cmd.box[0]];
  return cmd.alpha >= 255u ? m : udiv255(m * cmd.alpha);
}

// Pixel bounding box [x0,x1) x [y0,y1) of a command, clipped to the rows [y_begin, y_end) of a `width`-wide target;
// all zeros when the command cannot touch it.  `bb_fixed` = 24.8 bounds of an analytic command's edges.
struct CmdBox { int x0, y0, x1, y1; };

B2D_HD CmdBox command_pixel_box(const b2dgpu_command& cmd, uint32_t edge_count, int fx0, int fy0, int fx1, int fy1,
                                int width, int y_begin, int y_end) {
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  if (cmd.type == B2DGPU_CMD_FILL_BOX_A || cmd.type == B2DGPU_CMD_FILL_BOX_MASK_A) {
    x0 = cmd.box[0]; y0 = cmd.box[1]; x1 = cmd.box[2]; y1 = cmd.box[3];
  }
  else if (cmd.type == B2DGPU_CMD_FILL_BOX_U) {
    x0 = cmd.box[0] >> 8; y0 = cmd.box[1] >> 8; x1 = (cmd.box[2] + 0xFF) >> 8; y1 = (cmd.box[3] + 0xFF) >> 8;
  }
  else if (edge_count && fx0 <= fx1 && fy0 < fy1) {
    x0 = fx0 >> 8; x1 = (fx1 >> 8) + 1;
    y0 = fy0 >> 8; y1 = ((fy1 - 1) >> 8) + 1;
  }
  x0 = tmax(x0, 0); y0 = tmax(y0, y_begin); x1 = tmin(x1, width); y1 = tmin(y1, y_end);
  CmdBox b;
  if (x0 >= x1 || y0 >= y1 || cmd.alpha == 0) { b.x0 = b.y0 = b.x1 = b.y1 = 0; }
  else { b.x0 = x0; b.y0 = y0; b.x1 = x1; b.y1 = y1; }
  return b;
}

} // namespace b2d
