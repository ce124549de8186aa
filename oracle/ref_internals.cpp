// ref_internals.cpp - TEST INFRASTRUCTURE.  Entry points into the UNMODIFIED reference's private templates, compiled from
// /root/reference by oracle/Makefile.ref into oracle/_ref/libref_internals.so (the reference's own objects are linked in:
// its internal symbols are hidden in libblend2d_ref.so).  Nothing of the reference is copied: this file only
// instantiates and calls it.  Only tests/ may load the result.
//
//   ref_comp_op_pixels      CompOp_{SrcCopy,SrcOver,Plus}_Op::op_prgb32_prgb32(d, s, m)        pipeline/reference/compopgeneric_p.h:24-81
//   ref_fill_box_a_solid    FillDispatch<kBoxA, CompOp_Base<Op, P32, FetchSolid, 4>>::fill_func  pipeline/reference/fixedpiperuntime.cpp:58-69,
//                                                                                                fillgeneric_p.h:22-65
//   ref_build_edges         EdgeBuilder<int>::begin / add_path / done over an EdgeStorage<int>   raster/edgebuilder_p.h:934-1070,
//                                                                                                raster/edgestorage_p.h:38-178
//   ref_rasterize_edges     AnalyticRasterizer::prepare / rasterize<kOptionBandOffset> into cells       raster/analyticrasterizer_p.h:289-1210
#include <blend2d/core/api-build_p.h>
#include <blend2d/core/path_p.h>
#include <blend2d/pipeline/reference/compopgeneric_p.h>
#include <blend2d/pipeline/reference/fillgeneric_p.h>
#include <blend2d/raster/edgebuilder_p.h>
#include <blend2d/raster/analyticrasterizer_p.h>
#include <blend2d/support/arenaallocator_p.h>

#include <stdlib.h>
#include <string.h>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

using namespace bl;
using namespace bl::Pipeline;
typedef Reference::Pixel::P32_A8R8G8B8 P32;

template<typename Op>
static void comp_op_pixels(const uint32_t* d, const uint32_t* s, const uint32_t* m, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    P32 dp = P32::from_value(d[i]), sp = P32::from_value(s[i]);
    out[i] = (m[i] == 255u && Op::kOptimizeOpaque ? Op::op_prgb32_prgb32(dp, sp) : Op::op_prgb32_prgb32(dp, sp, m[i])).value();
  }
}

template<typename Op>
static FillFunc box_a_solid_func() {
  return Reference::FillDispatch<FillType::kBoxA, Reference::CompOp_Base<Op, P32, Reference::FetchSolid<P32>, 4>>::Fill::fill_func;
}

} // namespace

// comp_op: BL_COMP_OP_SRC_OVER (0), SRC_COPY (1), PLUS (12).  Returns 0 on success.
REF_API int ref_comp_op_pixels(uint32_t comp_op, const uint32_t* d, const uint32_t* s, const uint32_t* m, uint32_t* out, size_t n) {
  switch (comp_op) {
    case BL_COMP_OP_SRC_OVER: comp_op_pixels<Reference::CompOp_SrcOver_Op<P32>>(d, s, m, out, n); return 0;
    case BL_COMP_OP_SRC_COPY: comp_op_pixels<Reference::CompOp_SrcCopy_Op<P32>>(d, s, m, out, n); return 0;
    case BL_COMP_OP_PLUS:     comp_op_pixels<Reference::CompOp_Plus_Op<P32>>(d, s, m, out, n); return 0;
    default: return 1;
  }
}

// Fills [x0, x1) x [y0, y1) of a PRGB32 image through the reference's own FillBoxA pipeline with a solid source.
REF_API int ref_fill_box_a_solid(uint32_t comp_op, uint32_t* pixels, intptr_t stride, int w, int h, int x0, int y0, int x1, int y1,
                                 uint32_t prgb32, uint32_t alpha) {
  if (!(x0 >= 0 && y0 >= 0 && x1 <= w && y1 <= h && x0 < x1 && y0 < y1) || alpha > 255u) return 2;
  FillFunc fn = comp_op == BL_COMP_OP_SRC_OVER ? box_a_solid_func<Reference::CompOp_SrcOver_Op<P32>>()
              : comp_op == BL_COMP_OP_SRC_COPY ? box_a_solid_func<Reference::CompOp_SrcCopy_Op<P32>>()
              : comp_op == BL_COMP_OP_PLUS ? box_a_solid_func<Reference::CompOp_Plus_Op<P32>>() : nullptr;
  if (!fn) return 1;
  ContextData ctx_data;
  ctx_data.reset();
  ctx_data.dst.pixel_data = pixels;
  ctx_data.dst.stride = stride;
  ctx_data.dst.size.reset(w, h);
  ctx_data.dst.format = BL_FORMAT_PRGB32;
  FillData fill_data;
  fill_data.init_box_a_8bpc(alpha, x0, y0, x1, y1);
  FetchData fetch_data;
  memset(&fetch_data, 0, sizeof(fetch_data));
  fetch_data.solid.prgb32 = prgb32;
  fn(&ctx_data, &fill_data, &fetch_data);
  return 0;
}

// Runs the reference's EdgeBuilder on one path and returns its edges as lines in their ORIGINAL direction (x0, y0, x1,
// y1 in 24.8 fixed point), band list after band list.  `clip` = final_clip_box_fixed_d, `m` = the fixed-point final
// transform, `tolerance_sq` = toleranceFixedD squared.  Returns the number of lines (they are only stored while they fit).
REF_API int64_t ref_build_edges(const double* vertices, const uint8_t* commands, size_t n, int closed,
                                const double* m, uint32_t transform_type, const double* clip, double tolerance_sq,
                                int canvas_h, uint32_t band_height, int32_t* edges_out, size_t capacity) {
  using namespace bl::RasterEngine;
  if (!band_height || (band_height & (band_height - 1u))) return -1;
  const uint32_t band_count = (uint32_t(canvas_h) + band_height - 1u) / band_height;
  EdgeList<int>* lists = static_cast<EdgeList<int>*>(calloc(band_count + 1u, sizeof(EdgeList<int>)));
  if (!lists) return -1;

  int64_t count = -1;
  {
    ArenaAllocator arena(65536 - ArenaAllocator::kBlockOverhead, 8);
    EdgeStorage<int> storage;
    storage.init_data(lists, band_count, band_count, band_height);
    EdgeBuilder<int> builder(&arena, &storage, BLBox(clip[0], clip[1], clip[2], clip[3]), tolerance_sq);

    BLPathView view;
    view.command_data = commands;
    view.vertex_data = reinterpret_cast<const BLPoint*>(vertices);
    view.size = n;
    const BLMatrix2D transform(m[0], m[1], m[2], m[3], m[4], m[5]);

    builder.begin();
    BLResult r = builder.add_path(view, closed != 0, transform, BLTransformType(transform_type));
    if (r == BL_SUCCESS) r = builder.done();
    if (r == BL_SUCCESS) {
      count = 0;
      for (uint32_t b = 0; b < band_count; b++) {
        for (const EdgeVector<int>* ev = lists[b].first(); ev; ev = ev->next) {
          const size_t pts = ev->count();
          const bool flipped = ev->sign_bit() != 0;
          for (size_t i = 1; i < pts; i++) {
            if (size_t(count) < capacity) {
              int32_t* o = edges_out + size_t(count) * 4;
              const EdgePoint<int>& a = ev->pts[flipped ? i : i - 1];
              const EdgePoint<int>& c = ev->pts[flipped ? i - 1 : i];
              o[0] = a.x; o[1] = a.y; o[2] = c.x; o[3] = c.y;
            }
            count++;
          }
        }
      }
    }
  }
  free(lists);
  return count;
}

// Rasterizes `n` lines (x0, y0, x1, y1 in 24.8 fixed point, any direction, inside [0, w] x [0, h]) with the reference's
// AnalyticRasterizer as ONE band of h scanlines and returns the accumulated cells: cells_out[y * (w + 2) + x], x <= w + 1.
// This is the quantity FillAnalytic's scanline walk sums up (cover << 9 / area merged per cell, cell_merge :1202-1210).
REF_API int ref_rasterize_edges(const int32_t* lines, size_t n, int w, int h, uint32_t* cells_out) {
  using namespace bl::RasterEngine;
  if (w <= 0 || h <= 0) return 1;
  const size_t required_width = IntOps::align_up(uint32_t(w) + 1u + BL_PIPE_PIXELS_PER_ONE_BIT, BL_PIPE_PIXELS_PER_ONE_BIT);
  const size_t bit_stride = IntOps::word_count_from_bit_count<BLBitWord>(required_width / BL_PIPE_PIXELS_PER_ONE_BIT) * sizeof(BLBitWord);
  const size_t cell_stride = required_width * sizeof(uint32_t);
  const size_t bits_size = size_t(h) * bit_stride;
  const size_t cells_start = IntOps::align_up(bits_size, size_t(16));
  uint8_t* buffer = static_cast<uint8_t*>(calloc(cells_start + size_t(h) * cell_stride + 16, 1));
  if (!buffer) return 2;
  uint32_t* cells = IntOps::align_up(reinterpret_cast<uint32_t*>(buffer + cells_start), 16);

  AnalyticRasterizer ras;
  ras.init(reinterpret_cast<BLBitWord*>(buffer), bit_stride, cells, cell_stride, 0, uint32_t(h));
  for (size_t i = 0; i < n; i++) {
    EdgePoint<int> p0{lines[i * 4 + 0], lines[i * 4 + 1]}, p1{lines[i * 4 + 2], lines[i * 4 + 3]};
    uint32_t sign_bit = 0;
    if (p0.y > p1.y) { EdgePoint<int> t = p0; p0 = p1; p1 = t; sign_bit = 1; }
    ras.set_sign_mask_from_bit(sign_bit);
    if (!ras.prepare(p0, p1)) continue;
    ras.template rasterize<AnalyticRasterizer::kOptionBandOffset>();
  }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w + 2; x++)
      cells_out[size_t(y) * size_t(w + 2) + size_t(x)] = size_t(x) < required_width ? cells[size_t(y) * required_width + size_t(x)] : 0u;
  free(buffer);
  return 0;
}
