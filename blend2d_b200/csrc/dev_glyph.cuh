// dev_glyph.cuh - TrueType glyph outline instancing (glyph outline decode + cache on the device).
//
// The reference decodes a `glyf` outline and applies the glyph's matrix in one pass
// (blend2d/opentype/otglyf.cpp:376-448, the same arithmetic in its SIMD form otglyfsimdimpl_p.h:765-978): the current
// point starts at the translation, every TrueType point adds  (dx * m00 + dy * m10,  dx * m01 + dy * m11)  to it in
// double precision, on-curve / off-curve points become MOVE / ON / QUAD vertices, two consecutive off-curve points get
// an implied on-point  current - delta * 0.5  between them, a contour that starts or ends off-curve is closed through
// the averaged points, and every contour ends with a CLOSE command that occupies one (NaN) vertex slot.
//
// Which vertices a glyph produces does not depend on the matrix - only where they are.  The glyph cache therefore keeps,
// per glyph, the TrueType deltas + on-curve bits + contour ends and the segment list of the resulting BLPath with
// vertex indices relative to the glyph (built once on the host from the reference's own decoder); `glyph_emit` replays
// the accumulation for one instance matrix and hands out the vertices in the reference's order.  The running sums round
// exactly like the reference's because the operations and their order are the same (no FMA, dev_flatten.cuh).
//
// Shared by the CUDA kernel (k_glyph_instances, kernels.cu) and, as plain C++, by the Blend2D-side binding (shim/),
// which runs it with the identity matrix against the reference's decoder before it trusts a cache entry.
//
// Cache blob, 32-bit words per glyph:
//   [0] point count | contour count << 16        [1] vertices per instance        [2] segments per instance
//   contour end indices, two 16-bit values per word
//   points: (uint16) dx | (uint16) dy << 16
//   on-curve bits, 32 points per word
//   segments: two words each: first vertex (relative), (second vertex (relative) << 2) | kind   (b2dgpu_segment layout)
#pragma once
#include "dev_common.cuh"

namespace b2d {

struct GlyphBlobView {
  const uint32_t* words;
  uint32_t points, contours, vertices, segments;
  B2D_HD const uint32_t* contour_words() const { return words + 3; }
  B2D_HD const uint32_t* point_words() const { return contour_words() + (contours + 1u) / 2u; }
  B2D_HD const uint32_t* on_curve_words() const { return point_words() + points; }
  B2D_HD const uint32_t* segment_words() const { return on_curve_words() + (points + 31u) / 32u; }
  B2D_HD uint32_t total_words() const { return uint32_t(segment_words() - words) + segments * 2u; }
  B2D_HD uint32_t contour_end(uint32_t c) const { return (contour_words()[c >> 1] >> ((c & 1u) * 16u)) & 0xFFFFu; }
  B2D_HD bool on_curve(uint32_t i) const { return (on_curve_words()[i >> 5] >> (i & 31u)) & 1u; }
};

B2D_HD GlyphBlobView glyph_blob_view(const uint32_t* words) {
  GlyphBlobView v;
  v.words = words;
  v.points = words[0] & 0xFFFFu; v.contours = words[0] >> 16;
  v.vertices = words[1]; v.segments = words[2];
  return v;
}

// Replays the decoder for the matrix m = { m00, m01, m10, m11, m20, m21 }.  `put(index, x, y)` receives vertex
// `index` of the glyph's path; CLOSE slots are skipped (no segment refers to them).  Returns the number of vertex slots,
// which equals GlyphBlobView::vertices for a well-formed entry.
template<typename Put>
B2D_HD uint32_t glyph_emit(const GlyphBlobView& g, const double* m, Put& put) {
  const double m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
  double cx = m[4], cy = m[5];
  uint32_t out = 0;
  uint32_t i = 0;
  for (uint32_t c = 0; c < g.contours; c++) {
    const uint32_t i_end = g.contour_end(c) + 1u;
    if (i_end <= i || i_end > g.points) return 0xFFFFFFFFu;

    // first point of the contour (otglyf.cpp:392-403)
    {
      const uint32_t w = g.point_words()[i];
      const double dx = double(int16_t(w & 0xFFFFu)), dy = double(int16_t(w >> 16));
      cx += dx * m00 + dy * m10;
      cy += dx * m01 + dy * m11;
    }
    const bool starts_on = g.on_curve(i);
    bool prev_on = starts_on;
    i++;
    if (i >= i_end) return 0xFFFFFFFFu;                  // one-point contour: the reference's variants differ; not cached

    const double ix = cx, iy = cy;                       // initial point
    const uint32_t first_index = out;
    double fx = 0.0, fy = 0.0;                           // vertex at first_index (read back when the contour starts off-curve)
    bool have_first = false;
    if (starts_on) { put(out, ix, iy); fx = ix; fy = iy; have_first = true; out++; }

    bool last_on = starts_on;
    for (; i < i_end; i++) {
      const uint32_t w = g.point_words()[i];
      const double dx = double(int16_t(w & 0xFFFFu)), dy = double(int16_t(w >> 16));
      const double ddx = dx * m00 + dy * m10, ddy = dx * m01 + dy * m11;
      cx += ddx; cy += ddy;
      const bool on = g.on_curve(i);
      if (!on && !prev_on) {
        // two off-curve points in a row: implied on-point half a delta back (otglyf.cpp:421-425)
        const double ox = cx - ddx * 0.5, oy = cy - ddy * 0.5;
        put(out, ox, oy);
        if (!have_first) { fx = ox; fy = oy; have_first = true; }
        out++;
      }
      put(out, cx, cy);
      if (!have_first) { fx = cx; fy = cy; have_first = true; }
      out++;
      prev_on = on;
      last_on = on;
    }

    if (!starts_on) {
      // the contour started off-curve (otglyf.cpp:431-444)
      double ex = fx, ey = fy;
      if (!last_on) {
        put(out, (cx + ix) * 0.5, (cy + iy) * 0.5); out++;
        ex = (ix + fx) * 0.5; ey = (iy + fy) * 0.5;
      }
      put(out, ix, iy); out++;
      put(out, ex, ey); out++;
    }
    else if (!last_on) {
      put(out, ix, iy); out++;
    }
    out++;                                               // CLOSE slot
    (void)first_index;
  }
  return out;
}

} // namespace b2d
