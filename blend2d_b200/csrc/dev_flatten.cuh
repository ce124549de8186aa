// dev_flatten.cuh - the edge builder: one path segment -> clipped, flattened 24.8 fixed-point edges.
//
// What the reference does (blend2d/raster/edgebuilder_p.h, blend2d/geometry/bezier_p.h) as ONE sequential state
// machine over a whole path, restated here as an independent function PER SEGMENT so that segments can be processed
// by independent GPU threads:
//   line_to()                     edgebuilder_p.h:1123-1605  clip flags, start/end point clipping, border lines
//   quad_to()/cubic_to()/conic_to :1618-2022                 reject, monotone split, safe/unsafe flattening
//   flatten_safe_mono_curve()     :2029-2063                 midpoint subdivision, explicit stack of depth 32
//   flatten_unsafe_mono_curve()   :2071-2445                 subdivision interleaved with clipping
//   border accumulation           :2546-2622                 lines left/right of the clip box become vertical
//                                                            lines AT the clip edge
//   Geometry::split_with_options  bezier_p.h:366-400, split_with_ts :332-364
//   Geometry::split_cubic_to_spline bezier_p.h:834-920, Math::quad_roots support/math_p.h:574-593
//
// Why per-segment processing gives the same pixels: the analytic rasterizer sums (cover, area) contributions of
// individual lines into u32 cells, so only the MULTISET of integer lines matters, not their order or their grouping
// into edge vectors.  The reference merges consecutive border intervals before truncating them to integers
// (`accumulate_*_border`); un-merged intervals [a,b],[b,c] truncate the shared point identically and vertical lines
// at the same x add linearly in y, so they produce the same cells as the merged [a,c].
//
// All arithmetic is IEEE double with the reference's evaluation order and NO fused multiply-add (the reference
// build has no -mfma; this file must be compiled with nvcc -fmad=false / g++ -ffp-contract=off).
#pragma once
#include "dev_common.cuh"
#include <math.h>

namespace b2d {

struct P2 { double x, y; };
B2D_HD P2 mk(double x, double y) { P2 p; p.x = x; p.y = y; return p; }
B2D_HD P2 operator+(P2 a, P2 b) { return mk(a.x + b.x, a.y + b.y); }
B2D_HD P2 operator-(P2 a, P2 b) { return mk(a.x - b.x, a.y - b.y); }
B2D_HD P2 operator*(P2 a, P2 b) { return mk(a.x * b.x, a.y * b.y); }
B2D_HD P2 operator/(P2 a, P2 b) { return mk(a.x / b.x, a.y / b.y); }
B2D_HD P2 operator*(P2 a, double s) { return mk(a.x * s, a.y * s); }
B2D_HD P2 operator*(double s, P2 a) { return mk(s * a.x, s * a.y); }

B2D_HD double cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }       // Geometry::cross (commons_p.h:64)
B2D_HD double mag_sq(P2 v) { return v.x * v.x + v.y * v.y; }              // Geometry::magnitude_squared (:66)

// Math::trunc_to_int is a C cast (support/math_p.h:345-346): cvttsd2si returns INT_MIN when out of range / NaN.
B2D_HD int trunc_i(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return int(0x80000000u);
  return int(v);
}

struct ClipBox { double x0, y0, x1, y1; int ix0, ix1; };

enum : uint32_t { kClipX0 = 1u, kClipY0 = 2u, kClipX1 = 4u, kClipY1 = 8u };   // edgebuilder_p.h:25-47

B2D_HD uint32_t clip_x_flags(P2 p, const ClipBox& c) { return (uint32_t(!(p.x >= c.x0)) << 0) | (uint32_t(!(p.x <= c.x1)) << 2); }
B2D_HD uint32_t clip_y_flags(P2 p, const ClipBox& c) { return (uint32_t(!(p.y >= c.y0)) << 1) | (uint32_t(!(p.y <= c.y1)) << 3); }
B2D_HD uint32_t clip_flags(P2 p, const ClipBox& c) { return clip_x_flags(p, c) | clip_y_flags(p, c); }

// -----------------------------------------------------------------------------------------------------------------
// Edge output.  `Out::edge(x0, y0, x1, y1)` receives one integer line in its ORIGINAL direction (y0 != y1); y0 > y1
// is what the reference stores as "sign bit set, points reversed".
// -----------------------------------------------------------------------------------------------------------------
template<typename Out>
struct EdgeEmitter {
  Out& out;
  const ClipBox& clip;

  B2D_HD EdgeEmitter(Out& o, const ClipBox& c) : out(o), clip(c) {}

  // add_line_segment() (:1072-1091) and the unclipped line path (:1142-1152).
  B2D_HD void line(double x0, double y0, double x1, double y1) {
    int fx0 = trunc_i(x0), fy0 = trunc_i(y0), fx1 = trunc_i(x1), fy1 = trunc_i(y1);
    if (fy0 != fy1) out.edge(fx0, fy0, fx1, fy1);
  }

  // accumulate_left_border()/accumulate_right_border() + _emit_*_border() (:2556-2622), un-merged.
  B2D_HD void border(bool left, double y0, double y1) {
    int fy0 = trunc_i(y0), fy1 = trunc_i(y1);
    if (fy0 != fy1) {
      int x = left ? clip.ix0 : clip.ix1;
      out.edge(x, fy0, x, fy1);
    }
  }
  B2D_HD void border_signed(bool left, double y0, double y1, uint32_t sign_bit) {
    if (sign_bit) border(left, y1, y0); else border(left, y0, y1);
  }
};

// -----------------------------------------------------------------------------------------------------------------
// Lines (line_to, :1123-1605), one line a -> b at a time.
// -----------------------------------------------------------------------------------------------------------------
template<typename Out>
B2D_HD void build_line(P2 a, P2 b, const ClipBox& c, Out& out) {
  EdgeEmitter<Out> em(out, c);
  uint32_t a_flags = clip_flags(a, c);
  uint32_t b_flags = clip_flags(b, c);

  P2 p, d;
  double bor_y0 = 0.0, bor_y1;

  if (!a_flags) {
    if (!b_flags) { em.line(a.x, a.y, b.x, b.y); return; }
    p = a;
    d = b - a;
  }
  else {
    if (a_flags & kClipY0) {
      if (!(c.y0 < b.y)) return;                                 // completely above
      a_flags = clip_x_flags(a, c) | (uint32_t(!(a.y >= c.y0)) << 1);
      b_flags = clip_x_flags(b, c) | (uint32_t(!(b.y <= c.y1)) << 3);
      bor_y0 = c.y0;
      uint32_t common = a_flags & b_flags;
      if (common) {
        bor_y1 = tmin(c.y1, b.y);
        em.border((common & kClipX0) != 0, bor_y0, bor_y1);
        return;
      }
    }
    else if (a_flags & kClipY1) {
      if (!(c.y1 > b.y)) return;                                 // completely below
      a_flags = clip_x_flags(a, c) | (uint32_t(!(a.y <= c.y1)) << 3);
      b_flags = clip_x_flags(b, c) | (uint32_t(!(b.y >= c.y0)) << 1);
      bor_y0 = c.y1;
      uint32_t common = a_flags & b_flags;
      if (common) {
        bor_y1 = tmax(c.y0, b.y);
        em.border((common & kClipX0) != 0, bor_y0, bor_y1);
        return;
      }
    }
    else if (a_flags & kClipX0) {
      bor_y0 = tclamp(a.y, c.y0, c.y1);
      if (!(c.x0 < b.x)) {                                       // completely left
        bor_y1 = tclamp(b.y, c.y0, c.y1);
        if (bor_y0 != bor_y1) em.border(true, bor_y0, bor_y1);
        return;
      }
      a_flags = (uint32_t(!(a.x >= c.x0)) << 0) | clip_y_flags(a, c);
      b_flags = (uint32_t(!(b.x <= c.x1)) << 2) | clip_y_flags(b, c);
    }
    else {
      bor_y0 = tclamp(a.y, c.y0, c.y1);
      if (!(c.x1 > b.x)) {                                       // completely right
        bor_y1 = tclamp(b.y, c.y0, c.y1);
        if (bor_y0 != bor_y1) em.border(false, bor_y0, bor_y1);
        return;
      }
      a_flags = (uint32_t(!(a.x <= c.x1)) << 2) | clip_y_flags(a, c);
      b_flags = (uint32_t(!(b.x >= c.x0)) << 0) | clip_y_flags(b, c);
    }

    // Clip the start point (:1433-1497).
    d = b - a;
    p = mk(c.x1, c.y1);

    switch (a_flags) {
      case 0: p = a; break;
      case kClipX0 | kClipY0: p.x = c.x0;  // fallthrough
      case kClipX1 | kClipY0:
        p.y = a.y + (p.x - a.x) * d.y / d.x;
        a_flags = clip_y_flags(p, c);
        if (p.y >= c.y0) break;
        // fallthrough
      case kClipY0:
        p.y = c.y0;
        p.x = a.x + (p.y - a.y) * d.x / d.y;
        a_flags = clip_x_flags(p, c);
        break;
      case kClipX0 | kClipY1: p.x = c.x0;  // fallthrough
      case kClipX1 | kClipY1:
        p.y = a.y + (p.x - a.x) * d.y / d.x;
        a_flags = clip_y_flags(p, c);
        if (p.y <= c.y1) break;
        // fallthrough
      case kClipY1:
        p.y = c.y1;
        p.x = a.x + (p.y - a.y) * d.x / d.y;
        a_flags = clip_x_flags(p, c);
        break;
      case kClipX0: p.x = c.x0;            // fallthrough
      case kClipX1:
        p.y = a.y + (p.x - a.x) * d.y / d.x;
        a_flags = clip_y_flags(p, c);
        break;
      default:
        return;                                                  // NaNs: BL_ERROR_INVALID_GEOMETRY
    }

    if (a_flags) {
      // The line never enters the clip box.
      bor_y1 = tclamp(b.y, c.y0, c.y1);
      if (p.x <= c.x0) em.border(true, bor_y0, bor_y1);
      else if (p.x >= c.x1) em.border(false, bor_y0, bor_y1);
      return;
    }

    bor_y1 = tclamp(p.y, c.y0, c.y1);
    if (bor_y0 != bor_y1) em.border(p.x <= c.x0, bor_y0, bor_y1);

    if (!b_flags) { em.line(p.x, p.y, b.x, b.y); return; }
  }

  // Clip the end point (:1538-1598).
  P2 q = mk(c.x1, c.y1);
  switch (b_flags) {
    case kClipX0 | kClipY0: q.x = c.x0;    // fallthrough
    case kClipX1 | kClipY0:
      q.y = a.y + (q.x - a.x) * d.y / d.x;
      if (q.y >= c.y0) break;
      // fallthrough
    case kClipY0:
      q.y = c.y0;
      q.x = a.x + (q.y - a.y) * d.x / d.y;
      break;
    case kClipX0 | kClipY1: q.x = c.x0;    // fallthrough
    case kClipX1 | kClipY1:
      q.y = a.y + (q.x - a.x) * d.y / d.x;
      if (q.y <= c.y1) break;
      // fallthrough
    case kClipY1:
      q.y = c.y1;
      q.x = a.x + (q.y - a.y) * d.x / d.y;
      break;
    case kClipX0: q.x = c.x0;              // fallthrough
    case kClipX1:
      q.y = a.y + (q.x - a.x) * d.y / d.x;
      break;
    default:
      return;
  }

  em.line(p.x, p.y, q.x, q.y);
  double clipped_by = tclamp(b.y, c.y0, c.y1);
  if (q.y != clipped_by) em.border(q.x == c.x0, q.y, clipped_by);
}

// -----------------------------------------------------------------------------------------------------------------
// Monotone-curve flattening.  N = 3 (quad / conic) or 4 (cubic).  FlattenMonoQuad/Cubic (:440-663).
// -----------------------------------------------------------------------------------------------------------------
enum : int { kFlattenRecursionLimit = 32 };

template<int N>
struct MonoCurve {
  P2 p[N];
  P2 stack[kFlattenRecursionLimit * N];
  int sp;
  int base;          // stack level this walk started at (see begin_at)
  double tol_sq;

  struct Step { double value, limit; P2 a, b, c, d, e, mid; };

  B2D_HD void begin(const P2* src, uint32_t sign_bit) {
    sp = 0; base = 0;
    #pragma unroll
    for (int i = 0; i < N; i++) p[i] = sign_bit ? src[N - 1 - i] : src[i];
  }
  // Starts a walk at a node INSIDE the subdivision tree of a monotone piece: `pending` = the number of second halves
  // the reference's depth-first walk would hold on its stack when it reaches this node (one per "first half" step on
  // the path from the root), so that can_push() fails at exactly the same nodes.  The walk never pops below it.
  B2D_HD void begin_at(const P2* node, uint32_t pending) {
    sp = base = int(pending) * N;
    #pragma unroll
    for (int i = 0; i < N; i++) p[i] = node[i];
  }
  B2D_HD P2 first() const { return p[0]; }
  B2D_HD P2 last() const { return p[N - 1]; }
  B2D_HD bool can_pop() const { return sp != base; }
  B2D_HD bool can_push() const { return sp != kFlattenRecursionLimit * N; }

  // bound_left_to_right()/bound_right_to_left(): keep control points inside the end-point box.
  B2D_HD void bound(bool left_to_right) {
    double xlo = left_to_right ? p[0].x : p[N - 1].x;
    double xhi = left_to_right ? p[N - 1].x : p[0].x;
    #pragma unroll
    for (int i = 1; i < N - 1; i++) {
      p[i].x = tclamp(p[i].x, xlo, xhi);
      p[i].y = tclamp(p[i].y, p[0].y, p[N - 1].y);
    }
  }

  B2D_HD bool is_flat(Step& st) const {
    if (N == 3) {
      P2 v1 = p[1] - p[0];
      P2 v2 = p[2] - p[0];
      double dd = cross2(v2, v1);
      double len_sq = mag_sq(v2);
      st.value = dd * dd;
      st.limit = tol_sq * len_sq;
    }
    else {
      P2 v = p[N - 1] - p[0];
      double c1 = cross2(v, p[1] - p[0]);
      double c2 = cross2(v, p[N - 2] - p[0]);
      double d1 = c1 * c1;
      double d2 = c2 * c2;
      double len_sq = mag_sq(v);
      st.value = tmax(d1, d2);
      st.limit = tol_sq * len_sq;
    }
    return st.value <= st.limit;
  }

  B2D_HD void split(Step& st) const {
    if (N == 3) {
      st.a = (p[0] + p[1]) * 0.5;          // p01
      st.b = (p[1] + p[2]) * 0.5;          // p12
      st.mid = (st.a + st.b) * 0.5;        // p012
    }
    else {
      st.a = (p[0] + p[1]) * 0.5;          // p01
      st.b = (p[1] + p[2]) * 0.5;          // p12
      st.c = (p[2] + p[N - 1]) * 0.5;      // p23
      st.d = (st.a + st.b) * 0.5;          // p012
      st.e = (st.b + st.c) * 0.5;          // p123
      st.mid = (st.d + st.e) * 0.5;        // p0123
    }
  }

  B2D_HD void push(const Step& st) {
    if (N == 3) {
      stack[sp + 0] = st.mid; stack[sp + 1] = st.b; stack[sp + 2] = p[2];
      sp += 3;
      p[1] = st.a; p[2] = st.mid;
    }
    else {
      stack[sp + 0] = st.mid; stack[sp + 1] = st.e; stack[sp + 2] = st.c; stack[sp + 3] = p[N - 1];
      sp += 4;
      p[1] = st.a; p[2] = st.d; p[N - 1] = st.mid;
    }
  }

  B2D_HD void discard_and_advance(const Step& st) {
    if (N == 3) { p[0] = st.mid; p[1] = st.b; }
    else { p[0] = st.mid; p[1] = st.e; p[2] = st.c; }
  }

  B2D_HD void pop() {
    sp -= N;
    #pragma unroll
    for (int i = 0; i < N; i++) p[i] = stack[sp + i];
  }
};

// Appender (:835-879): a chain of truncated points of one monotone (y non-decreasing) piece.
template<typename Out>
struct Chain {
  Out& out;
  uint32_t sign_bit;
  int px, py;

  B2D_HD Chain(Out& o, uint32_t s) : out(o), sign_bit(s), px(0), py(0) {}
  B2D_HD void open_at(double x, double y) { px = trunc_i(x); py = trunc_i(y); }
  B2D_HD void add_line(double x, double y) {
    int fx = trunc_i(x), fy = trunc_i(y);
    if (fy != py) {
      if (sign_bit) out.edge(fx, fy, px, py); else out.edge(px, py, fx, fy);
    }
    px = fx; py = fy;
  }
};

// flatten_safe_mono_curve (:2029-2063).
// The walk of flatten_safe_mono_curve from the current node of `mc` (the whole piece, or - device edge builder - one
// node of its subdivision tree: consecutive nodes share their end point, so each walk opens its chain where the
// previous one ended and the union of the emitted lines is the reference's chain).
template<int N, typename Out>
B2D_HD void safe_walk(MonoCurve<N>& mc, uint32_t sign_bit, Out& out) {
  Chain<Out> chain(out, sign_bit);
  chain.open_at(mc.first().x, mc.first().y);
  for (;;) {
    typename MonoCurve<N>::Step st;
    if (!mc.is_flat(st) && mc.can_push()) {
      mc.split(st);
      mc.push(st);
      continue;
    }
    chain.add_line(mc.last().x, mc.last().y);
    if (!mc.can_pop()) break;
    mc.pop();
  }
}

template<int N, typename Out>
B2D_HD void flatten_safe(MonoCurve<N>& mc, const P2* src, uint32_t sign_bit, Out& out) {
  mc.begin(src, sign_bit);
  mc.bound(mc.first().x < mc.last().x);

  Chain<Out> chain(out, sign_bit);
  chain.open_at(mc.first().x, mc.first().y);
  for (;;) {
    typename MonoCurve<N>::Step st;
    if (!mc.is_flat(st) && mc.can_push()) {
      mc.split(st);
      mc.push(st);
      continue;
    }
    chain.add_line(mc.last().x, mc.last().y);
    if (!mc.can_pop()) break;
    mc.pop();
  }
}

// flatten_unsafe_mono_curve (:2071-2445).  The reference has two textually mirrored branches (left-to-right and
// right-to-left); here a right-to-left piece is reflected in x (x -> -x, clip x0/x1 swapped and negated), run through
// the single left-to-right routine, and reflected back right before truncation.  Negation is exact in IEEE
// arithmetic and every comparison/min/max in the reference's second branch is the mirror image of the first.
// First half of flatten_unsafe_mono_curve: everything that looks at the WHOLE monotone piece - the early outs, the
// "practically a vertical line" case, the reflection of a right-to-left piece and bound().  Returns false when the piece
// is finished (nothing visible, or the vertical case emitted it); otherwise `mc` holds the piece in the left-to-right
// frame and `ltr` says whether that frame is the original one.
template<int N, typename Out>
B2D_HD bool unsafe_prepare(MonoCurve<N>& mc, const P2* src, uint32_t sign_bit, const ClipBox& c, Out& out, bool& ltr) {
  EdgeEmitter<Out> em(out, c);
  mc.begin(src, sign_bit);

  double y_start = mc.first().y;
  double y_end = tmin(mc.last().y, c.y1);
  if ((y_start >= y_end) | (y_end <= c.y0)) return false;

  const double kDeltaLimit = 0.00390625;
  double x_delta = mc.first().x - mc.last().x;
  x_delta = x_delta < 0.0 ? -x_delta : x_delta;                       // bl_abs

  if (x_delta <= kDeltaLimit) {
    // Practically a vertical line.
    Chain<Out> chain(out, sign_bit);
    y_start = tmax(y_start, c.y0);
    double x_min = tmin(mc.first().x, mc.last().x);
    double x_max = tmax(mc.first().x, mc.last().x);
    if (x_max <= c.x0) em.border_signed(true, y_start, y_end, sign_bit);
    else if (x_min >= c.x1) em.border_signed(false, y_start, y_end, sign_bit);
    else {
      chain.open_at(mc.first().x, y_start);
      chain.add_line(mc.last().x, y_end);
    }
    return false;
  }

  ltr = mc.first().x < mc.last().x;
  if (!ltr) {
    #pragma unroll
    for (int i = 0; i < N; i++) mc.p[i].x = -mc.p[i].x;
  }
  mc.bound(true);
  return true;
}

// Second half: the clipped walk from the current node of `mc` (in the left-to-right frame).  For the whole piece this
// is the reference's walk.  The device edge builder also runs it on the nodes of the piece's subdivision tree one by
// one: the piece is monotone in x and y, so the state the reference carries from one node to the next - which side of
// the clip box it is on, where the open chain ends - is a function of the node's first point, which every walk
// re-derives below; borders come out as one interval per node instead of one per piece (un-merged, see the top of
// this file).  tests/hostsim checks the equivalence on random curves.
template<int N, typename Out>
B2D_HD void unsafe_walk(MonoCurve<N>& mc, bool ltr, uint32_t sign_bit, const ClipBox& c, Out& out) {
  EdgeEmitter<Out> em(out, c);
  double y_start = mc.first().y;
  double y_end = tmin(mc.last().y, c.y1);
  if ((y_start >= y_end) | (y_end <= c.y0)) return;

  Chain<Out> chain(out, sign_bit);
  const double sgn = ltr ? 1.0 : -1.0;       // reflection factor applied to every x that leaves this routine
  double cx0 = c.x0, cx1 = c.x1;             // clip x range in the (possibly reflected) frame
  if (!ltr) { cx0 = -c.x1; cx1 = -c.x0; }
  // In the reflected frame "x0 side" is the reference's x1 side: borders go to the opposite clip edge.
  const bool x0_is_left = ltr;

  // 0 = not out, 1 = out on the frame's x0 side, 2 = out on the frame's x1 side.
  uint32_t out_side = 0;
  typename MonoCurve<N>::Step st;

  enum { kEnterNone, kEnterAddLine, kEnterX0Clip, kEnterX0Pop };
  int enter = kEnterNone;

  if (y_start < c.y0) {
    // Above the clip box: subdivide until the piece crossing y0 is found (:2121-2171).
    y_start = c.y0;
    bool o = false;
    for (;;) {
      o = (mc.first().x >= cx1);
      if (o) break;

      if (!mc.is_flat(st)) {
        mc.split(st);
        if (st.mid.y <= c.y0) { mc.discard_and_advance(st); continue; }
        if (mc.can_push()) { mc.push(st); continue; }
      }

      if (mc.last().y > c.y0) {
        o = mc.last().x < cx0;
        if (o) { enter = kEnterX0Pop; break; }

        P2 dd = mc.last() - mc.first();
        double x_clipped = mc.first().x + (c.y0 - mc.first().y) * (dd.x / dd.y);
        if (x_clipped <= cx0) { enter = kEnterX0Clip; break; }

        o = (x_clipped >= cx1);
        if (o) break;

        chain.open_at(sgn * x_clipped, c.y0);
        enter = kEnterAddLine;
        break;
      }

      if (!mc.can_pop()) break;
      mc.pop();
    }
    if (enter == kEnterNone) {
      out_side = o ? 2u : 0u;
      goto Finish;
    }
  }
  else if (!(y_start < c.y1)) {
    return;                                                       // below the bottom
  }

  if (enter == kEnterNone) {
    if (mc.first().x < cx0) enter = -1;                           // "before x0" loop from its top
    else if (mc.first().x < cx1) { chain.open_at(sgn * mc.first().x, mc.first().y); enter = -2; }  // visible
    else { out_side = 2u; goto Finish; }
  }

  if (enter == -1 || enter == kEnterX0Clip || enter == kEnterX0Pop) {
    // Left of the clip box: subdivide until the piece crossing x0 is found (:2173-2222).
    bool o = false;
    bool found = false;
    for (;;) {
      if (enter == kEnterX0Clip) { enter = -1; goto X0Clip; }
      if (enter == kEnterX0Pop) { enter = -1; o = true; goto X0Pop; }

      o = (mc.first().y >= c.y1);
      if (o) break;

      if (!mc.is_flat(st)) {
        mc.split(st);
        if (st.mid.x <= cx0) { mc.discard_and_advance(st); continue; }
        if (mc.can_push()) { mc.push(st); continue; }
      }

      if (mc.last().x > cx0) {
X0Clip:
        P2 dd = mc.last() - mc.first();
        double y_clipped = mc.first().y + (cx0 - mc.first().x) * (dd.y / dd.x);
        o = (y_clipped >= y_end);
        if (o) break;

        if (y_start < y_clipped) em.border_signed(x0_is_left, y_start, y_clipped, sign_bit);
        chain.open_at(sgn * cx0, y_clipped);
        found = true;
        break;
      }

      o = (mc.last().y >= y_end);
      if (o) break;
X0Pop:
      if (!mc.can_pop()) break;
      mc.pop();
    }
    if (!found) {
      out_side = o ? 1u : 0u;
      goto Finish;
    }
    enter = kEnterAddLine;
  }

  {
    // Visible part (:2223-2265).
    bool o = false;
    for (;;) {
      if (enter == kEnterAddLine) { enter = -2; goto AddLine; }

      if (!mc.is_flat(st)) {
        mc.split(st);
        if (mc.can_push()) { mc.push(st); continue; }
      }
AddLine:
      o = mc.last().x > cx1;
      if (o) {
        P2 dd = mc.last() - mc.first();
        double y_clipped = mc.first().y + (cx1 - mc.first().x) * (dd.y / dd.x);
        if (y_clipped <= y_end) {
          y_start = y_clipped;
          chain.add_line(sgn * cx1, y_clipped);
          break;
        }
      }

      o = mc.last().y >= c.y1;
      if (o) {
        P2 dd = mc.last() - mc.first();
        double x_clipped = tmin(mc.first().x + (c.y1 - mc.first().y) * (dd.x / dd.y), cx1);
        chain.add_line(sgn * x_clipped, c.y1);
        o = false;
        break;
      }

      chain.add_line(sgn * mc.last().x, mc.last().y);
      if (!mc.can_pop()) break;
      mc.pop();
    }
    out_side = o ? 2u : 0u;
  }

Finish:
  if (out_side && y_start < y_end) {
    bool frame_x0 = (out_side == 1u);
    bool left = frame_x0 ? x0_is_left : !x0_is_left;
    em.border_signed(left, y_start, y_end, sign_bit);
  }
}

template<int N, typename Out>
B2D_HD void flatten_unsafe(MonoCurve<N>& mc, const P2* src, uint32_t sign_bit, const ClipBox& c, Out& out) {
  bool ltr = true;
  if (unsafe_prepare<N>(mc, src, sign_bit, c, out, ltr)) unsafe_walk<N>(mc, ltr, sign_bit, c, out);
}

// -----------------------------------------------------------------------------------------------------------------
// Curve front-ends (quad_to :1618-1743, cubic_to :1757-1884, conic_to :1897-2022).
// -----------------------------------------------------------------------------------------------------------------

// Geometry::split_with_ts (bezier_p.h:332-364) for a quad whose t values (incl. the final 1.0) are in ts[0..n).
B2D_HD int split_quad_at(const P2* curve, P2* outp, const double* ts, int n) {
  P2 v1 = curve[1] - curve[0];
  P2 v2 = curve[2] - curve[1];
  P2 qa = v2 - v1, qb = v1 + v1, qc = curve[0];                    // coefficients_of (:195-199)

  int pieces = 0;
  double t_cut = 0.0;
  outp[0] = curve[0];
  P2 last = curve[2];
  for (int i = 0; i < n; i++) {
    double t_val = ts[i];
    double dt = (t_val - t_cut) * 0.5;
    P2 cp = (qa * (t_val * 2.0) + qb) * dt;
    P2 tp = (qa * t_val + qb) * t_val + qc;
    if (i + 1 == n) tp = last;
    outp[1] = tp - cp;
    outp[2] = tp;
    outp += 2;
    pieces++;
    t_cut = t_val;
  }
  return pieces;
}

// quad_to up to the monotone pieces: the trivial reject / border case is emitted here (returns 0), otherwise
// spline[0 .. 2 * pieces] holds the pieces and `any` tells whether a control point lies outside the clip box (unsafe).
template<typename Out>
B2D_HD int prepare_quad(P2 p0, P2 p1, P2 p2, const ClipBox& c, Out& out, P2* spline /* [7] */, uint32_t& any) {
  uint32_t f0 = clip_flags(p0, c), f1 = clip_flags(p1, c), f2 = clip_flags(p2, c);
  uint32_t common = f0 & f1 & f2;
  any = f0 | f1 | f2;
  if (common) {
    if (common & (kClipY0 | kClipY1)) return 0;
    EdgeEmitter<Out> em(out, c);
    em.border((common & kClipX0) != 0, tclamp(p0.y, c.y0, c.y1), tclamp(p2.y, c.y0, c.y1));
    return 0;
  }

  spline[0] = p0; spline[1] = p1; spline[2] = p2;

  // Geometry::split_with_options<kExtremaXY> (bezier_p.h:366-400).
  double ts[3]; int n = 0;
  {
    P2 ext = (p0 - p1) / (p0 - p1 * 2.0 + p2);
    double t0 = tmin(ext.x, ext.y);
    double t1 = tmax(ext.x, ext.y);
    if ((t0 > 0.0) & (t0 < 1.0)) ts[n++] = t0;
    if ((t1 > tmax(t0, 0.0)) & (t1 < 1.0)) ts[n++] = t1;
  }
  int pieces = 1;
  if (n) {
    ts[n++] = 1.0;
    P2 src[3] = { p0, p1, p2 };
    pieces = split_quad_at(src, spline, ts, n);
  }
  return pieces;
}

template<typename Out>
B2D_HD void build_quad(P2 p0, P2 p1, P2 p2, const ClipBox& c, double tol_sq, Out& out) {
  P2 spline[7];
  uint32_t any;
  const int pieces = prepare_quad(p0, p1, p2, c, out, spline, any);

  MonoCurve<3> mc;
  mc.tol_sq = tol_sq;
  for (int i = 0; i < pieces; i++) {
    const P2* piece = spline + i * 2;
    uint32_t sign_bit = piece[0].y > piece[2].y;
    if (any) flatten_unsafe<3>(mc, piece, sign_bit, c, out);
    else flatten_safe<3>(mc, piece, sign_bit, out);
  }
}

// Math::quad_roots (support/math_p.h:574-593).
B2D_HD int quad_roots(double* dst, double a, double b, double cc, double t_min, double t_max) {
  double d = tmax(b * b - 4.0 * a * cc, 0.0);
  double s = sqrt(d);
  double q = -0.5 * (b + copysign(s, b));
  double t0 = q / a;
  double t1 = cc / q;
  double x0 = tmin(t0, t1);
  double x1 = tmax(t1, t0);
  dst[0] = x0;
  int n = int((x0 >= t_min) & (x0 <= t_max));
  dst[n] = x1;
  n += int((x1 > x0) & (x1 >= t_min) & (x1 <= t_max));
  return n;
}

// cubic_to up to the monotone pieces (see prepare_quad); spline[0 .. 3 * pieces].
template<typename Out>
B2D_HD int prepare_cubic(P2 p0, P2 p1, P2 p2, P2 p3, const ClipBox& c, Out& out, P2* spline /* [25] */, uint32_t& any) {
  uint32_t f0 = clip_flags(p0, c), f1 = clip_flags(p1, c), f2 = clip_flags(p2, c), f3 = clip_flags(p3, c);
  uint32_t common = f0 & f1 & f2 & f3;
  any = f0 | f1 | f2 | f3;
  if (common) {
    if (common & (kClipY0 | kClipY1)) return 0;
    EdgeEmitter<Out> em(out, c);
    em.border((common & kClipX0) != 0, tclamp(p0.y, c.y0, c.y1), tclamp(p3.y, c.y0, c.y1));
    return 0;
  }

  spline[0] = p0; spline[1] = p1; spline[2] = p2; spline[3] = p3;
  int pieces = 1;

  // Geometry::split_cubic_to_spline<kExtremaXYInflectionsCusp> (bezier_p.h:834-920).
  {
    const double kAfter0 = 1e-40, kBefore1 = 0.999999999999999889;     // support/mathconst_p.h:31-32
    double ts[9]; int n = 0;

    P2 v1 = p1 - p0, v2 = p2 - p1, v3 = p3 - p2;
    P2 ca = v3 - v2 - v2 + v1;                                         // coefficients_of (:619-630)
    P2 cb = 3.0 * (v2 - v1);
    P2 cc = 3.0 * v1;
    P2 cd = p0;

    double q0 = cross2(cb, ca);
    double q1 = cross2(cc, ca);
    double q2 = cross2(cc, cb);

    double t_cusp = (q1 / q0) * -0.5;
    if ((t_cusp > 0.0) & (t_cusp < 1.0)) ts[n++] = t_cusp;
    n += quad_roots(ts + n, q0 * 6.0, q1 * 6.0, q2 * 2.0, kAfter0, kBefore1);

    P2 da = 3.0 * (v3 - v2 - v2 + v1);                                 // derivative_coefficients_of (:632-642)
    P2 db = 6.0 * (v2 - v1);
    P2 dc = 3.0 * v1;
    n += quad_roots(ts + n, da.x, db.x, dc.x, kAfter0, kBefore1);
    n += quad_roots(ts + n, da.y, db.y, dc.y, kAfter0, kBefore1);

    if (n) {
      for (int i = 1; i < n; i++) {                                    // insertion_sort
        double v = ts[i]; int j = i;
        while (j > 0 && ts[j - 1] > v) { ts[j] = ts[j - 1]; j--; }
        ts[j] = v;
      }
      ts[n++] = 1.0;

      P2* o = spline;
      o[0] = p0;
      P2 last = p3;
      pieces = 0;
      int i = 0;
      double t_cut = 0.0;
      do {
        double t_val = ts[i++];
        if (t_val == t_cut) continue;

        const double k1Div3 = 1.0 / 3.0;
        double dt = (t_val - t_cut) * k1Div3;
        P2 tp = ((ca * t_val + cb) * t_val + cc) * t_val + cd;
        if (i == n) tp = last;

        P2 cp1 = ((ca * (t_cut * 3.0) + cb * 2.0) * t_cut + cc) * dt;
        P2 cp2 = ((ca * (t_val * 3.0) + cb * 2.0) * t_val + cc) * dt;

        o[1] = o[0] + cp1;
        o[2] = tp - cp2;
        o[3] = tp;
        o += 3;
        pieces++;
        t_cut = t_val;
      } while (i != n);
      // `if (spline_end == spline_ptr) spline_end += 3` (:1850-1851): every t was skipped -> original curve.
      if (pieces == 0) { spline[1] = p1; spline[2] = p2; spline[3] = p3; pieces = 1; }
    }
  }
  return pieces;
}

template<typename Out>
B2D_HD void build_cubic(P2 p0, P2 p1, P2 p2, P2 p3, const ClipBox& c, double tol_sq, Out& out) {
  P2 spline[8 * 3 + 1];
  uint32_t any;
  const int pieces = prepare_cubic(p0, p1, p2, p3, c, out, spline, any);

  MonoCurve<4> mc;
  mc.tol_sq = tol_sq;
  for (int i = 0; i < pieces; i++) {
    const P2* piece = spline + i * 3;
    uint32_t sign_bit = piece[0].y > piece[3].y;
    if (any) flatten_unsafe<4>(mc, piece, sign_bit, c, out);
    else flatten_safe<4>(mc, piece, sign_bit, out);
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Subdivision-tree nodes (device edge builder).  A monotone piece is expanded breadth first into at most kNodeCap nodes,
// which are then walked independently (safe_walk / unsafe_walk).  The children of a node are exactly what the
// reference's push() leaves as "current curve" (first half) and "stack top" (second half).
// -----------------------------------------------------------------------------------------------------------------
enum : int { kNodeCap = 32 };
enum : uint32_t { kNodePendingMask = 0xFFu, kNodeSign = 0x100u, kNodeUnsafe = 0x200u, kNodeLtr = 0x400u };

// Root node of monotone piece `piece` (N points): what flatten_safe / flatten_unsafe set up before they start walking.
// Returns false when the piece needs no walk (clipped away, or emitted by the vertical-line case).
template<int N, typename Out>
B2D_HD bool piece_root(MonoCurve<N>& mc, const P2* piece, bool unsafe, const ClipBox& c, Out& out, uint32_t& meta) {
  const uint32_t sign_bit = piece[0].y > piece[N - 1].y;
  meta = sign_bit ? kNodeSign : 0u;
  if (unsafe) {
    bool ltr = true;
    if (!unsafe_prepare<N>(mc, piece, sign_bit, c, out, ltr)) return false;
    meta |= kNodeUnsafe | (ltr ? kNodeLtr : 0u);
  }
  else {
    mc.begin(piece, sign_bit);
    mc.bound(mc.first().x < mc.last().x);
  }
  return true;
}

// One breadth-first step on the node in mc.p: false = leaf (flat, or the reference's stack would be full), true =
// `first` / `second` receive the halves (pending + 1 / pending).
template<int N>
B2D_HD bool node_split(const MonoCurve<N>& mc, uint32_t pending, P2* first, P2* second) {
  typename MonoCurve<N>::Step st;
  if (mc.is_flat(st) || pending == uint32_t(kFlattenRecursionLimit)) return false;
  mc.split(st);
  if (N == 3) {
    first[0] = mc.p[0]; first[1] = st.a; first[2] = st.mid;
    second[0] = st.mid; second[1] = st.b; second[2] = mc.p[2];
  }
  else {
    first[0] = mc.p[0]; first[1] = st.a; first[2] = st.d; first[3] = st.mid;
    second[0] = st.mid; second[1] = st.e; second[2] = st.c; second[3] = mc.p[N - 1];
  }
  return true;
}

template<int N, typename Out>
B2D_HD void node_walk(MonoCurve<N>& mc, const P2* node, uint32_t meta, const ClipBox& c, Out& out) {
  mc.begin_at(node, meta & kNodePendingMask);
  const uint32_t sign_bit = (meta & kNodeSign) ? 1u : 0u;
  if (meta & kNodeUnsafe) unsafe_walk<N>(mc, (meta & kNodeLtr) != 0u, sign_bit, c, out);
  else safe_walk<N>(mc, sign_bit, out);
}

// -----------------------------------------------------------------------------------------------------------------
// Transform (EdgeTransformScale / EdgeTransformAffine, :67-95; BLMatrix2D::map_point).
// -----------------------------------------------------------------------------------------------------------------
struct GeomXform { double m00, m01, m10, m11, m20, m21; uint32_t affine; };

B2D_HD P2 xform(const GeomXform& t, P2 s) {
  if (!t.affine) return mk(s.x * t.m00 + t.m20, s.y * t.m11 + t.m21);
  return mk(s.x * t.m00 + s.y * t.m10 + t.m20, s.x * t.m01 + s.y * t.m11 + t.m21);
}

} // namespace b2d
