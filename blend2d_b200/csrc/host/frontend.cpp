// frontend.cpp - host-side mirror of the Blend2D raster-context front end for the rendering hot path.
//
// What this file restates from the reference (host-only code; it never touches pixels):
//   style -> FetchData           raster/rastercontext.cpp:363-494, pipeline/pipedefs.cpp:53-699
//   gradient LUTs                core/gradient.cpp:202-330, pixelops/interpolation.cpp, interpolation_avx2.cpp
//   comp-op simplification       core/compopsimplifyimpl_p.h:79-163, 423-437, 533-544, 571-583
//   alpha quantisation           raster/rastercontext.cpp:1920-1939, 2000-2029
//   rect / path dispatch         raster/rastercontext.cpp:853-920, 2634-2920, 3382-3458
//   BLMatrix2D                   core/matrix.cpp:139-330, 442-463
// Commands are queued exactly like the reference's asynchronous mode (rastercontext.cpp:2413-2523) and handed to the
// GPU pipeline runtime through b2dgpu_submit() on flush (rastercontext.cpp:1021-1072).
#include "../../../include/b2d_host.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <memory>
#include <vector>
#include <unordered_set>

namespace {

template<typename T> inline T bmin(T a, T b) { return b < a ? b : a; }
template<typename T> inline T bmax(T a, T b) { return a < b ? b : a; }
template<typename T> inline T bclamp(T a, T lo, T hi) { return bmin(hi, bmax(lo, a)); }

// ---------------------------------------------------------------------------------------------------------------
// Math helpers with the reference's rounding behaviour (support/math_p.h:329-420).
// ---------------------------------------------------------------------------------------------------------------
inline int round_to_int(double x) { int y = int(lrint(x)); return y + (double(y) - x == -0.5); }
inline int64_t floor_to_int64(double x) { int64_t y = int64_t(x); return y - int64_t(double(y) > x); }
inline int trunc_to_int(double x) { return int(x); }
inline bool is_finite_d(double x) { return isfinite(x); }

// ---------------------------------------------------------------------------------------------------------------
// BLMatrix2D
// ---------------------------------------------------------------------------------------------------------------
struct Matrix { double m00, m01, m10, m11, m20, m21; };
const Matrix kIdentity = { 1.0, 0.0, 0.0, 1.0, 0.0, 0.0 };

enum : uint32_t { kTTIdentity = 0, kTTTranslate = 1, kTTScale = 2, kTTSwap = 3, kTTAffine = 4, kTTInvalid = 5 };

uint32_t matrix_type(const Matrix& m) {                                      // bl_matrix2d_get_type
  uint32_t msk = (uint32_t(m.m00 != 0.0) << 3) | (uint32_t(m.m01 != 0.0) << 2) | (uint32_t(m.m10 != 0.0) << 1) | uint32_t(m.m11 != 0.0);
  const uint32_t valid = (1u << 3) | (1u << 6) | (1u << 7) | (1u << 9) | (1u << 11) | (1u << 12) | (1u << 13) | (1u << 14) | (1u << 15);
  double d = m.m00 * m.m11 - m.m01 * m.m10;
  if (!((1u << msk) & valid) || !is_finite_d(d) || !is_finite_d(m.m20) || !is_finite_d(m.m21)) return kTTInvalid;
  if (msk != 9u) return msk == 6u ? kTTSwap : kTTAffine;
  if (!((m.m00 == 1.0) & (m.m11 == 1.0))) return kTTScale;
  if (!((m.m20 == 0.0) & (m.m21 == 0.0))) return kTTTranslate;
  return kTTIdentity;
}

void matrix_multiply(Matrix& dst, const Matrix& a, const Matrix& b) {          // TransformInternal::multiply
  Matrix r;
  r.m00 = a.m00 * b.m00 + a.m01 * b.m10;
  r.m01 = a.m00 * b.m01 + a.m01 * b.m11;
  r.m10 = a.m10 * b.m00 + a.m11 * b.m10;
  r.m11 = a.m10 * b.m01 + a.m11 * b.m11;
  r.m20 = a.m20 * b.m00 + a.m21 * b.m10 + b.m20;
  r.m21 = a.m20 * b.m01 + a.m21 * b.m11 + b.m21;
  dst = r;
}

bool matrix_invert(Matrix& dst, const Matrix& s) {                             // bl_matrix2d_invert
  double d = s.m00 * s.m11 - s.m01 * s.m10;
  if (d == 0.0 || !is_finite_d(d)) return false;
  double t00 = s.m11, t01 = -s.m01, t10 = -s.m10, t11 = s.m00;
  t00 /= d; t01 /= d; t10 /= d; t11 /= d;
  double t20 = -(s.m20 * t00 + s.m21 * t10);
  double t21 = -(s.m20 * t01 + s.m21 * t11);
  dst = Matrix{ t00, t01, t10, t11, t20, t21 };
  return true;
}

inline void map_point(const Matrix& m, double x, double y, double& ox, double& oy) {
  ox = x * m.m00 + y * m.m10 + m.m20;
  oy = x * m.m01 + y * m.m11 + m.m21;
}

bool matrix_apply_op(Matrix& a, uint32_t op, const double* data) {            // bl_matrix2d_apply_op
  switch (op) {
    case 0: a = kIdentity; return true;
    case 1: memcpy(&a, data, sizeof(Matrix)); return true;
    case 2: { double x = data[0], y = data[1]; a.m20 += x * a.m00 + y * a.m10; a.m21 += x * a.m01 + y * a.m11; return true; }
    case 3: { double x = data[0], y = data[1]; a.m00 *= x; a.m01 *= x; a.m10 *= y; a.m11 *= y; return true; }
    case 4: {
      double xt = tan(data[0]), yt = tan(data[1]);
      double t00 = yt * a.m10, t01 = yt * a.m11;
      a.m10 += xt * a.m00; a.m11 += xt * a.m01; a.m00 += t00; a.m01 += t01;
      return true;
    }
    case 5: case 6: {
      double angle = data[0], as = sin(angle), ac = cos(angle);
      double t00 = as * a.m10 + ac * a.m00, t01 = as * a.m11 + ac * a.m01;
      double t10 = ac * a.m10 - as * a.m00, t11 = ac * a.m11 - as * a.m01;
      if (op == 6) {
        double px = data[1], py = data[2];
        double tx = px - ac * px + as * py, ty = py - as * px - ac * py;
        double t20 = tx * a.m00 + ty * a.m10 + a.m20, t21 = tx * a.m01 + ty * a.m11 + a.m21;
        a.m20 = t20; a.m21 = t21;
      }
      a.m00 = t00; a.m01 = t01; a.m10 = t10; a.m11 = t11;
      return true;
    }
    case 7: { Matrix b; memcpy(&b, data, sizeof(Matrix)); matrix_multiply(a, b, a); return true; }
    case 8: a.m20 += data[0]; a.m21 += data[1]; return true;
    case 9: { double x = data[0], y = data[1]; a.m00 *= x; a.m01 *= y; a.m10 *= x; a.m11 *= y; a.m20 *= x; a.m21 *= y; return true; }
    case 13: { Matrix b; memcpy(&b, data, sizeof(Matrix)); matrix_multiply(a, a, b); return true; }
    default: return false;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pixel helpers
// ---------------------------------------------------------------------------------------------------------------
inline uint32_t div255(uint32_t x) { return ((x + 128u) * 257u) >> 16; }

uint32_t premultiply_argb32(uint32_t v) {                                     // cvt_prgb32_8888_from_argb32_8888
  uint32_t a = v >> 24;
  uint32_t r = div255(((v >> 16) & 0xFFu) * a), g = div255(((v >> 8) & 0xFFu) * a), b = div255((v & 0xFFu) * a);
  return (a << 24) | (r << 16) | (g << 8) | b;
}

inline uint32_t udiv65535(uint32_t x) { return ((x + 0x8000u) + ((x + 0x8000u) >> 16)) >> 16; }

uint64_t premultiply_argb64(uint64_t v) {                                     // cvt_prgb64_8888_from_argb64_8888
  uint32_t a = uint32_t(v >> 48);
  uint64_t r = udiv65535(uint32_t((v >> 32) & 0xFFFFu) * a), g = udiv65535(uint32_t((v >> 16) & 0xFFFFu) * a), b = udiv65535(uint32_t(v & 0xFFFFu) * a);
  return (uint64_t(a) << 48) | (r << 32) | (g << 16) | b;
}

inline uint32_t rgba32_from_rgba64(uint64_t v) {                               // RgbaInternal::rgba32FromRgba64
  return (uint32_t(v >> 56) << 24) | (uint32_t((v >> 40) & 0xFFu) << 16) | (uint32_t((v >> 24) & 0xFFu) << 8) | uint32_t((v >> 8) & 0xFFu);
}

inline uint32_t format_from_rgba32(uint32_t rgba32) {                          // formatFromRgba32
  return rgba32 == 0u ? B2DGPU_FORMAT_ZERO32 : rgba32 >= 0xFF000000u ? B2DGPU_FORMAT_FRGB32 : B2DGPU_FORMAT_PRGB32;
}

// ---------------------------------------------------------------------------------------------------------------
// Comp-op simplification (core/compopsimplifyimpl_p.h).  Result: simplified (op, dst, src) or "nop", plus the solid
// override the reference applies (transparent / opaque black / opaque white).
// ---------------------------------------------------------------------------------------------------------------
enum SolidId : uint32_t { kSolidNone = 0, kSolidTransparent = 1, kSolidOpaqueBlack = 2, kSolidOpaqueWhite = 3, kSolidNop = 4 };
enum : uint32_t { kOpSrcOver = 0, kOpSrcCopy = 1, kOpSrcIn = 2, kOpSrcOut = 3, kOpSrcAtop = 4, kOpDstOver = 5, kOpDstCopy = 6, kOpDstIn = 7,
                  kOpDstOut = 8, kOpDstAtop = 9, kOpXor = 10, kOpClear = 11, kOpPlus = 12, kOpMinus = 13, kOpModulate = 14, kOpMultiply = 15,
                  kOpScreen = 16, kOpOverlay = 17, kOpDarken = 18, kOpLighten = 19, kOpColorDodge = 20, kOpColorBurn = 21, kOpLinearBurn = 22,
                  kOpLinearLight = 23, kOpPinLight = 24, kOpHardLight = 25, kOpSoftLight = 26, kOpDifference = 27, kOpExclusion = 28 };

struct Simplified { uint32_t op, dst, src, solid; bool implemented; };

const uint32_t P = B2DGPU_FORMAT_PRGB32, X = B2DGPU_FORMAT_XRGB32, A = B2DGPU_FORMAT_A8, F = B2DGPU_FORMAT_FRGB32, Z = B2DGPU_FORMAT_ZERO32;

Simplified mk_op(uint32_t op, uint32_t d, uint32_t s, uint32_t solid = kSolidNone) { return Simplified{ op, d, s, solid, true }; }
Simplified mk_nop() { return Simplified{ kOpDstCopy, 0, 0, kSolidNop, true }; }
Simplified mk_unimpl(uint32_t op, uint32_t d, uint32_t s) { return Simplified{ op, d, s, kSolidNone, false }; }

Simplified s_clear(uint32_t d, uint32_t s) {
  if (d == P) return mk_op(kOpSrcCopy, P, P, kSolidTransparent);
  if (d == X) return mk_op(kOpSrcCopy, P, P, kSolidOpaqueBlack);
  if (d == A) return mk_op(kOpSrcCopy, A, P, kSolidTransparent);
  return mk_unimpl(kOpClear, d, s);
}

Simplified s_src_copy(uint32_t d, uint32_t s) {
  if (d == P && (s == Z || s == F)) return mk_op(kOpSrcCopy, P, P);
  if (d == X && (s == P || s == Z || s == X)) return mk_op(kOpSrcCopy, P, X);
  if (d == X && s == F) return mk_op(kOpSrcCopy, P, P);
  if (d == A && s == Z) return s_clear(A, Z);
  if (d == A && (s == X || s == F)) return mk_op(kOpSrcCopy, A, P, kSolidOpaqueWhite);
  return mk_op(kOpSrcCopy, d, s);
}

Simplified s_src_over(uint32_t d, uint32_t s) {
  if (d == P && s == Z) return mk_nop();
  if (d == P && s == X) return s_src_copy(P, X);
  if (d == P && s == F) return s_src_copy(P, F);
  if (d == X && s == P) return s_src_over(P, P);
  if (d == X && s == Z) return mk_nop();
  if (d == X && s == X) return s_src_copy(P, X);
  if (d == X && s == F) return s_src_copy(P, F);
  if (d == A && s == Z) return mk_nop();
  if (d == A && s == X) return s_src_copy(A, X);
  if (d == A && s == F) return s_src_copy(A, F);
  return mk_op(kOpSrcOver, d, s);
}

Simplified s_plus(uint32_t d, uint32_t s) {
  if (d == P && s == Z) return mk_nop();
  if (d == P && s == F) return s_plus(P, P);
  if (d == X && (s == P || s == X || s == F)) return s_plus(P, P);
  if (d == X && s == Z) return mk_nop();
  if (d == A && s == Z) return mk_nop();
  if (d == A && (s == X || s == F)) return mk_op(kOpPlus, A, P, kSolidOpaqueWhite);
  return mk_op(kOpPlus, d, s);
}

Simplified s_multiply(uint32_t d, uint32_t s) {
  if (d == P && s == Z) return mk_nop();
  if (d == P && s == F) return s_multiply(P, X);
  if (d == X && s == Z) return mk_nop();
  if (d == X && (s == F || s == X)) return mk_unimpl(kOpModulate, X, X);
  if (d == A || s == A) return mk_unimpl(kOpDstOver, d, s);
  if (d == X) return mk_unimpl(kOpMultiply, d, s);           // XRGB32 x PRGB32 keeps a dedicated JIT variant
  return mk_op(kOpMultiply, d, s);
}

Simplified s_screen(uint32_t d, uint32_t s) {
  if (d == P && s == Z) return mk_nop();
  if (d == P && s == F) return s_screen(P, P);
  if (d == X && (s == P || s == F)) return s_screen(P, P);
  if (d == X && s == Z) return mk_nop();
  if (d == X && s == X) return s_screen(P, X);
  if (d == A || s == A) return s_src_over(d, s);
  return mk_op(kOpScreen, d, s);
}

Simplified simplify(uint32_t op, uint32_t d, uint32_t s) {
  switch (op) {
    case kOpSrcOver: return s_src_over(d, s);
    case kOpSrcCopy: return s_src_copy(d, s);
    case kOpClear: return s_clear(d, s);
    case kOpPlus: return s_plus(d, s);
    case kOpMultiply: return s_multiply(d, s);
    case kOpScreen: return s_screen(d, s);
    // The other operators of the GPU runtime (dev_pixel.cuh comp_jit_ext): this mirror only knows the case the
    // reference leaves untouched, PRGB32 x PRGB32 (core/compopsimplifyimpl_p.h: every [Op PRGBxPRGB] entry is the operator
    // itself); sources without alpha are rewritten there (SrcAtop -> SrcIn, Xor -> SrcOut, ...) - applications get
    // that through the real frontend (shim/).
    case kOpSrcIn: case kOpSrcOut: case kOpSrcAtop: case kOpDstOver: case kOpDstIn: case kOpDstOut: case kOpDstAtop: case kOpXor:
    case kOpMinus: case kOpModulate: case kOpDarken: case kOpLighten: case kOpLinearBurn: case kOpDifference: case kOpExclusion:
    case kOpOverlay: case kOpColorDodge: case kOpColorBurn: case kOpLinearLight: case kOpPinLight: case kOpHardLight: case kOpSoftLight:
      return d == P && s == P ? mk_op(op, d, s) : mk_unimpl(op, d, s);
    default: return mk_unimpl(op, d, s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient LUTs
// ---------------------------------------------------------------------------------------------------------------
struct Stop { double offset; uint64_t rgba; };

inline uint32_t premul_top8(uint64_t c) {
  uint32_t a = uint32_t(c >> 56), r = uint32_t((c >> 40) & 0xFFu), g = uint32_t((c >> 24) & 0xFFu), b = uint32_t((c >> 8) & 0xFFu);
  return (a << 24) | (div255(r * a) << 16) | (div255(g * a) << 8) | div255(b * a);
}

// interpolate_prgb32 as executed on AVX2 hosts (pixelops/interpolation_avx2.cpp:18-205) in scalar form: 8-bit channel
// positions in 9.23 fixed point, per-channel step = trunc((c1 - c0) * (2^23 / n)) in double, premultiply after.
// The inner loop is pure 32-bit integer work; the AVX2 clone (picked at load time) vectorises it ~4x, which matters:
// bl_bench-style scenes create one gradient - hence one table - per fill.
__attribute__((target_clones("avx2", "default"), optimize("O3")))
void make_lut32(uint32_t* d, uint32_t size, const Stop* stops, size_t n) {
  uint64_t c0 = stops[0].rgba, c1 = c0;
  uint32_t u0 = 0, u1;
  size_t si = size_t(stops[0].offset == 0.0 && n > 1);
  uint32_t d_size = size - 1;
  double f_width = double(int32_t(d_size) << 8);
  uint32_t* dp = d;

  do {
    c1 = stops[si].rgba;
    u1 = uint32_t(round_to_int(stops[si].offset * f_width));
    dp = d + (u0 >> 8);
    uint32_t i = (u1 >> 8) - (u0 >> 8);
    u0 = u1;

    if (i <= 1) {
      uint32_t p0 = premul_top8(c0), p1 = premul_top8(c1);
      c0 = c1;
      *dp++ = p0;
      if (i == 0) continue;
      *dp++ = p1;
    }
    else {
      uint32_t cnt = i + 1;
      double scale = double(1 << 23) / double(int(i));
      uint32_t cx[4]; int32_t dx[4];
      for (int ch = 0; ch < 4; ch++) {
        int sh = 8 + ch * 16;                                     // top byte of each 16-bit channel: b, g, r, a
        int32_t a0 = int32_t((c0 >> sh) & 0xFFu), a1 = int32_t((c1 >> sh) & 0xFFu);
        dx[ch] = int32_t(double(a1 - a0) * scale);                // cvttpd2dq
        cx[ch] = (uint32_t(a0) << 23) + (1u << 22);
      }
      for (uint32_t k = 0; k < cnt; k++) {
        uint32_t b = cx[0] >> 23, g = cx[1] >> 23, r = cx[2] >> 23, a = cx[3] >> 23;
        for (int ch = 0; ch < 4; ch++) cx[ch] += uint32_t(dx[ch]);
        *dp++ = (a << 24) | (div255(r * a) << 16) | (div255(g * a) << 8) | div255(b * a);
      }
      c0 = c1;
    }
  } while (++si < n);

  uint32_t rest = uint32_t((d + d_size + 1) - dp);
  uint32_t last = premul_top8(c0);
  while (rest--) *dp++ = last;
  d[0] = premul_top8(stops[0].rgba);
}

// interpolate_prgb64 (pixelops/interpolation.cpp:163-296; no SIMD variant exists).
void make_lut64(uint64_t* d_ptr, uint32_t d_size, const Stop* s_ptr, size_t s_size) {
  uint64_t* dp = d_ptr;
  uint32_t i = d_size;
  uint64_t c0 = s_ptr[0].rgba, c1 = c0;
  uint32_t p0 = 0, p1;
  size_t si = 0;
  double f_width = double(int32_t(--d_size) << 8);
  uint64_t cp = premultiply_argb64(c0);
  uint64_t cp_first = cp;
  bool solid_only = (s_size == 1);

  auto span = [&](void) {
    // Writes `i` entries starting at dp for the transition c0 -> c1 (or a solid run when they are equal).
    cp = premultiply_argb64(c0);
    if (c0 == c1) { do { *dp++ = cp; } while (--i); return; }
    *dp++ = cp;
    if (--i) {
      const uint32_t kShift = 15, kMask = 0xFFFFu << kShift;
      uint32_t r_pos = uint32_t((c0 >> (32 - kShift)) & kMask), g_pos = uint32_t((c0 >> (16 - kShift)) & kMask), b_pos = uint32_t((c0 << kShift) & kMask);
      uint32_t r_inc = uint32_t((c1 >> (32 - kShift)) & kMask), g_inc = uint32_t((c1 >> (16 - kShift)) & kMask), b_inc = uint32_t((c1 << kShift) & kMask);
      r_inc = uint32_t(int32_t(r_inc - r_pos) / int32_t(i));
      g_inc = uint32_t(int32_t(g_inc - g_pos) / int32_t(i));
      b_inc = uint32_t(int32_t(b_inc - b_pos) / int32_t(i));
      r_pos += 1u << (kShift - 1); g_pos += 1u << (kShift - 1); b_pos += 1u << (kShift - 1);
      if (((c0 & c1) & 0xFFFF000000000000ull) == 0xFFFF000000000000ull) {
        do {
          r_pos += r_inc; g_pos += g_inc; b_pos += b_inc;
          *dp++ = (uint64_t(r_pos & kMask) << (32 - kShift)) | (uint64_t(g_pos & kMask) << (16 - kShift)) | (uint64_t(b_pos & kMask) >> kShift) | 0xFFFF000000000000ull;
        } while (--i);
      }
      else {
        uint32_t a_pos = uint32_t((c0 >> (48 - kShift)) & kMask), a_inc = uint32_t((c1 >> (48 - kShift)) & kMask);
        a_inc = uint32_t(int32_t(a_inc - a_pos) / int32_t(i));
        a_pos += 1u << (kShift - 1);
        do {
          a_pos += a_inc; r_pos += r_inc; g_pos += g_inc; b_pos += b_inc;
          uint32_t ca = a_pos >> kShift;
          uint64_t cr = udiv65535((r_pos >> kShift) * ca), cg = udiv65535((g_pos >> kShift) * ca), cb = udiv65535((b_pos >> kShift) * ca);
          *dp++ = (uint64_t(ca) << 48) | (cr << 32) | (cg << 16) | cb;
        } while (--i);
      }
    }
    c0 = c1;
  };

  if (solid_only) { do { *dp++ = cp; } while (--i); }
  else {
    do {
      c1 = s_ptr[si].rgba;
      p1 = uint32_t(round_to_int(s_ptr[si].offset * f_width));
      dp = d_ptr + (p0 >> 8);
      i = (p1 >> 8) - (p0 >> 8);
      if (i == 0) c0 = c1;
      p0 = p1;
      i++;
      span();
    } while (++si < s_size);
    i = uint32_t((d_ptr + d_size + 1) - dp);
    if (i != 0) { c1 = c0; span(); }
  }
  d_ptr[0] = cp_first;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// Objects
// ---------------------------------------------------------------------------------------------------------------
struct b2d_image {
  int refs;
  int32_t w, h;
  uint32_t format;
  int bpp;
  intptr_t stride;
  uint8_t* data;
};

struct b2d_gradient {
  int refs;
  uint32_t type, extend_mode;
  double values[6];
  Matrix transform;
  uint32_t transform_type;
  std::vector<Stop> stops;
  // BLGradientInfo (core/gradient.cpp:202-281)
  bool empty, solid;
  uint32_t format, lut_size;
  std::unique_ptr<uint32_t[]> lut32;            // tables are written in full by make_lut*: no zero fill
  std::unique_ptr<uint64_t[]> lut64;
};

struct b2d_pattern {
  int refs;
  b2d_image* image;
  int32_t area[4];
  uint32_t extend_mode;
  Matrix transform;
  uint32_t transform_type;
};

static int image_bpp(uint32_t format) { return format == B2DGPU_FORMAT_A8 ? 1 : (format == B2DGPU_FORMAT_PRGB32 || format == B2DGPU_FORMAT_XRGB32) ? 4 : 0; }

extern "C" b2dgpu_result b2d_image_create(int32_t w, int32_t h, uint32_t format, b2d_image** out) {
  if (!out) return B2DGPU_ERROR_INVALID_VALUE;
  *out = nullptr;
  int bpp = image_bpp(format);
  if (!bpp || w <= 0 || h <= 0 || w > 65535 || h > 65535) return B2DGPU_ERROR_INVALID_VALUE;
  b2d_image* img = new (std::nothrow) b2d_image();
  if (!img) return B2DGPU_ERROR_OUT_OF_MEMORY;
  img->refs = 1;
  img->w = w; img->h = h; img->format = format; img->bpp = bpp; img->stride = intptr_t(w) * bpp;
  void* p = nullptr;
  if (posix_memalign(&p, 64, size_t(img->stride) * size_t(h) + 64) != 0) { delete img; return B2DGPU_ERROR_OUT_OF_MEMORY; }
  memset(p, 0, size_t(img->stride) * size_t(h));
  img->data = static_cast<uint8_t*>(p);
  *out = img;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_image_destroy(b2d_image* img) {
  if (!img) return B2DGPU_ERROR_INVALID_VALUE;
  if (--img->refs > 0) return B2DGPU_SUCCESS;         // still retained by a pattern / queued command
  free(img->data);
  delete img;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_image_get_data(b2d_image* img, b2dgpu_image_data* out) {
  if (!img || !out) return B2DGPU_ERROR_INVALID_VALUE;
  out->pixel_data = img->data; out->stride = img->stride; out->w = img->w; out->h = img->h; out->format = img->format; out->flags = 0;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_gradient_create(uint32_t type, const double* values, uint32_t extend_mode, const b2d_gradient_stop* stops,
                                             uint32_t stop_count, const double* matrix, b2d_gradient** out) {
  if (!out || !values || type > B2D_GRADIENT_CONIC || extend_mode > B2D_EXTEND_REFLECT) return B2DGPU_ERROR_INVALID_VALUE;
  *out = nullptr;
  b2d_gradient* g = new (std::nothrow) b2d_gradient();
  if (!g) return B2DGPU_ERROR_OUT_OF_MEMORY;
  g->refs = 1;
  g->type = type; g->extend_mode = extend_mode;
  size_t nv = type == B2D_GRADIENT_RADIAL ? 6 : 4;
  memset(g->values, 0, sizeof(g->values));
  memcpy(g->values, values, nv * sizeof(double));
  if (matrix) memcpy(&g->transform, matrix, sizeof(Matrix)); else g->transform = kIdentity;
  g->transform_type = matrix ? matrix_type(g->transform) : kTTIdentity;
  g->stops.reserve(stop_count);
  for (uint32_t i = 0; i < stop_count; i++) {
    if (i && stops[i].offset < stops[i - 1].offset) { delete g; return B2DGPU_ERROR_INVALID_VALUE; }
    g->stops.push_back(Stop{ bclamp(stops[i].offset, 0.0, 1.0), stops[i].rgba64 });
  }

  // ensure_info()
  g->empty = g->stops.empty();
  g->solid = false; g->format = P; g->lut_size = 0;
  if (!g->empty) {
    const uint32_t kAlphaNotOne = 1, kAlphaNotZero = 2, kTransition = 4;
    uint32_t flags = 0;
    uint64_t prev = g->stops[0].rgba & 0xFF00FF00FF00FF00ull;
    if (prev < 0xFF00000000000000ull) flags |= kAlphaNotOne;
    if (prev > 0x00FFFFFFFFFFFFFFull) flags |= kAlphaNotZero;
    for (size_t i = 1; i < g->stops.size(); i++) {
      uint64_t v = g->stops[i].rgba & 0xFF00FF00FF00FF00ull;
      if (v == prev) continue;
      flags |= kTransition;
      if (v < 0xFF00000000000000ull) flags |= kAlphaNotOne;
      if (v > 0x00FFFFFFFFFFFFFFull) flags |= kAlphaNotZero;
      prev = v;
    }
    if (!(flags & kAlphaNotZero)) flags &= ~kTransition;
    uint32_t lut_size = 256;
    if (flags & kTransition) {
      size_t n = g->stops.size();
      if (n == 1) lut_size = 256;
      else if (n == 2) lut_size = (g->stops[1].offset - g->stops[0].offset >= 0.998) ? 256 : 512;
      else if (n == 3) lut_size = (g->stops[0].offset <= 0.002 && g->stops[1].offset == 0.5 && g->stops[2].offset >= 0.998) ? 512 : 1024;
      else lut_size = 1024;
    }
    g->solid = !(flags & kTransition);
    g->format = (flags & kAlphaNotOne) ? P : F;
    g->lut_size = lut_size;
  }
  *out = g;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_gradient_destroy(b2d_gradient* g) {
  if (!g) return B2DGPU_ERROR_INVALID_VALUE;
  if (--g->refs > 0) return B2DGPU_SUCCESS;
  delete g;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_pattern_create(b2d_image* image, const int32_t* area, uint32_t extend_mode, const double* matrix, b2d_pattern** out) {
  if (!out || !image || extend_mode > B2D_EXTEND_REFLECT_X_REPEAT_Y) return B2DGPU_ERROR_INVALID_VALUE;
  *out = nullptr;
  b2d_pattern* p = new (std::nothrow) b2d_pattern();
  if (!p) return B2DGPU_ERROR_OUT_OF_MEMORY;
  p->refs = 1;
  p->image = image;
  image->refs++;
  if (area) {
    memcpy(p->area, area, sizeof(p->area));
    if (p->area[0] < 0 || p->area[1] < 0 || p->area[2] < 0 || p->area[3] < 0 || p->area[0] + p->area[2] > image->w || p->area[1] + p->area[3] > image->h) { image->refs--; delete p; return B2DGPU_ERROR_INVALID_VALUE; }
  }
  else { p->area[0] = 0; p->area[1] = 0; p->area[2] = image->w; p->area[3] = image->h; }
  p->extend_mode = extend_mode;
  if (matrix) memcpy(&p->transform, matrix, sizeof(Matrix)); else p->transform = kIdentity;
  p->transform_type = matrix ? matrix_type(p->transform) : kTTIdentity;
  *out = p;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_pattern_destroy(b2d_pattern* p) {
  if (!p) return B2DGPU_ERROR_INVALID_VALUE;
  if (--p->refs > 0) return B2DGPU_SUCCESS;
  b2d_image_destroy(p->image);
  delete p;
  return B2DGPU_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------
// FetchData initialisers (pipeline/pipedefs.cpp)
// ---------------------------------------------------------------------------------------------------------------
namespace {

inline uint32_t extend_x_of(uint32_t m) { static const uint8_t t[9] = { 0, 1, 2, 0, 0, 1, 1, 2, 2 }; return t[m]; }
inline uint32_t extend_y_of(uint32_t m) { static const uint8_t t[9] = { 0, 1, 2, 1, 2, 0, 2, 0, 1 }; return t[m]; }

const uint32_t kPending = 0xFFFFFFFFu;      // "signature has the pending flag": the style cannot be rendered

uint32_t init_pattern_tx_ty(b2dgpu_fetch_pattern& fd, uint32_t fetch_base, uint32_t extend_mode, int tx, int ty, bool is_fractional) {
  uint32_t ex = extend_x_of(extend_mode), ey = extend_y_of(extend_mode);
  int rx = 0, ry = 0;
  if (fd.src.w <= 1) ex = B2D_EXTEND_PAD;
  if (fd.src.h <= 1) ey = B2D_EXTEND_PAD;

  if (ex >= B2D_EXTEND_REPEAT) {
    bool is_reflect = ex == B2D_EXTEND_REFLECT;
    rx = int(fd.src.w) << uint32_t(is_reflect);
    if (unsigned(tx) >= unsigned(rx)) tx %= rx;
    if (tx < 0) tx += rx;
    if (is_fractional) ex = 1;
  }

  b2dgpu_vert_extend& ve = fd.simple.v_extend;
  ve.stride[0] = fd.src.stride; ve.stride[1] = 0;
  ve.y_stop[0] = uint32_t(fd.src.h); ve.y_stop[1] = 0;
  ve.y_rewind_offset = 0;
  ve.pixel_ptr_rewind_offset = (ey != B2D_EXTEND_REPEAT ? intptr_t(0) : intptr_t(fd.src.h - 1)) * fd.src.stride;

  if (ey >= B2D_EXTEND_REPEAT) {
    ry = int(fd.src.h) << uint32_t(ey == B2D_EXTEND_REFLECT);
    if (unsigned(ty) >= unsigned(ry)) ty %= ry;
    if (ty < 0) ty += ry;
    ve.stride[1] = (ey == B2D_EXTEND_REPEAT) ? fd.src.stride : -fd.src.stride;
    ve.y_stop[1] = uint32_t(fd.src.h);
    ve.y_rewind_offset = uint32_t(fd.src.h);
  }

  fd.simple.tx = tx; fd.simple.ty = ty; fd.simple.rx = rx; fd.simple.ry = ry;
  memset(fd.simple.ix, 0, sizeof(fd.simple.ix));
  return fetch_base + ex;
}

uint32_t init_pattern_fx_fy(b2dgpu_fetch_pattern& fd, uint32_t extend_mode, uint32_t quality, int64_t tx64, int64_t ty64) {
  uint32_t fetch_base = B2DGPU_FETCH_PATTERN_ALIGNED_PAD;
  uint32_t wx = uint32_t(tx64 & 0xFF), wy = uint32_t(ty64 & 0xFF);
  int tx = -int(tx64 >> 8), ty = -int(ty64 >> 8);
  bool is_fractional = (wx | wy) != 0;
  if (is_fractional) {
    if (quality == B2D_PATTERN_QUALITY_NEAREST) {
      tx -= (wx >= 128); ty -= (wy >= 128);
      is_fractional = false;
    }
    else {
      fd.simple.wa = ((wy) * (wx)) >> 8;
      fd.simple.wb = ((wy) * (256 - wx) + 255) >> 8;
      fd.simple.wc = ((256 - wy) * (wx)) >> 8;
      fd.simple.wd = ((256 - wy) * (256 - wx) + 255) >> 8;
      tx--; ty--;
      if (wy == 0) fetch_base = B2DGPU_FETCH_PATTERN_FX_PAD;
      else if (wx == 0) fetch_base = B2DGPU_FETCH_PATTERN_FY_PAD;
      else fetch_base = B2DGPU_FETCH_PATTERN_FXFY_PAD;
    }
  }
  return init_pattern_tx_ty(fd, fetch_base, extend_mode, tx, ty, is_fractional);
}

inline bool near_one(double x) { double d = x - 1.0; return (d < 0 ? -d : d) <= 1e-14; }
inline bool near_zero(double x) { return (x < 0 ? -x : x) <= 1e-14; }

uint32_t init_pattern_affine(b2dgpu_fetch_pattern& fd, uint32_t extend_mode, uint32_t quality, uint32_t bytes_per_pixel, const Matrix& transform) {
  Matrix inv;
  if (!matrix_invert(inv, transform)) return kPending;
  int tw = int(fd.src.w), th = int(fd.src.h);
  if (tw == 0) return kPending;

  double xx = inv.m00, xy = inv.m01, yx = inv.m10, yy = inv.m11;
  if (near_one(xx) & near_zero(xy) & near_zero(yx) & near_one(yy)) {
    int64_t tx64 = floor_to_int64(-inv.m20 * 256.0);
    int64_t ty64 = floor_to_int64(-inv.m21 * 256.0);
    return init_pattern_fx_fy(fd, extend_mode, quality, tx64, ty64);
  }

  uint32_t fetch_type = quality == B2D_PATTERN_QUALITY_NEAREST ? B2DGPU_FETCH_PATTERN_AFFINE_NN_ANY : B2DGPU_FETCH_PATTERN_AFFINE_BI_ANY;
  uint32_t opt = bmax(tw, th) < 32767 && fd.src.stride >= 0 && fd.src.stride <= intptr_t(32767);
  if (quality == B2D_PATTERN_QUALITY_BILINEAR) opt = 0;
  fetch_type += opt;

  uint32_t ex = extend_x_of(extend_mode), ey = extend_y_of(extend_mode);
  double tx = inv.m20, ty = inv.m21;
  tx += 0.5 * (xx + yx);
  ty += 0.5 * (xy + yy);

  const double fp_scale = 4294967296.0;
  int ox = INT32_MAX, oy = INT32_MAX, rx = 0, ry = 0;
  b2dgpu_pattern_affine& a = fd.affine;
  a.min_x = 0; a.min_y = 0;
  a.max_x = tw - 1; a.max_y = th - 1;
  a.cor_x = tw - 1; a.cor_y = th - 1;

  if (ex != B2D_EXTEND_PAD) {
    a.min_x = INT32_MIN;
    if (ex == B2D_EXTEND_REPEAT) a.cor_x = 0;
    ox = tw;
    if (ex == B2D_EXTEND_REFLECT) tw *= 2;
    if (xx < 0.0) {
      xx = -xx; yx = -yx; tx = double(tw) - tx;
      if (ex == B2D_EXTEND_REPEAT) { ox = 0; a.cor_x = a.max_x; a.max_x = -1; }
    }
    ox--;
  }
  if (ey != B2D_EXTEND_PAD) {
    a.min_y = INT32_MIN;
    if (ey == B2D_EXTEND_REPEAT) a.cor_y = 0;
    oy = th;
    if (ey == B2D_EXTEND_REFLECT) th *= 2;
    if (xy < 0.0) {
      xy = -xy; yy = -yy; ty = double(th) - ty;
      if (ey == B2D_EXTEND_REPEAT) { oy = 0; a.cor_y = a.max_y; a.max_y = -1; }
    }
    oy--;
  }

  if (quality != B2D_PATTERN_QUALITY_NEAREST) { tx -= 0.5; ty -= 0.5; }

  double tw_d = double(tw), th_d = double(th);
  if (ex == B2D_EXTEND_PAD) tw_d = 2147483647.0;
  else { tx = fmod(tx, tw_d); rx = tw; if (xx >= tw_d) xx = fmod(xx, tw_d); }
  if (ey == B2D_EXTEND_PAD) th_d = 2147483647.0;
  else { ty = fmod(ty, th_d); ry = th; if (xy >= th_d) xy = fmod(xy, th_d); }

  xx *= fp_scale; xy *= fp_scale; yx *= fp_scale; yy *= fp_scale; tx *= fp_scale; ty *= fp_scale;

  double all_min = bmin(bmin(yy, yx), bmin(xy, xx));
  double all_max = bmax(bmax(xx, xy), bmax(yx, yy));
  all_min = bmin(all_min, bmin(ty, tx));
  all_max = bmax(bmax(tx, ty), all_max);

  if (all_min >= double(INT64_MIN + 1) && all_max <= double(INT64_MAX)) {
    a.xx.i64 = floor_to_int64(xx); a.xy.i64 = floor_to_int64(xy);
    a.yx.i64 = floor_to_int64(yx); a.yy.i64 = floor_to_int64(yy);
    a.tx.i64 = floor_to_int64(tx); a.ty.i64 = floor_to_int64(ty);
  }
  else {
    a.xx.i64 = a.xy.i64 = a.yx.i64 = a.yy.i64 = a.tx.i64 = a.ty.i64 = 0;
  }

  a.rx.i64 = int64_t(uint64_t(int64_t(rx)) << 32);
  a.ry.i64 = int64_t(uint64_t(int64_t(ry)) << 32);
  a.ox.i32[1] = ox; a.ox.i32[0] = INT32_MAX;
  a.oy.i32[1] = oy; a.oy.i32[0] = INT32_MAX;
  a.tw = tw_d; a.th = th_d;
  a.xx2.u64 = a.xx.u64 << 1;
  a.xy2.u64 = a.xy.u64 << 1;
  if (ex >= B2D_EXTEND_REPEAT && a.xx2.u32[1] >= uint32_t(tw)) a.xx2.u32[1] %= uint32_t(tw);
  if (ey >= B2D_EXTEND_REPEAT && a.xy2.u32[1] >= uint32_t(th)) a.xy2.u32[1] %= uint32_t(th);
  a.addr_mul32[0] = int32_t(bytes_per_pixel);
  a.addr_mul32[1] = int32_t(fd.src.stride);
  return fetch_type;
}

uint32_t init_linear_gradient(b2dgpu_fetch_gradient& fd, const double* v, uint32_t extend_mode, uint32_t quality, const Matrix& transform) {
  Matrix inv;
  if (!matrix_invert(inv, transform)) return kPending;
  double p0x = v[0], p0y = v[1], p1x = v[2], p1y = v[3];
  uint32_t lut_size = fd.lut.size;
  uint32_t maxi = extend_mode == B2D_EXTEND_REFLECT ? lut_size * 2u - 1u : lut_size - 1u;
  uint32_t rori = extend_mode == B2D_EXTEND_REFLECT ? maxi : 0u;

  double ax = p1x - p0x, ay = p1y - p0y;
  double dist = ax * ax + ay * ay;
  double mx, my;
  map_point(transform, p0x, p0y, mx, my);
  double ox = 0.5 - mx, oy = 0.5 - my;

  double dt = ax * inv.m00 + ay * inv.m01;
  double dy = ax * inv.m10 + ay * inv.m11;
  double scale = double(int64_t(uint64_t(lut_size) << 32)) / dist;
  double offset = ox * dt + oy * dy;
  dt *= scale; dy *= scale; offset *= scale;

  fd.linear.dy.i64 = floor_to_int64(dy);
  fd.linear.dt.i64 = floor_to_int64(dt);
  fd.linear.pt[0].i64 = floor_to_int64(offset);
  fd.linear.pt[1].u64 = fd.linear.pt[0].u64 + fd.linear.dt.u64;
  fd.linear.maxi = maxi; fd.linear.rori = rori;
  uint32_t base = quality < B2D_GRADIENT_QUALITY_DITHER ? B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD : B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_PAD;
  return base + uint32_t(extend_mode != B2D_EXTEND_PAD);
}

inline double sq(double x) { return x * x; }
inline double dabs(double x) { return x < 0 ? -x : x; }

uint32_t init_radial_gradient(b2dgpu_fetch_gradient& fd, const double* v, uint32_t extend_mode, uint32_t quality, const Matrix& transform) {
  Matrix inv;
  if (!matrix_invert(inv, transform)) return kPending;
  uint32_t lut_size = fd.lut.size;
  uint32_t maxi = extend_mode == B2D_EXTEND_REFLECT ? lut_size * 2u - 1u : lut_size - 1u;
  uint32_t rori = extend_mode == B2D_EXTEND_REFLECT ? maxi : 0u;
  b2dgpu_gradient_radial& r = fd.radial;
  r.maxi = maxi; r.rori = rori;

  double cpx = v[0], cpy = v[1], fpx = v[2], fpy = v[3];
  double cr = v[4], fr = v[5];
  double dpx = cpx - fpx, dpy = cpy - fpy;
  double dr = cr - fr;

  double sq_d = sq(dpx) + sq(dpy);
  double d = sqrt(sq_d);
  double dist_from_border = dabs(d - dr);
  const double dist_limit = 0.5;
  if (dist_from_border < dist_limit) {
    double dp0x = (dpx * (dr - dist_limit)) / d, dp0y = (dpy * (dr - dist_limit)) / d;
    double dp1x = (dpx * (dr + dist_limit)) / d, dp1y = (dpy * (dr + dist_limit)) / d;
    double dp0_dist = dabs(sq(dp0x) + sq(dp0y) - sq_d);
    double dp1_dist = dabs(sq(dp1x) + sq(dp1y) - sq_d);
    if (dp0_dist < dp1_dist) { dpx = dp0x; dpy = dp0y; } else { dpx = dp1x; dpy = dp1y; }
    fpx = cpx - dpx; fpy = cpy - dpy;
    sq_d = sq(dpx) + sq(dpy);
  }

  double a = sq(dr) - sq_d;
  double sq_fr = sq(fr);
  double scale = double(lut_size);
  double xx = inv.m00, xy = inv.m01, yx = inv.m10, yy = inv.m11;
  double tpx = (inv.m20 + (xx + xy) * 0.5) - fpx;
  double tpy = (inv.m21 + (yx + yy) * 0.5) - fpy;

  r.tx = tpx; r.ty = tpy; r.yx = yx; r.yy = yy;
  double a_mul_4 = a * 4.0;
  double inv2a = (scale * 0.5) / a;
  double sq_inv2a = sq(inv2a);
  r.amul4 = a_mul_4; r.inv2a = inv2a; r.sq_inv2a = sq_inv2a; r.sq_fr = sq_fr;

  double sq_xx_plus_sq_yx = sq(xx) + sq(xy);
  double b0 = 2.0 * (dr * fr + tpx * dpx + tpy * dpy);
  double bx = 2.0 * (dpx * xx + dpy * xy);
  double by = 2.0 * (dpx * yx + dpy * yy);
  r.b0 = -b0; r.by = -by;

  double bx_mul_2 = bx * 2.0;
  double sq_bx = sq(bx);
  double dd0 = sq_bx + bx_mul_2 * b0 + a_mul_4 * (sq_xx_plus_sq_yx + 2.0 * (tpx * xx + tpy * xy));
  double ddy = bx_mul_2 * by + a_mul_4 * (2.0 * (xx * yx + yy * xy));
  double ddd_half = (sq_bx + a_mul_4 * sq_xx_plus_sq_yx);
  double ddd_half_inv = ddd_half * sq_inv2a;
  r.dd0 = dd0 - ddd_half;
  r.ddy = ddy;
  r.f32_bd = float(-bx * inv2a);
  r.f32_ddd = float(ddd_half_inv);
  uint32_t base = quality < B2D_GRADIENT_QUALITY_DITHER ? B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD : B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_PAD;
  return base + uint32_t(extend_mode != B2D_EXTEND_PAD);
}

uint32_t init_conic_gradient(b2dgpu_fetch_gradient& fd, const double* v, uint32_t quality, const Matrix& transform) {
  static const double q256[4] = { 4.071421038552e+1, -1.311160794048e+1, 6.017670215625, -1.623253505085 };
  double cx, cy;
  double angle = v[2], repeat = v[3];
  uint32_t lut_size = fd.lut.size;

  map_point(transform, v[0], v[1], cx, cy);
  cx = 0.5 - cx; cy = 0.5 - cy;

  double vx = 1.0 * transform.m00 + 0.0 * transform.m10;           // map_vector(1, 0)
  double vy = 1.0 * transform.m01 + 0.0 * transform.m11;
  double matrix_angle = atan2(vy, vx);

  Matrix updated = transform;
  double rot[3] = { -matrix_angle, cx, cy };
  matrix_apply_op(updated, 6, rot);

  angle += matrix_angle;
  double q = angle / -6.283185307179586476925;
  double off = q - floor(q);
  if (off != 0.0) off = -1.0 + off;

  Matrix inv;
  if (!matrix_invert(inv, updated)) return kPending;

  b2dgpu_gradient_conic& c = fd.conic;
  c.tx = cx * inv.m00 + cy * inv.m10;
  c.ty = cx * inv.m01 + cy * inv.m11;
  c.yx = inv.m10;
  c.yy = inv.m11;

  double lut_d = double(int(lut_size));
  double rep_size = lut_d * repeat;
  double q_scale = rep_size / 256.0;
  for (int i = 0; i < 4; i++) c.q_coeff[i] = float(q256[i] * q_scale);
  c.n_div_1_2_4[0] = float(rep_size);
  c.n_div_1_2_4[1] = float(rep_size * 0.5);
  c.n_div_1_2_4[2] = float(rep_size * 0.25);
  c.offset = float(off * rep_size - 0.5);
  c.xx = float(inv.m00);
  c.maxi = uint32_t(INT32_MAX);
  c.rori = lut_size - 1u;
  return quality < B2D_GRADIENT_QUALITY_DITHER ? B2DGPU_FETCH_GRADIENT_CONIC_NN : B2DGPU_FETCH_GRADIENT_CONIC_DITHER;
}

// A resolved style: what RenderFetchData + its signature hold in the reference.
struct Style {
  uint32_t kind;                // 0 = solid, 1 = non-solid, 2 = disabled (nop)
  uint32_t format;              // FormatExt of the source
  uint32_t solid_prgb32;
  uint32_t fetch_type;          // non-solid
  int32_t fetch_index;          // index in the current batch (-1 = not yet materialised in this batch)
  b2dgpu_fetch_data fd;
  const struct b2d_gradient* device_lut;   // non-null: fd.gradient.lut.data is NULL, the runtime interpolates the table from these stops
};

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------------------------
struct b2d_context {
  b2d_image* image;
  b2dgpu_runtime* rt;
  bool own_rt;
  b2dgpu_target* target;
  uint32_t dst_format;
  int32_t origin_x, origin_y;
  uint32_t queue_limit;
  bool adaptive_queue;          // command_queue_limit == 0: small first batch (the GPU starts early), then doubling
  bool record_only;             // commands are only queued (peek_batch); nothing can be flushed
  bool registered;              // the image's pixels were page-locked by this context

  // State.
  uint32_t comp_op;
  uint32_t fill_rule;
  double global_alpha, fill_alpha;
  uint32_t fill_alpha_i;
  uint32_t gradient_quality, pattern_quality;
  double tolerance;
  Matrix user;                  // == final transform (meta transform is identity)
  Matrix final_fixed;
  uint32_t final_type, final_fixed_type;
  bool integral_translation;
  int32_t tr_x, tr_y;
  Style style;
  uint32_t error_flags;

  // Batch under construction.
  std::vector<b2dgpu_command> cmds;
  std::vector<b2dgpu_fetch_data> fetch;
  std::vector<double> vtx;
  std::vector<b2dgpu_segment> segs;
  std::vector<b2dgpu_geometry_state> states;
  std::vector<b2dgpu_lut_request> lut_requests;      // gradient tables the device builds (b2dgpu_lut_request)
  std::vector<b2dgpu_gradient_stop> lut_stops;
  bool device_luts;                                   // GPU contexts: gradient tables are interpolated on the device
  bool state_valid;             // states.back() matches the current transform
  // Style objects whose memory (LUTs, pixels) the queued FetchData still points to - released after the flush,
  // like RenderFetchData's reference to its style (renderfetchdata_p.h:43, 181-184).
  std::vector<b2d_gradient*> kept_gradients;
  std::vector<b2d_pattern*> kept_patterns;
  std::vector<b2d_image*> kept_images;          // masks of queued fill_mask commands
  std::unordered_set<uint32_t> known_signatures; // PipeLookupCache stand-in
  bool dirty;                   // device canvas differs from the host image
};

namespace {

void update_transform(b2d_context* c) {                                       // on_after_user_transform_changed
  c->final_fixed = c->user;
  double s[2] = { 256.0, 256.0 };
  matrix_apply_op(c->final_fixed, 9, s);
  c->final_type = matrix_type(c->user);
  c->final_fixed_type = bmax<uint32_t>(c->final_type, kTTScale);
  c->integral_translation = false;
  if (c->final_type <= kTTTranslate) {
    const double lim = 16777216.0 * 256.0;
    if (c->final_fixed.m20 >= -lim && c->final_fixed.m20 <= lim && c->final_fixed.m21 >= -lim && c->final_fixed.m21 <= lim) {
      int64_t tx64 = floor_to_int64(c->final_fixed.m20), ty64 = floor_to_int64(c->final_fixed.m21);
      if (((tx64 | ty64) & 0xFF) == 0) { c->tr_x = int(tx64 >> 8); c->tr_y = int(ty64 >> 8); c->integral_translation = true; }
    }
  }
  c->state_valid = false;
}

void update_alpha(b2d_context* c) {
  c->fill_alpha_i = uint32_t(round_to_int(c->global_alpha * 255.0 * c->fill_alpha));
}

void release_kept(b2d_context* c, bool keep_current) {
  // The current fill style may still be used by the next batch: keep the most recent object of each kind.
  b2d_gradient* cur_g = (keep_current && !c->kept_gradients.empty()) ? c->kept_gradients.back() : nullptr;
  b2d_pattern* cur_p = (keep_current && !c->kept_patterns.empty()) ? c->kept_patterns.back() : nullptr;
  for (b2d_gradient* g : c->kept_gradients) if (g != cur_g) b2d_gradient_destroy(g);
  for (b2d_pattern* p : c->kept_patterns) if (p != cur_p) b2d_pattern_destroy(p);
  for (b2d_image* im : c->kept_images) b2d_image_destroy(im);
  c->kept_gradients.clear(); c->kept_patterns.clear(); c->kept_images.clear();
  if (cur_g) c->kept_gradients.push_back(cur_g);
  if (cur_p) c->kept_patterns.push_back(cur_p);
}

b2dgpu_result flush_batch(b2d_context* c) {
  if (c->cmds.empty()) return B2DGPU_SUCCESS;
  if (c->record_only) return B2DGPU_ERROR_INVALID_STATE;
  b2dgpu_batch_view v;
  memset(&v, 0, sizeof(v));
  b2d_context_peek_batch(c, &v);
  b2dgpu_result r = b2dgpu_submit(c->rt, c->target, &v);
  c->cmds.clear(); c->fetch.clear(); c->vtx.clear(); c->segs.clear(); c->states.clear(); c->lut_requests.clear(); c->lut_stops.clear();
  c->state_valid = false;
  c->style.fetch_index = -1;
  c->dirty = true;
  release_kept(c, true);
  if (r) c->error_flags |= 1;
  return r;
}

// Resolved render call: signature bits (dst|src|op), alpha, and where the source comes from.
struct Resolved { bool nop; uint32_t sig; uint32_t alpha; bool solid; uint32_t solid_prgb32; int32_t fetch_index; b2dgpu_result err; };

Resolved resolve(b2d_context* c, bool is_clear) {
  Resolved r; memset(&r, 0, sizeof(r));
  r.nop = true; r.fetch_index = -1;
  Style& st = c->style;
  uint32_t op = is_clear ? uint32_t(kOpSrcCopy) : c->comp_op;
  uint32_t src_format = is_clear ? P : st.format;
  if (!is_clear && (st.kind == 2 || c->fill_alpha_i == 0)) return r;

  // resolve_clear_op (rastercontext.cpp:1231-1243): SRC_COPY of the transparent solid from solid_override_fill_table.
  Simplified s = simplify(op, c->dst_format, src_format);
  if (s.solid == kSolidNop) return r;
  if (!s.implemented) { r.err = B2DGPU_ERROR_NOT_IMPLEMENTED; return r; }

  r.nop = false;
  r.alpha = is_clear ? 255u : c->fill_alpha_i;
  r.sig = B2DGPU_MAKE_SIG(s.dst, s.src, s.op, 0, 0);
  if (is_clear) {
    r.solid = true;
    r.solid_prgb32 = s.solid == kSolidOpaqueBlack ? 0xFF000000u : s.solid == kSolidOpaqueWhite ? 0xFFFFFFFFu : 0x00000000u;
  }
  else if (s.solid != kSolidNone) {
    // solid_fetch_data_override_table: the style is replaced by a constant colour.
    r.solid = true;
    r.solid_prgb32 = s.solid == kSolidTransparent ? 0x00000000u : s.solid == kSolidOpaqueBlack ? 0xFF000000u : 0xFFFFFFFFu;
  }
  else if (st.kind == 0) { r.solid = true; r.solid_prgb32 = st.solid_prgb32; }
  else {
    r.solid = false;
    r.sig |= st.fetch_type << 16;
    if (st.fetch_index < 0) {
      st.fetch_index = int32_t(c->fetch.size()); c->fetch.push_back(st.fd);
      if (st.device_lut) {
        b2dgpu_lut_request q;
        q.fetch_index = uint32_t(st.fetch_index); q.stop_offset = uint32_t(c->lut_stops.size());
        q.stop_count = uint32_t(st.device_lut->stops.size()); q.lut_size = st.fd.gradient.lut.size;
        for (const Stop& sp : st.device_lut->stops) c->lut_stops.push_back(b2dgpu_gradient_stop{ sp.offset, sp.rgba });
        c->lut_requests.push_back(q);
      }
    }
    r.fetch_index = st.fetch_index;
  }
  return r;
}

b2dgpu_command make_command(const Resolved& r, uint32_t type, uint32_t fill_type) {
  b2dgpu_command cmd; memset(&cmd, 0, sizeof(cmd));
  cmd.type = type;
  cmd.signature = r.sig | (fill_type << 14);
  cmd.alpha = r.alpha;
  cmd.solid_prgb32 = r.solid_prgb32;
  cmd.fetch_index = r.fetch_index < 0 ? 0u : uint32_t(r.fetch_index);
  cmd.fill_rule_mask = B2DGPU_FILL_RULE_MASK_NON_ZERO;
  return cmd;
}

b2dgpu_result push_command(b2d_context* c, const b2dgpu_command& cmd) {
  // Pipeline lookup (ensure_fetch_and_dispatch_data -> PipeProvider::get, rastercontext.cpp:1920-2029): the first time
  // a signature is used the runtime is asked for it; BL_ERROR_NOT_IMPLEMENTED surfaces to the caller like the
  // reference's static runtime does (fixedpiperuntime.cpp:314-315).
  if (c->rt && c->known_signatures.find(cmd.signature) == c->known_signatures.end()) {
    b2dgpu_dispatch_data dd;
    b2dgpu_result lr = b2dgpu_runtime_get(c->rt, cmd.signature, &dd, nullptr);
    if (lr) return lr;
    c->known_signatures.insert(cmd.signature);
  }
  c->cmds.push_back(cmd);
  if (c->cmds.size() >= c->queue_limit) {
    b2dgpu_result fr = flush_batch(c);
    // Adaptive batching: the first implicit flush of a frame comes early so the GPU starts while the host is still
    // recording; later batches double in size, which amortises the per-batch kernels (b2dgpu_submit is pipelined).
    if (c->adaptive_queue && c->queue_limit < 8192u) c->queue_limit *= 2u;
    return fr;
  }
  return B2DGPU_SUCCESS;
}

b2dgpu_result fill_box_a(b2d_context* c, const Resolved& r, int x0, int y0, int x1, int y1) {
  b2dgpu_command cmd = make_command(r, B2DGPU_CMD_FILL_BOX_A, B2DGPU_FILL_BOX_A);
  cmd.box[0] = x0; cmd.box[1] = y0; cmd.box[2] = x1; cmd.box[3] = y1;
  return push_command(c, cmd);
}

b2dgpu_result fill_box_f(b2d_context* c, const Resolved& r, int x0, int y0, int x1, int y1) {       // fill_clipped_box_f
  if (((x0 | y0 | x1 | y1) & 0xFF) == 0) return fill_box_a(c, r, x0 >> 8, y0 >> 8, x1 >> 8, y1 >> 8);
  b2dgpu_command cmd = make_command(r, B2DGPU_CMD_FILL_BOX_U, B2DGPU_FILL_MASK);
  cmd.box[0] = x0; cmd.box[1] = y0; cmd.box[2] = x1; cmd.box[3] = y1;
  return push_command(c, cmd);
}

uint32_t current_state(b2d_context* c, const Matrix& m, uint32_t transform_type) {
  b2dgpu_geometry_state gs; memset(&gs, 0, sizeof(gs));
  gs.m[0] = m.m00; gs.m[1] = m.m01; gs.m[2] = m.m10; gs.m[3] = m.m11; gs.m[4] = m.m20; gs.m[5] = m.m21;
  gs.clip[0] = 0.0; gs.clip[1] = 0.0; gs.clip[2] = double(c->image->w) * 256.0; gs.clip[3] = double(c->image->h) * 256.0;
  double t = c->tolerance * 256.0;
  gs.tolerance_sq = t * t;
  gs.transform_type = transform_type;
  if (!c->states.empty() && memcmp(&c->states.back(), &gs, sizeof(gs)) == 0) return uint32_t(c->states.size() - 1);
  c->states.push_back(gs);
  return uint32_t(c->states.size() - 1);
}

// Serialises a path view into segments (EdgeSourcePath + EdgeBuilder::add_from_source, edgebuilder_p.h:165-281,
// 1032-1070): every figure starts at a MOVE, is implicitly closed, and curve commands need all of their vertices.
b2dgpu_result fill_path_segments(b2d_context* c, const Resolved& r, const uint8_t* cmd, const double* vtx, uint32_t n,
                                 const Matrix& m, uint32_t transform_type, uint32_t fill_rule) {
  // The style was resolved into this batch already (r.fetch_index), so the batch cannot be flushed here; 2^30 vertices
  // (16 GiB) in one batch is not a workload, it is an error.
  if (c->vtx.size() / 2 + n > 0x3FFFFFF0u) return B2DGPU_ERROR_OUT_OF_MEMORY;
  uint32_t base = uint32_t(c->vtx.size() / 2);
  uint32_t seg_begin = uint32_t(c->segs.size());
  uint32_t cmd_index = uint32_t(c->cmds.size());
  c->vtx.insert(c->vtx.end(), vtx, vtx + size_t(n) * 2);

  auto add = [&](uint32_t p0, uint32_t p1, uint32_t kind) {
    b2dgpu_segment s; s.p0 = base + p0; s.p1_kind = ((base + p1) << 2) | kind; s.command = cmd_index;
    c->segs.push_back(s);
  };

  uint32_t i = 0;
  while (i < n) {
    if (cmd[i] != B2D_PATH_CMD_MOVE) { i++; continue; }
    uint32_t start = i, cur = i;
    i++;
    for (;;) {
      if (i < n && cmd[i] == B2D_PATH_CMD_ON) { add(cur, i, B2DGPU_SEG_LINE); cur = i; i++; }
      else if (i + 2 <= n && cmd[i] == B2D_PATH_CMD_QUAD) { add(cur, i, B2DGPU_SEG_QUAD); cur = i + 1; i += 2; }
      else if (i + 2 < n && cmd[i] == B2D_PATH_CMD_CUBIC) { add(cur, i, B2DGPU_SEG_CUBIC); cur = i + 2; i += 3; }
      else if (i + 2 < n && cmd[i] == B2D_PATH_CMD_CONIC) { add(cur, i, B2DGPU_SEG_CONIC); cur = i + 2; i += 2; }
      else { add(cur, start, B2DGPU_SEG_LINE); break; }
    }
  }

  uint32_t seg_count = uint32_t(c->segs.size()) - seg_begin;
  if (!seg_count) { c->vtx.resize(size_t(base) * 2); return B2DGPU_SUCCESS; }

  b2dgpu_command out = make_command(r, B2DGPU_CMD_FILL_GEOMETRY, B2DGPU_FILL_ANALYTIC);
  out.fill_rule_mask = fill_rule == B2D_FILL_RULE_NON_ZERO ? B2DGPU_FILL_RULE_MASK_NON_ZERO : B2DGPU_FILL_RULE_MASK_EVEN_ODD;
  out.data_offset = seg_begin;
  out.data_count = seg_count;
  out.state_index = current_state(c, m, transform_type);
  return push_command(c, out);
}

b2dgpu_result fill_box_d(b2d_context* c, const Resolved& r, double bx0, double by0, double bx1, double by1) {   // fill_unclipped_box_d
  const Matrix& t = c->final_fixed;
  if (c->final_fixed_type <= kTTSwap) {
    double x0 = bx0 * t.m00 + by0 * t.m10 + t.m20, y0 = bx0 * t.m01 + by0 * t.m11 + t.m21;
    double x1 = bx1 * t.m00 + by1 * t.m10 + t.m20, y1 = bx1 * t.m01 + by1 * t.m11 + t.m21;
    double mx0 = bmin(x0, x1), my0 = bmin(y0, y1), mx1 = bmax(x0, x1), my1 = bmax(y0, y1);
    double cw = double(c->image->w) * 256.0, ch = double(c->image->h) * 256.0;
    double fx0 = bmax(mx0, 0.0), fy0 = bmax(my0, 0.0), fx1 = bmin(mx1, cw), fy1 = bmin(my1, ch);
    if (!((fx0 < fx1) & (fy0 < fy1))) return B2DGPU_SUCCESS;
    int ix0 = trunc_to_int(fx0), iy0 = trunc_to_int(fy0), ix1 = trunc_to_int(fx1), iy1 = trunc_to_int(fy1);
    if (ix0 >= ix1 || iy0 >= iy1) return B2DGPU_SUCCESS;
    return fill_box_f(c, r, ix0, iy0, ix1, iy1);
  }
  double poly[8] = { bx0, by0, bx1, by0, bx1, by1, bx0, by1 };
  uint8_t cmds[4] = { B2D_PATH_CMD_MOVE, B2D_PATH_CMD_ON, B2D_PATH_CMD_ON, B2D_PATH_CMD_ON };
  return fill_path_segments(c, r, cmds, poly, 4, c->final_fixed, c->final_fixed_type, B2D_FILL_RULE_EVEN_ODD);
}

b2dgpu_result set_non_solid_style(b2d_context* c, uint32_t fetch_type, uint32_t format, const b2dgpu_fetch_data& fd) {
  Style& st = c->style;
  if (fetch_type == kPending) { st.kind = 2; return B2DGPU_SUCCESS; }
  st.kind = 1; st.format = format; st.fetch_type = fetch_type; st.fetch_index = -1; st.fd = fd; st.device_lut = nullptr;
  return B2DGPU_SUCCESS;
}

} // namespace

extern "C" b2dgpu_result b2d_context_create(b2d_image* target, const b2d_context_create_info* info, b2d_context** out) {
  if (!out || !target) return B2DGPU_ERROR_INVALID_VALUE;
  *out = nullptr;
  b2d_context* c = new (std::nothrow) b2d_context();
  if (!c) return B2DGPU_ERROR_OUT_OF_MEMORY;
  c->image = target;
  c->rt = info ? static_cast<b2dgpu_runtime*>(info->runtime) : nullptr;
  c->own_rt = false;
  c->target = nullptr;
  c->record_only = info && (info->flags & B2D_CONTEXT_CREATE_FLAG_RECORD_ONLY);
  // A recorded batch may be rendered by the test-only host simulator, which needs real tables; a GPU context lets the
  // device interpolate them (B2D_HOST_DEVICE_LUTS=0 keeps the host tables, for A/B tests).
  { const char* e = getenv("B2D_HOST_DEVICE_LUTS"); c->device_luts = !c->record_only && !(e && e[0] == '0'); }
  b2dgpu_result r = B2DGPU_SUCCESS;
  if (c->record_only) c->rt = nullptr;
  else if (!c->rt) {
    b2dgpu_create_info ci; memset(&ci, 0, sizeof(ci));
    ci.struct_size = sizeof(ci); ci.device = info ? info->device : 0; ci.stream = info ? info->stream : nullptr;
    r = b2dgpu_runtime_create(&ci, &c->rt);
    if (r) { delete c; return r; }
    c->own_rt = true;
  }
  if (!c->record_only) {
    int y0 = 0, y1 = target->h;
    if (info && (info->slab_y0 || info->slab_y1)) { y0 = info->slab_y0; y1 = info->slab_y1; }
    r = b2dgpu_target_create_slab(c->rt, target->w, target->h, y0, y1, target->format, &c->target);
  }
  c->registered = false;
  if (!r && !c->record_only) {
    b2dgpu_image_data id; b2d_image_get_data(target, &id);
    // Large canvases are page-locked for the lifetime of the context: flush(SYNC) then copies straight into them.
    const size_t image_bytes = size_t(id.stride < 0 ? -id.stride : id.stride) * size_t(target->h);
    if (image_bytes >= (size_t(4) << 20) && id.stride > 0)
      c->registered = b2dgpu_host_register(c->rt, id.pixel_data, image_bytes) == B2DGPU_SUCCESS;
    r = b2dgpu_target_upload(c->target, &id);
  }
  if (r) {
    if (c->registered) { b2dgpu_image_data id; b2d_image_get_data(target, &id); b2dgpu_host_unregister(c->rt, id.pixel_data); }
    if (c->target) b2dgpu_target_destroy(c->target);
    if (c->own_rt) b2dgpu_runtime_destroy(c->rt);
    delete c;
    return r;
  }
  c->dst_format = target->format;
  c->origin_x = info ? info->pixel_origin_x : 0;
  c->origin_y = info ? info->pixel_origin_y : 0;
  c->adaptive_queue = !c->record_only && !(info && info->command_queue_limit);
  c->queue_limit = c->record_only ? 0xFFFFFFFFu : c->adaptive_queue ? 512u : info->command_queue_limit;
  c->comp_op = kOpSrcOver;
  c->fill_rule = B2D_FILL_RULE_NON_ZERO;
  c->global_alpha = 1.0; c->fill_alpha = 1.0;
  c->gradient_quality = B2D_GRADIENT_QUALITY_NEAREST;
  c->pattern_quality = B2D_PATTERN_QUALITY_BILINEAR;
  c->tolerance = 0.2;
  c->user = kIdentity;
  c->error_flags = 0;
  c->state_valid = false;
  c->dirty = false;
  // Default fill style: opaque black (rastercontext.cpp:224-233).
  c->style.kind = 0; c->style.format = F; c->style.solid_prgb32 = 0xFF000000u; c->style.fetch_index = -1;
  update_transform(c);
  update_alpha(c);
  *out = c;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_flush(b2d_context* c, uint32_t flags) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  if (c->record_only) return B2DGPU_SUCCESS;
  b2dgpu_result r = flush_batch(c);
  if (r) return r;
  if (c->adaptive_queue) c->queue_limit = 512u;           // an explicit flush ends the frame: start small again
  if (flags & 0x80000000u) {
    if (c->dirty) {
      b2dgpu_image_data id; b2d_image_get_data(c->image, &id);
      r = b2dgpu_target_download(c->target, &id);      // synchronises the stream
      if (!r) c->dirty = false;
    }
    else r = b2dgpu_sync(c->rt);
  }
  return r;
}

extern "C" b2dgpu_result b2d_context_end(b2d_context* c) { return b2d_context_flush(c, 0x80000000u); }

extern "C" b2dgpu_result b2d_context_destroy(b2d_context* c) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  b2dgpu_result r = b2d_context_end(c);
  release_kept(c, false);
  if (c->registered) { b2dgpu_image_data id; b2d_image_get_data(c->image, &id); b2dgpu_host_unregister(c->rt, id.pixel_data); }
  if (c->target) b2dgpu_target_destroy(c->target);
  if (c->own_rt) b2dgpu_runtime_destroy(c->rt);
  delete c;
  return r;
}

extern "C" b2dgpu_runtime* b2d_context_runtime(b2d_context* c) { return c ? c->rt : nullptr; }
extern "C" b2dgpu_target* b2d_context_target(b2d_context* c) { return c ? c->target : nullptr; }

extern "C" b2dgpu_result b2d_context_peek_batch(b2d_context* c, b2dgpu_batch_view* v) {
  if (!c || !v) return B2DGPU_ERROR_INVALID_VALUE;
  memset(v, 0, sizeof(*v));
  v->struct_size = sizeof(*v);
  v->command_count = uint32_t(c->cmds.size()); v->commands = c->cmds.data();
  v->fetch_data = c->fetch.data(); v->fetch_count = uint32_t(c->fetch.size());
  v->vertices = c->vtx.data(); v->vertex_count = uint32_t(c->vtx.size() / 2);
  v->segments = c->segs.data(); v->segment_count = uint32_t(c->segs.size());
  v->geometry_states = c->states.data(); v->geometry_state_count = uint32_t(c->states.size());
  v->pixel_origin_x = c->origin_x; v->pixel_origin_y = c->origin_y;
  v->lut_requests = c->lut_requests.data(); v->lut_request_count = uint32_t(c->lut_requests.size());
  v->lut_stops = c->lut_stops.data(); v->lut_stop_count = uint32_t(c->lut_stops.size());
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_discard_batch(b2d_context* c) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  c->cmds.clear(); c->fetch.clear(); c->vtx.clear(); c->segs.clear(); c->states.clear(); c->lut_requests.clear(); c->lut_stops.clear();
  c->state_valid = false; c->style.fetch_index = -1;
  release_kept(c, true);
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_set_comp_op(b2d_context* c, uint32_t comp_op) {
  if (!c || comp_op > 28) return B2DGPU_ERROR_INVALID_VALUE;
  c->comp_op = comp_op;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_set_global_alpha(b2d_context* c, double alpha) {
  if (!c || alpha != alpha) return B2DGPU_ERROR_INVALID_VALUE;
  c->global_alpha = bclamp(alpha, 0.0, 1.0);
  update_alpha(c);
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_set_fill_alpha(b2d_context* c, double alpha) {
  if (!c || alpha != alpha) return B2DGPU_ERROR_INVALID_VALUE;
  c->fill_alpha = bclamp(alpha, 0.0, 1.0);
  update_alpha(c);
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_set_fill_rule(b2d_context* c, uint32_t fill_rule) {
  if (!c || fill_rule > 1) return B2DGPU_ERROR_INVALID_VALUE;
  c->fill_rule = fill_rule;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_set_hint(b2d_context* c, uint32_t hint, uint32_t value) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  switch (hint) {
    case B2D_HINT_RENDERING_QUALITY: return value <= 0 ? B2DGPU_SUCCESS : B2DGPU_ERROR_INVALID_VALUE;
    case B2D_HINT_GRADIENT_QUALITY: if (value > 2) return B2DGPU_ERROR_INVALID_VALUE; c->gradient_quality = value; return B2DGPU_SUCCESS;
    case B2D_HINT_PATTERN_QUALITY: if (value > 1) return B2DGPU_ERROR_INVALID_VALUE; c->pattern_quality = value; return B2DGPU_SUCCESS;
    default: return B2DGPU_ERROR_INVALID_VALUE;
  }
}

extern "C" b2dgpu_result b2d_context_set_flatten_tolerance(b2d_context* c, double tolerance) {
  if (!c || tolerance != tolerance) return B2DGPU_ERROR_INVALID_VALUE;
  c->tolerance = bclamp(tolerance, 0.0001, 0.9);       // rastercontext.cpp:1869-1885 (ContextInternal limits)
  c->state_valid = false;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_set_fill_style_rgba32(b2d_context* c, uint32_t rgba32) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  c->style.kind = 0;
  c->style.format = format_from_rgba32(rgba32);
  c->style.solid_prgb32 = premultiply_argb32(rgba32);
  c->style.fetch_index = -1;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_set_fill_style_gradient(b2d_context* c, const b2d_gradient* gc) {
  if (!c || !gc) return B2DGPU_ERROR_INVALID_VALUE;
  b2d_gradient* g = const_cast<b2d_gradient*>(gc);
  if (g->empty) { c->style.kind = 2; return B2DGPU_SUCCESS; }
  if (g->solid) {
    uint32_t rgba32 = rgba32_from_rgba64(g->stops.back().rgba);
    c->style.kind = 0; c->style.format = g->format; c->style.solid_prgb32 = premultiply_argb32(rgba32); c->style.fetch_index = -1;
    return B2DGPU_SUCCESS;
  }

  Matrix m = c->user;
  if (g->transform_type != kTTIdentity) matrix_multiply(m, g->transform, c->user);

  uint32_t quality = c->gradient_quality;
  if (c->dst_format == B2DGPU_FORMAT_A8) quality = B2D_GRADIENT_QUALITY_NEAREST;
  bool dither = quality >= B2D_GRADIENT_QUALITY_DITHER;
  uint32_t lut_size = dither ? bmin<uint32_t>(g->lut_size * 2, 1024) : g->lut_size;

  b2dgpu_fetch_data fd; memset(&fd, 0, sizeof(fd));
  if (dither) {
    if (!g->lut64) { g->lut64.reset(new uint64_t[lut_size]); make_lut64(g->lut64.get(), lut_size, g->stops.data(), g->stops.size()); }
    fd.gradient.lut.data = g->lut64.get();
  }
  else if (c->device_luts && !g->lut32) {
    fd.gradient.lut.data = nullptr;               // k_build_luts interpolates it from the stops (b2dgpu_lut_request)
  }
  else {
    if (!g->lut32) { g->lut32.reset(new uint32_t[lut_size]); make_lut32(g->lut32.get(), lut_size, g->stops.data(), g->stops.size()); }
    fd.gradient.lut.data = g->lut32.get();
  }
  fd.gradient.lut.size = lut_size;

  uint32_t ft;
  if (g->type == B2D_GRADIENT_LINEAR) ft = init_linear_gradient(fd.gradient, g->values, g->extend_mode, quality, m);
  else if (g->type == B2D_GRADIENT_RADIAL) ft = init_radial_gradient(fd.gradient, g->values, g->extend_mode, quality, m);
  else ft = init_conic_gradient(fd.gradient, g->values, quality, m);
  g->refs++;
  c->kept_gradients.push_back(g);
  b2dgpu_result sr = set_non_solid_style(c, ft, g->format, fd);
  if (sr == B2DGPU_SUCCESS && c->style.kind == 1 && !dither && !fd.gradient.lut.data) c->style.device_lut = g;
  return sr;
}

extern "C" b2dgpu_result b2d_context_set_fill_style_pattern(b2d_context* c, const b2d_pattern* p) {
  if (!c || !p) return B2DGPU_ERROR_INVALID_VALUE;
  if (!p->area[2] || !p->area[3]) { c->style.kind = 2; return B2DGPU_SUCCESS; }
  Matrix m = c->user;
  if (p->transform_type != kTTIdentity) matrix_multiply(m, p->transform, c->user);
  b2dgpu_fetch_data fd; memset(&fd, 0, sizeof(fd));
  const b2d_image* img = p->image;
  fd.pattern.src.pixel_data = img->data + intptr_t(p->area[1]) * img->stride + intptr_t(p->area[0]) * img->bpp;
  fd.pattern.src.stride = img->stride;
  fd.pattern.src.w = p->area[2]; fd.pattern.src.h = p->area[3];
  uint32_t ft = init_pattern_affine(fd.pattern, p->extend_mode, c->pattern_quality, uint32_t(img->bpp), m);
  const_cast<b2d_pattern*>(p)->refs++;
  c->kept_patterns.push_back(const_cast<b2d_pattern*>(p));
  return set_non_solid_style(c, ft, img->format, fd);
}

extern "C" b2dgpu_result b2d_context_apply_transform_op(b2d_context* c, uint32_t op, const double* data) {
  if (!c || (op != 0 && !data)) return B2DGPU_ERROR_INVALID_VALUE;
  if (!matrix_apply_op(c->user, op, data)) return B2DGPU_ERROR_INVALID_VALUE;
  update_transform(c);
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2d_context_clear_all(b2d_context* c) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  Resolved r = resolve(c, true);
  if (r.err) return r.err;
  if (r.nop) return B2DGPU_SUCCESS;
  return fill_box_a(c, r, 0, 0, c->image->w, c->image->h);
}

extern "C" b2dgpu_result b2d_context_fill_all(b2d_context* c) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  Resolved r = resolve(c, false);
  if (r.err) return r.err;
  if (r.nop) return B2DGPU_SUCCESS;
  return fill_box_a(c, r, 0, 0, c->image->w, c->image->h);
}

extern "C" b2dgpu_result b2d_context_fill_rect_i(b2d_context* c, int32_t x, int32_t y, int32_t w, int32_t h) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  if (c->final_type >= kTTInvalid) return B2DGPU_SUCCESS;
  Resolved r = resolve(c, false);
  if (r.err) return r.err;
  if (r.nop) return B2DGPU_SUCCESS;
  if (!c->integral_translation) {
    if ((w <= 0) | (h <= 0)) return B2DGPU_SUCCESS;
    return fill_box_d(c, r, double(x), double(y), double(x) + double(w), double(y) + double(h));
  }
  int64_t x0 = int64_t(x) + c->tr_x, y0 = int64_t(y) + c->tr_y;
  int64_t x1 = int64_t(w) + x0, y1 = int64_t(h) + y0;
  x0 = bmax<int64_t>(x0, 0); y0 = bmax<int64_t>(y0, 0);
  x1 = bmin<int64_t>(x1, c->image->w); y1 = bmin<int64_t>(y1, c->image->h);
  if ((x0 >= x1) | (y0 >= y1)) return B2DGPU_SUCCESS;
  return fill_box_a(c, r, int(x0), int(y0), int(x1), int(y1));
}

// bl_context_fill_mask_i: fill_mask_i_impl -> translate_and_clip_rect_to_blit_i -> fill_clipped_box_masked_a
// (rastercontext.cpp:3723-3742, 922-1000, 3135-3175).  Only what the reference itself implements: an integral
// translation; its fill_unclipped_mask_d() returns NOT_IMPLEMENTED for everything that is not pixel aligned.
extern "C" b2dgpu_result b2d_context_fill_mask_i(b2d_context* c, int32_t x, int32_t y, const b2d_image* mask, const int32_t* area) {
  if (!c || !mask) return B2DGPU_ERROR_INVALID_VALUE;
  if (c->final_type >= kTTInvalid) return B2DGPU_SUCCESS;
  int sx = 0, sy = 0, w = mask->w, h = mask->h;
  if (area) {
    unsigned max_w = unsigned(w) - unsigned(area[0]), max_h = unsigned(h) - unsigned(area[1]);
    if ((max_w > unsigned(w)) | (unsigned(area[2]) > max_w) | (max_h > unsigned(h)) | (unsigned(area[3]) > max_h)) return B2DGPU_ERROR_INVALID_VALUE;
    sx = area[0]; sy = area[1]; w = area[2]; h = area[3];
  }
  if (!c->integral_translation) return B2DGPU_ERROR_NOT_IMPLEMENTED;
  if (mask->format != B2DGPU_FORMAT_A8) return B2DGPU_ERROR_NOT_IMPLEMENTED;         // the reference reads the mask rows as bytes
  int64_t dx = int64_t(x) + c->tr_x, dy = int64_t(y) + c->tr_y;
  int64_t x0 = dx, y0 = dy, x1 = dx + w, y1 = dy + h;
  x0 = bmax<int64_t>(x0, 0); y0 = bmax<int64_t>(y0, 0);
  x1 = bmin<int64_t>(x1, c->image->w); y1 = bmin<int64_t>(y1, c->image->h);
  if ((x0 >= x1) | (y0 >= y1)) return B2DGPU_SUCCESS;
  sx += int(x0 - dx); sy += int(y0 - dy);
  Resolved r = resolve(c, false);
  if (r.err) return r.err;
  if (r.nop) return B2DGPU_SUCCESS;

  b2dgpu_fetch_data fd; memset(&fd, 0, sizeof(fd));
  fd.pattern.src.pixel_data = mask->data + intptr_t(sy) * mask->stride + intptr_t(sx);
  fd.pattern.src.stride = mask->stride;
  fd.pattern.src.w = int32_t(x1 - x0); fd.pattern.src.h = int32_t(y1 - y0);
  b2dgpu_command cmd = make_command(r, B2DGPU_CMD_FILL_BOX_MASK_A, B2DGPU_FILL_MASK);
  cmd.box[0] = int(x0); cmd.box[1] = int(y0); cmd.box[2] = int(x1); cmd.box[3] = int(y1);
  cmd.reserved[0] = uint32_t(c->fetch.size());
  c->fetch.push_back(fd);
  const_cast<b2d_image*>(mask)->refs++;                                 // retained until the batch is submitted
  c->kept_images.push_back(const_cast<b2d_image*>(mask));
  return push_command(c, cmd);
}

extern "C" b2dgpu_result b2d_context_fill_rect_d(b2d_context* c, double x, double y, double w, double h) {
  if (!c) return B2DGPU_ERROR_INVALID_VALUE;
  if (c->final_type >= kTTInvalid) return B2DGPU_SUCCESS;
  Resolved r = resolve(c, false);
  if (r.err) return r.err;
  if (r.nop) return B2DGPU_SUCCESS;
  return fill_box_d(c, r, x, y, x + w, y + h);
}

extern "C" b2dgpu_result b2d_context_fill_path_d(b2d_context* c, double ox, double oy, const uint8_t* cmd, const double* vtx, uint32_t count) {
  if (!c || (count && (!cmd || !vtx))) return B2DGPU_ERROR_INVALID_VALUE;
  if (!count || c->final_type >= kTTInvalid) return B2DGPU_SUCCESS;
  Resolved r = resolve(c, false);
  if (r.err) return r.err;
  if (r.nop) return B2DGPU_SUCCESS;
  const Matrix& ft = c->final_fixed;
  double fx, fy;
  map_point(ft, ox, oy, fx, fy);
  Matrix m = { ft.m00, ft.m01, ft.m10, ft.m11, fx, fy };
  return fill_path_segments(c, r, cmd, vtx, count, m, bmax<uint32_t>(c->final_fixed_type, kTTTranslate), c->fill_rule);
}

extern "C" b2dgpu_result b2d_context_fill_polygon_d(b2d_context* c, const double* pts, uint32_t count) {
  if (!c || (count && !pts)) return B2DGPU_ERROR_INVALID_VALUE;
  if (count < 1 || c->final_type >= kTTInvalid) return B2DGPU_SUCCESS;
  Resolved r = resolve(c, false);
  if (r.err) return r.err;
  if (r.nop) return B2DGPU_SUCCESS;
  std::vector<uint8_t> cmds(count, uint8_t(B2D_PATH_CMD_ON));
  cmds[0] = B2D_PATH_CMD_MOVE;
  return fill_path_segments(c, r, cmds.data(), pts, count, c->final_fixed, c->final_fixed_type, c->fill_rule);
}
