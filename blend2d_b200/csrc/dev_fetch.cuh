// dev_fetch.cuh - source fetchers (solid, gradients, patterns) in CLOSED FORM of the destination pixel (x, y).
//
// The reference fetchers (blend2d/pipeline/reference/fetchgeneric_p.h) are incremental: spanStartX()/fetch()/advance_y()
// walk a span pixel by pixel.  Every one of them is, by construction, a pure function of (x, y) - except the conic
// gradient, whose row origin is accumulated in double (noted below) - so the GPU evaluates that function directly and
// any thread can fetch any pixel.  Integer parts wrap exactly like the reference (u64 / u32 arithmetic); float parts
// use the reference's operation order with unfused multiply-add (compile with -fmad=false) and x86 conversion
// semantics (cvttss2si / cvtss2si return INT_MIN when out of range).
//
//   FetchLinearGradient      :939-1011    FetchRadialGradient  :1016-1132   FetchConicGradient :1143-1254
//   FetchGradientBase dither :899-934     (Bayer 16x16: blend2d/tables/tables_p.h:452-473)
//   FetchPatternAligned*     :470-596     FetchPatternFxFy*    :601-693     FetchPatternAffine{NN,BI} :354-465, :698-797
#pragma once
#include "dev_common.cuh"
#include "dev_pixel.cuh"
#include "../../include/b2dgpu.h"
#include <math.h>

namespace b2d {

// x86 float -> int conversions (cvttss2si / cvtss2si return INT_MIN for NaN and for anything outside the int range).
// Device variants: cvt.rzi / cvt.rni saturate, so values below -2^31 already give INT_MIN; NaN gives 0 instead of
// INT_MIN, which is indistinguishable in the only users, the gradient table indices (pad clamps both to 0, repeat /
// reflect / conic mask both to 0 because the masks are below 2^31); only v >= 2^31 needs the x86 result forced.
B2D_HD int x86_trunc_f32(float v) {
#if defined(__CUDA_ARCH__)
  return v >= 2147483648.0f ? int(0x80000000u) : __float2int_rz(v);
#else
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return int(0x80000000u);
  return int(v);
#endif
}
B2D_HD int x86_nearby_f32(float v) {
#if defined(__CUDA_ARCH__)
  return v >= 2147483648.0f ? int(0x80000000u) : __float2int_rn(v);
#else
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return int(0x80000000u);
  return int(lrintf(v));
#endif
}

// bl_abs(float) of the reference is `v < 0 ? -v : v`, which keeps -0.0f where fabsf() gives +0.0f.  In the two
// gradient fetchers that use it the sign of a zero never reaches the table index (it can only flow into additions,
// a square root and a float -> int conversion, all of which map -0 and +0 to the same index; the one division is
// 0 / 0 = NaN for either sign), so the device uses the free |x| operand modifier.
B2D_HD float f32_abs(float v) {
#if defined(__CUDA_ARCH__)
  return fabsf(v);
#else
  return v < 0.0f ? -v : v;
#endif
}
B2D_HD uint32_t f32_bits(float v) { union { float f; uint32_t u; } c; c.f = v; return c.u; }

// Everything a fetcher needs besides (x, y).
struct FetchEnv {
  const b2dgpu_fetch_data* fd;     // device copy; pointers inside already point to device memory
  const uint8_t* bayer;            // 16 rows x 32 entries
  uint32_t fetch_type;             // B2DGPU_FETCH_*
  uint32_t src_format;             // B2DGPU_FORMAT_* of a pattern source
  uint32_t solid;                  // FETCH_SOLID colour
  int origin_x, origin_y;          // ContextData::pixel_origin
};

// ---------------------------------------------------------------------------------------------------------------
// Gradient table lookup (+ optional ordered dithering).
// ---------------------------------------------------------------------------------------------------------------
B2D_HD uint32_t lut_fetch_nn(const b2dgpu_fetch_gradient& g, uint32_t idx) {
  return static_cast<const uint32_t*>(g.lut.data)[idx];
}

B2D_HD uint32_t lut_fetch_dither(const FetchEnv& env, const b2dgpu_fetch_gradient& g, uint32_t idx, uint32_t x, uint32_t y) {
  uint64_t v = static_cast<const uint64_t*>(g.lut.data)[idx];                // BLRgba64: a[63:48] r[47:32] g[31:16] b[15:0]
  uint32_t dm = ((uint32_t(env.origin_y) + y) & 15u) * 32u + (uint32_t(env.origin_x) & 15u) + (x & 15u);
  uint32_t dd = env.bayer[dm];
  uint32_t a = uint32_t(v >> 56);
  uint32_t r = tmin<uint32_t>((uint32_t((v >> 32) & 0xFFFFu) + dd) >> 8, a);
  uint32_t gg = tmin<uint32_t>((uint32_t((v >> 16) & 0xFFFFu) + dd) >> 8, a);
  uint32_t b = tmin<uint32_t>((uint32_t(v & 0xFFFFu) + dd) >> 8, a);
  return (a << 24) | (r << 16) | (gg << 8) | b;
}

B2D_HD uint32_t grad_index_pad(uint32_t idx, uint32_t maxi) { return uint32_t(tclamp<int32_t>(int32_t(idx), 0, int32_t(maxi))); }
B2D_HD uint32_t grad_index_ror(uint32_t idx, uint32_t maxi, uint32_t rori) { return tmin<uint32_t>(idx & maxi, (idx & maxi) ^ rori); }

// Per-row state of the float gradients: computed once per (command, row) by the lane that owns the row.
struct RadialRow { float b, d, dd; };
struct ConicRow { float tx, ay, by; };

B2D_HD RadialRow radial_row(const b2dgpu_gradient_radial& r, uint32_t y) {
  double yd = double(int32_t(y));
  double ptx = r.yx * yd + r.tx;
  double pty = r.yy * yd + r.ty;
  double b = yd * r.by + r.b0;
  double sq_dist = ptx * ptx + pty * pty;
  RadialRow o;
  o.b = float(b * r.inv2a);
  o.d = float((r.amul4 * (sq_dist - r.sq_fr) + b * b) * r.sq_inv2a);
  o.dd = float((yd * r.ddy + r.dd0) * r.sq_inv2a);
  return o;
}

B2D_HD uint32_t radial_index(const b2dgpu_gradient_radial& r, const RadialRow& row, uint32_t x) {
  float xf = float(int32_t(x));
  float sq_x = xf * xf;
  float a = sqrtf(f32_abs(sq_x * r.f32_ddd + (xf * row.dd + row.d)));
  float v = (xf * r.f32_bd + row.b) + a;
  return uint32_t(x86_trunc_f32(v));
}

// NOTE (documented +-1 LSB case): the reference accumulates the row origin `_tp += _yy_yx` from the first row of
// each fill call / band; the closed form below differs from that running sum by at most a few ulps of a double before
// it is rounded to float.
B2D_HD ConicRow conic_row(const b2dgpu_gradient_conic& c, uint32_t y) {
  double yd = double(int(y));
  double tpx = c.tx + c.yx * yd;
  double tpy = c.ty + c.yy * yd;
  ConicRow o;
  o.tx = float(tpx);
  float ay = float(tpy);
  o.by = (f32_bits(ay) >> 31) ? c.n_div_1_2_4[0] : 0.0f;
  o.ay = f32_abs(ay);
  return o;
}

B2D_HD uint32_t conic_index(const b2dgpu_gradient_conic& c, const ConicRow& row, uint32_t x) {
  float xf = float(int(x));
  float xv = xf * c.xx + row.tx;
  float ax = f32_abs(xv);
  float xy_min = tmin(ax, row.ay);
  float xy_max = tmax(ax, row.ay);
  float s = (ax == xy_min) ? c.n_div_1_2_4[2] : 0.0f;
  float p = xy_min / xy_max;
  float p_sq = p * p;
  float v = p_sq * c.q_coeff[3] + c.q_coeff[2];
  v = v * p_sq + c.q_coeff[1];
  v = v * p_sq + c.q_coeff[0];
  v = f32_abs(v * p + (-s));
  v = f32_abs(v - ((f32_bits(xv) >> 31) ? c.n_div_1_2_4[1] : 0.0f));
  v = f32_abs(v - row.by) + c.offset;
  return uint32_t(tmin<int32_t>(x86_nearby_f32(v), int32_t(c.maxi))) & c.rori;
}

// ---------------------------------------------------------------------------------------------------------------
// Patterns.
// ---------------------------------------------------------------------------------------------------------------
B2D_HD uint32_t load_src_pixel(const uint8_t* row, uint32_t x, uint32_t src_format) {
  if (src_format == B2DGPU_FORMAT_A8) return adapt_src_a8(row[x]);
  uint32_t p = reinterpret_cast<const uint32_t*>(row)[x];
  return src_format == B2DGPU_FORMAT_XRGB32 ? adapt_src_xrgb32(p) : p;
}

// FetchPatternVertAAExtendCtxAny::init (:94-149) as a function of y: index of the source row.
B2D_HD int pattern_row(const b2dgpu_fetch_pattern& p, uint32_t y) {
  int64_t yy = int64_t(y) + int64_t(p.simple.ty);
  int64_t h = p.src.h;
  int64_t ry = p.simple.ry;
  if (ry == 0) return int(tclamp<int64_t>(yy, 0, h - 1));
  yy = int64_t(uint32_t(yy) % uint32_t(ry));
  return int(yy >= h ? (h - 1) - (yy - h) : yy);
}

// Horizontal extend contexts (:186-352) as functions of x: index of the source pixel.
B2D_HD uint32_t pattern_col_pad(const b2dgpu_fetch_pattern& p, uint32_t x) {
  int64_t v = int64_t(x) + int64_t(p.simple.tx);
  return uint32_t(tclamp<int64_t>(v, 0, int64_t(p.src.w) - 1));
}
B2D_HD uint32_t pattern_col_repeat(const b2dgpu_fetch_pattern& p, uint32_t x) {
  uint64_t v = uint64_t(x) + uint64_t(int64_t(p.simple.tx));
  return uint32_t(v % uint64_t(int64_t(p.src.w)));
}
B2D_HD uint32_t pattern_col_ror(const b2dgpu_fetch_pattern& p, uint32_t x) {
  uint64_t rx = uint64_t(int64_t(p.simple.rx));
  int64_t v = int64_t((uint64_t(x) + uint64_t(int64_t(p.simple.tx))) % rx);
  if (v >= int64_t(p.src.w)) v -= int64_t(rx);
  return uint32_t(v ^ (v >> 63));
}

B2D_HD uint32_t pattern_col(const b2dgpu_fetch_pattern& p, uint32_t x, uint32_t mode /*0 pad, 1 repeat, 2 ror*/) {
  return mode == 0 ? pattern_col_pad(p, x) : mode == 1 ? pattern_col_repeat(p, x) : pattern_col_ror(p, x);
}

B2D_HD uint32_t fetch_pattern_aligned(const b2dgpu_fetch_pattern& p, uint32_t src_format, uint32_t mode, uint32_t x, uint32_t y) {
  const uint8_t* row = p.src.pixel_data + intptr_t(pattern_row(p, y)) * p.src.stride;
  return load_src_pixel(row, pattern_col(p, x, mode), src_format);
}

// FetchPatternAlignedBlit (:470-537): plain translation, the fill box is guaranteed to be inside the source.
B2D_HD uint32_t fetch_pattern_blit(const b2dgpu_fetch_pattern& p, uint32_t src_format, uint32_t x, uint32_t y) {
  const uint8_t* row = p.src.pixel_data + intptr_t(int32_t(y - uint32_t(p.simple.ty))) * p.src.stride;
  return load_src_pixel(row, x - uint32_t(p.simple.tx), src_format);
}

// FetchPatternFxFyAny (:601-693): fixed 2x2 weights, two-row / two-column footprint.
B2D_HD uint32_t fetch_pattern_fxfy(const b2dgpu_fetch_pattern& p, uint32_t src_format, uint32_t mode, uint32_t x, uint32_t y) {
  const uint8_t* row0 = p.src.pixel_data + intptr_t(pattern_row(p, y)) * p.src.stride;
  const uint8_t* row1 = p.src.pixel_data + intptr_t(pattern_row(p, y + 1)) * p.src.stride;
  uint32_t c0 = pattern_col(p, x, mode);
  uint32_t c1 = pattern_col(p, x + 1, mode);
  Lanes2 acc = add(mul(unpack(load_src_pixel(row0, c0, src_format)), p.simple.wa),
                   mul(unpack(load_src_pixel(row1, c0, src_format)), p.simple.wc));
  Lanes2 cur = add(mul(unpack(load_src_pixel(row0, c1, src_format)), p.simple.wb),
                   mul(unpack(load_src_pixel(row1, c1, src_format)), p.simple.wd));
  return pack(div256(add(cur, acc)));
}

// FetchPatternAffineCtx (:354-465).  px/py are 32.32 fixed point; `normalize` is what spanStartX() applies and what
// advance_x() maintains incrementally (one overflow correction per step keeps the same canonical range).
B2D_HD uint64_t affine_normalize(uint64_t v, int32_t tw, int32_t r, int32_t o) {
  uint32_t x = uint32_t(int32_t(v >> 32) % tw);
  if (int32_t(x) < 0) x += uint32_t(r);
  if (int32_t(x) > o) x -= uint32_t(r);
  return (uint64_t(x) << 32) | (v & 0xFFFFFFFFu);
}

struct AffinePos { uint64_t px, py; };

B2D_HD AffinePos affine_pos(const b2dgpu_pattern_affine& a, uint32_t x, uint32_t y) {
  uint64_t tx = a.tx.u64 + a.yx.u64 * uint64_t(y) + a.xx.u64 * uint64_t(x);
  uint64_t ty = a.ty.u64 + a.yy.u64 * uint64_t(y) + a.xy.u64 * uint64_t(x);
  AffinePos o;
  o.px = affine_normalize(tx, int32_t(a.tw), int32_t(a.rx.u64 >> 32), int32_t(a.ox.u64 >> 32));
  o.py = affine_normalize(ty, int32_t(a.th), int32_t(a.ry.u64 >> 32), int32_t(a.oy.u64 >> 32));
  return o;
}

B2D_HD uint32_t affine_fold(int32_t v, int32_t vmin, int32_t vmax, int32_t cor) {
  v = tmax(v, vmin);
  if (v > vmax) v = cor;
  return uint32_t(v ^ (v >> 31));
}

B2D_HD uint32_t fetch_pattern_affine_nn(const b2dgpu_fetch_pattern& p, uint32_t src_format, uint32_t x, uint32_t y) {
  const b2dgpu_pattern_affine& a = p.affine;
  AffinePos pos = affine_pos(a, x, y);
  uint32_t ix = affine_fold(int32_t(pos.px >> 32), a.min_x, a.max_x, a.cor_x);
  uint32_t iy = affine_fold(int32_t(pos.py >> 32), a.min_y, a.max_y, a.cor_y);
  return load_src_pixel(p.src.pixel_data + intptr_t(iy) * p.src.stride, ix, src_format);
}

B2D_HD uint32_t fetch_pattern_affine_bi(const b2dgpu_fetch_pattern& p, uint32_t src_format, uint32_t x, uint32_t y) {
  const b2dgpu_pattern_affine& a = p.affine;
  AffinePos pos = affine_pos(a, x, y);
  int32_t xi = int32_t(pos.px >> 32), yi = int32_t(pos.py >> 32);
  uint32_t x0 = affine_fold(xi, a.min_x, a.max_x, a.cor_x);
  uint32_t y0 = affine_fold(yi, a.min_y, a.max_y, a.cor_y);
  uint32_t x1 = affine_fold(xi + 1, a.min_x, a.max_x, a.cor_x);
  uint32_t y1 = affine_fold(yi + 1, a.min_y, a.max_y, a.cor_y);
  uint32_t wx = uint32_t(pos.px & 0xFFFFFFFFu) >> 24;
  uint32_t wy = uint32_t(pos.py & 0xFFFFFFFFu) >> 24;
  uint32_t ix = 256u - wx, iy = 256u - wy;

  const uint8_t* line0 = p.src.pixel_data + intptr_t(y0) * p.src.stride;
  const uint8_t* line1 = p.src.pixel_data + intptr_t(y1) * p.src.stride;

  Lanes2 p0 = add(mul(unpack(load_src_pixel(line0, x0, src_format)), iy), mul(unpack(load_src_pixel(line1, x0, src_format)), wy));
  Lanes2 p1 = add(mul(unpack(load_src_pixel(line0, x1, src_format)), iy), mul(unpack(load_src_pixel(line1, x1, src_format)), wy));
  p0 = mul(div256(p0), ix);
  p1 = mul(div256(p1), wx);
  return pack(div256(add(p0, p1)));
}

// ---------------------------------------------------------------------------------------------------------------
// Row context + per-pixel fetch.  `RowCtx` is what a lane keeps for the 4 pixels it owns in one row.
// ---------------------------------------------------------------------------------------------------------------
struct RowCtx {
  RadialRow radial;
  ConicRow conic;
};

B2D_HD void fetch_row_init(const FetchEnv& env, uint32_t y, RowCtx& rc) {
  uint32_t ft = env.fetch_type;
  if (ft >= B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD && ft <= B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_ROR)
    rc.radial = radial_row(env.fd->gradient.radial, y);
  else if (ft >= B2DGPU_FETCH_GRADIENT_CONIC_NN)
    rc.conic = conic_row(env.fd->gradient.conic, y);
}

B2D_HD uint32_t fetch_pattern_pixel(const b2dgpu_fetch_pattern& p, uint32_t ft, uint32_t src_format, uint32_t x, uint32_t y) {
  switch (ft) {
    case B2DGPU_FETCH_PATTERN_ALIGNED_BLIT:   return fetch_pattern_blit(p, src_format, x, y);
    case B2DGPU_FETCH_PATTERN_ALIGNED_PAD:    return fetch_pattern_aligned(p, src_format, 0, x, y);
    case B2DGPU_FETCH_PATTERN_ALIGNED_REPEAT: return fetch_pattern_aligned(p, src_format, 1, x, y);
    case B2DGPU_FETCH_PATTERN_ALIGNED_ROR:    return fetch_pattern_aligned(p, src_format, 2, x, y);
    case B2DGPU_FETCH_PATTERN_FX_PAD:
    case B2DGPU_FETCH_PATTERN_FY_PAD:
    case B2DGPU_FETCH_PATTERN_FXFY_PAD:       return fetch_pattern_fxfy(p, src_format, 0, x, y);
    case B2DGPU_FETCH_PATTERN_FX_ROR:
    case B2DGPU_FETCH_PATTERN_FY_ROR:
    case B2DGPU_FETCH_PATTERN_FXFY_ROR:       return fetch_pattern_fxfy(p, src_format, 2, x, y);
    case B2DGPU_FETCH_PATTERN_AFFINE_NN_ANY:
    case B2DGPU_FETCH_PATTERN_AFFINE_NN_OPT:  return fetch_pattern_affine_nn(p, src_format, x, y);
    default:                                  return fetch_pattern_affine_bi(p, src_format, x, y);
  }
}

B2D_HD uint32_t fetch_pixel(const FetchEnv& env, const RowCtx& rc, uint32_t x, uint32_t y) {
  const uint32_t ft = env.fetch_type;
  if (ft == B2DGPU_FETCH_SOLID) return env.solid;

  if (ft >= B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD) {
    const b2dgpu_fetch_gradient& g = env.fd->gradient;
    uint32_t idx;
    bool dither;
    if (ft <= B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_ROR) {
      const b2dgpu_gradient_linear& l = g.linear;
      uint64_t pt = l.pt[0].u64 + uint64_t(y) * l.dy.u64 + uint64_t(x) * l.dt.u64;
      idx = uint32_t(pt >> 32);
      bool pad = (ft == B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD) || (ft == B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_PAD);
      idx = pad ? grad_index_pad(idx, l.maxi) : grad_index_ror(idx, l.maxi, l.rori);
      dither = ft >= B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_PAD;
    }
    else if (ft <= B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_ROR) {
      const b2dgpu_gradient_radial& r = g.radial;
      idx = radial_index(r, rc.radial, x);
      bool pad = (ft == B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD) || (ft == B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_PAD);
      idx = pad ? grad_index_pad(idx, r.maxi) : grad_index_ror(idx, r.maxi, r.rori);
      dither = ft >= B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_PAD;
    }
    else {
      idx = conic_index(g.conic, rc.conic, x);
      dither = ft == B2DGPU_FETCH_GRADIENT_CONIC_DITHER;
    }
    return dither ? lut_fetch_dither(env, g, idx, x, y) : lut_fetch_nn(g, idx);
  }

  return fetch_pattern_pixel(env.fd->pattern, ft, env.src_format, x, y);
}

// Everything that is not a solid colour or a nearest-neighbour gradient (dithered gradients, all patterns): kept out of
// line so that the compositor's hot loop stays small (see B2D_HD_COLD).
// Every argument is a scalar: a FetchEnv passed by reference would have to live in local memory in the CALLER's hot
// loop (ncu round 2: five STL per replayed command in k_tile_render just to keep it addressable).
B2D_HD_COLD uint32_t fetch_pixel_cold(const b2dgpu_fetch_data* fd, const uint8_t* bayer, uint32_t type_and_format, int origin_x, int origin_y,
                                      uint32_t x, uint32_t y) {
  FetchEnv env;
  env.fd = fd; env.bayer = bayer;
  env.fetch_type = type_and_format & 0xFFu; env.src_format = type_and_format >> 8;
  env.solid = 0;
  env.origin_x = origin_x; env.origin_y = origin_y;
  RowCtx rc;
  fetch_row_init(env, y, rc);
  return fetch_pixel(env, rc, x, y);
}

// Per-row state of the nearest-neighbour gradient fetchers as three 32-bit words, so that it can be computed once per
// (command, row) and kept in shared memory: linear = the 64-bit row origin pt0 + y * dy (a, b), radial = RadialRow,
// conic = ConicRow (float bits).
struct RowCtx3 { uint32_t a, b, c; };

B2D_HD float f32_from_bits(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }

B2D_HD RowCtx3 fetch_row_ctx(uint32_t ft, const b2dgpu_fetch_data& fd, uint32_t y) {
  RowCtx3 o; o.a = o.b = o.c = 0;
  if (ft <= B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_ROR) {
    const uint64_t pt = fd.gradient.linear.pt[0].u64 + uint64_t(y) * fd.gradient.linear.dy.u64;
    o.a = uint32_t(pt); o.b = uint32_t(pt >> 32);
  }
  else if (ft <= B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_ROR) {
    const RadialRow r = radial_row(fd.gradient.radial, y);
    o.a = f32_bits(r.b); o.b = f32_bits(r.d); o.c = f32_bits(r.dd);
  }
  else {
    const ConicRow r = conic_row(fd.gradient.conic, y);
    o.a = f32_bits(r.tx); o.b = f32_bits(r.ay); o.c = f32_bits(r.by);
  }
  return o;
}

// Fetches the (up to) 4 consecutive pixels x..x+3 of row y whose mask is non-zero.  `rowctx` (optional) = the result of
// fetch_row_ctx() for this command and row.  The fetch-type dispatch is hoisted
// out of the pixel loop: a command is uniform over the whole CTA, so every warp takes the same branch.
B2D_HD void fetch4(const FetchEnv& env, uint32_t x, uint32_t y, const uint32_t* m, uint32_t* s, const RowCtx3* rowctx = nullptr) {
  const uint32_t ft = env.fetch_type;
  if (ft == B2DGPU_FETCH_SOLID) {
    s[0] = s[1] = s[2] = s[3] = env.solid;
    return;
  }
  // Nearest-neighbour gradients are evaluated for all four pixels without looking at the masks: the table index is
  // always in range (pad clamps, repeat / reflect mask it), so the loads are safe, and straight-line code with four
  // independent dependency chains is what the instruction front end and the LSU want.  The compositor ignores the
  // result where the mask is zero.
  if (ft == B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD || ft == B2DGPU_FETCH_GRADIENT_LINEAR_NN_ROR) {
    const b2dgpu_fetch_gradient& g = env.fd->gradient;
    const b2dgpu_gradient_linear& l = g.linear;
    const bool pad = ft == B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD;
    const uint32_t maxi = l.maxi, rori = l.rori;
    const uint64_t dt = l.dt.u64;
    uint64_t pt = (rowctx ? (uint64_t(rowctx->a) | (uint64_t(rowctx->b) << 32)) : l.pt[0].u64 + uint64_t(y) * l.dy.u64) + uint64_t(x) * dt;
    #pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t idx = uint32_t(pt >> 32);
      idx = pad ? grad_index_pad(idx, maxi) : grad_index_ror(idx, maxi, rori);
      s[i] = lut_fetch_nn(g, idx);
      pt += dt;
    }
    return;
  }
  if (ft == B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD || ft == B2DGPU_FETCH_GRADIENT_RADIAL_NN_ROR) {
    const b2dgpu_fetch_gradient& g = env.fd->gradient;
    const b2dgpu_gradient_radial& r = g.radial;
    const bool pad = ft == B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD;
    RadialRow row;
    if (rowctx) { row.b = f32_from_bits(rowctx->a); row.d = f32_from_bits(rowctx->b); row.dd = f32_from_bits(rowctx->c); }
    else row = radial_row(r, y);
    #pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t idx = radial_index(r, row, x + i);
      idx = pad ? grad_index_pad(idx, r.maxi) : grad_index_ror(idx, r.maxi, r.rori);
      s[i] = lut_fetch_nn(g, idx);
    }
    return;
  }
  if (ft == B2DGPU_FETCH_GRADIENT_CONIC_NN) {
    const b2dgpu_fetch_gradient& g = env.fd->gradient;
    ConicRow row;
    if (rowctx) { row.tx = f32_from_bits(rowctx->a); row.ay = f32_from_bits(rowctx->b); row.by = f32_from_bits(rowctx->c); }
    else row = conic_row(g.conic, y);
    #pragma unroll
    for (int i = 0; i < 4; i++) s[i] = lut_fetch_nn(g, conic_index(g.conic, row, x + i));
    return;
  }
  if (ft <= B2DGPU_FETCH_PATTERN_AFFINE_BI_OPT) {
    // Patterns (the aligned blit assumes the pixel is inside the source, so the masks guard the fetch).
    const b2dgpu_fetch_pattern& p = env.fd->pattern;
    const uint32_t sf = env.src_format;
    #pragma unroll
    for (int i = 0; i < 4; i++)
      if (m[i]) s[i] = fetch_pattern_pixel(p, ft, sf, x + i, y);
    return;
  }
  #pragma unroll
  for (int i = 0; i < 4; i++)
    if (m[i]) s[i] = fetch_pixel_cold(env.fd, env.bayer, ft | (env.src_format << 8), env.origin_x, env.origin_y, x + i, y);
}

} // namespace b2d
