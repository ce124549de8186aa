/*
 * b2d_oracle.c - CPU restatement (plain C) of the reference's pixel path, written for CHECKING, not for speed.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, load or call
 * this file.  Nothing under blend2d_b200/ does; the product has no CPU path.
 *
 * What is restated (each function cites the reference lines it follows, relative to /root/reference):
 *   - composition operators on premultiplied PRGB32 and on A8      pipeline/reference/compopgeneric_p.h:24-81,
 *                                                                  pipeline/reference/pixelgeneric_p.h:292-405
 *     plus Multiply / Screen, which only exist in the JIT          pipeline/jit/compoppart.cpp:4325-4461, 4591-4640
 *   - coverage accumulator -> mask                                 pipeline/reference/fillgeneric_p.h:381-387
 *   - FillAnalytic scanline walk over dense cells                  pipeline/reference/fillgeneric_p.h:174-379
 *   - axis-unaligned box -> mask-command program and its walker    pipeline/pipedefs_p.h:644-815,
 *                                                                  pipeline/reference/fillgeneric_p.h:67-162
 *   - the analytic rasterizer in its ORIGINAL whole-edge (non-banded, multi-scanline) form
 *                                                                  raster/analyticrasterizer_p.h:289-360, 466-1165
 *   - unclipped polygon -> 24.8 edges                              raster/edgebuilder_p.h:1136-1152
 *   - linear gradient fetch (incremental form)                     pipeline/reference/fetchgeneric_p.h:939-1011
 *
 * PINNING: every restated piece that the reference itself can execute (SrcOver, SrcCopy, masks, rasterizer, BoxU,
 * linear gradient) is checked against the UNMODIFIED reference binary (oracle/_ref/libblend2d_ref.so) in
 * tests/test_oracle.py and against the committed fixtures in tests/golden/.  The reference ships no golden images
 * (SURVEY.md section 4), so "outputs of the reference itself run here" are the golden vectors.
 * Plus / Multiply / Screen: PARITY UNPINNED - the reference's portable pipeline rejects them
 * (pipeline/reference/fixedpiperuntime.cpp:254) and its JIT cannot be built here (asmjit is not vendored); they are
 * restated from the JIT source and only their inputs (masks, source pixels) come from the reference binary.
 *
 * NOT restated here: curve flattening / clipping and the radial, conic and pattern fetchers.  Those are checked
 * directly against the reference binary (tests/test_parity_gpu.py, tests/test_hostsim_parity.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------------
 * Pixel arithmetic: 4 x 16-bit lanes in one 64-bit word, laid out [.3.1.2.0] like U32_8888 (pixelgeneric_p.h:292-296).
 * ---------------------------------------------------------------------------------------------------------------- */
static uint64_t orc_rep(uint32_t v) { uint64_t x = v | (v << 16); return x | (x << 32); }                 /* Repeat::u16x4 */
static uint64_t orc_unpack(uint32_t p) { return (uint64_t)(p & 0x00FF00FFu) | ((uint64_t)(p & 0xFF00FF00u) << 24); }
static uint32_t orc_pack(uint64_t u) { return (uint32_t)(((u >> 24) | u) & 0xFFFFFFFFu); }
static uint64_t orc_div255(uint64_t u) {                                                                  /* :390-393 */
  u += orc_rep(0x80u);
  return ((u + ((u >> 8) & orc_rep(0xFFu))) >> 8) & orc_rep(0xFFu);
}
static uint64_t orc_addus8(uint64_t a, uint64_t b) {                                                       /* :399-403 */
  uint64_t val = a + b;
  uint64_t msk = ((val >> 8) & orc_rep(0x1u)) * 0xFFu;
  return (val | msk) & orc_rep(0xFFu);
}

/* CompOp_SrcCopy_Op::op_prgb32_prgb32(d, s, m)                                      compopgeneric_p.h:38-40 */
static uint32_t orc_src_copy(uint32_t d, uint32_t s, uint32_t m) {
  return orc_pack(orc_div255(orc_unpack(d) * (m ^ 0xFFu) + orc_unpack(s) * m));
}
/* CompOp_SrcOver_Op::op_prgb32_prgb32(d, s, m)                                      compopgeneric_p.h:54-62 */
static uint32_t orc_src_over(uint32_t d, uint32_t s, uint32_t m) {
  uint32_t sm = orc_pack(orc_div255(orc_unpack(s) * m));
  return sm + orc_pack(orc_div255(orc_unpack(d) * ((sm >> 24) ^ 0xFFu)));
}
/* CompOp_Plus_Op::op_prgb32_prgb32(d, s, m)                                         compopgeneric_p.h:78-80 */
static uint32_t orc_plus(uint32_t d, uint32_t s, uint32_t m) {
  return orc_pack(orc_addus8(orc_unpack(d), orc_div255(orc_unpack(s) * m)));
}

/* 16-bit SIMD lane helpers of the JIT: v_mul_u16 (pmullw), v_add_i16 (paddw), v_div255_u16 = paddw 0x80; pmulhuw 0x101
 * (pipeline/jit/pipecompiler_p.h:84-95), packing with unsigned saturation (packuswb). */
static uint32_t jit_mul16(uint32_t a, uint32_t b) { return (a * b) & 0xFFFFu; }
static uint32_t jit_add16(uint32_t a, uint32_t b) { return (a + b) & 0xFFFFu; }
static uint32_t jit_div255(uint32_t x) { return (((x + 0x80u) & 0xFFFFu) * 0x101u) >> 16; }
static uint32_t jit_packus(uint32_t x) { int16_t v = (int16_t)x; return v < 0 ? 0u : v > 255 ? 255u : (uint32_t)v; }

/* Multiply, masked, Da and Sa used                                                  compoppart.cpp:4408-4441
 *   S = div255(S * m);  D' = div255(D * (S + (255 - Sa)) + S * (255 - Da))   per channel, alpha included. */
static uint32_t orc_multiply(uint32_t d, uint32_t s, uint32_t m) {
  uint32_t sv[4], dv[4], out = 0;
  for (int i = 0; i < 4; i++) { sv[i] = jit_div255(jit_mul16((s >> (i * 8)) & 0xFFu, m)); dv[i] = (d >> (i * 8)) & 0xFFu; }
  uint32_t isa = 255u - sv[3], ida = 255u - dv[3];                                                        /* v_inv255_u16 */
  for (int i = 0; i < 4; i++) {
    uint32_t y = jit_add16(isa, sv[i]);
    uint32_t v = jit_add16(jit_mul16(dv[i], y), jit_mul16(ida, sv[i]));
    out |= jit_packus(jit_div255(v)) << (i * 8);
  }
  return out;
}
/* Screen, masked                                                                    compoppart.cpp:4614-4630
 *   S = div255(S * m);  D' = div255(D * (255 - S)) + S */
static uint32_t orc_screen(uint32_t d, uint32_t s, uint32_t m) {
  uint32_t out = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t sc = jit_div255(jit_mul16((s >> (i * 8)) & 0xFFu, m));
    uint32_t dc = (d >> (i * 8)) & 0xFFu;
    out |= jit_packus(jit_add16(jit_div255(jit_mul16(dc, 255u - sc)), sc)) << (i * 8);
  }
  return out;
}

/* The remaining operators of the JIT for a PRGB32 destination, masked variants (`has_mask`), restated as the SEQUENCE of
 * vector instructions the JIT emits (compoppart.cpp v_mask_proc_rgba32_vec), on four 16-bit lanes b, g, r, a.
 * UNPINNED: the reference build here has no JIT (asmjit is not vendored) and its portable pipeline has none of them.
 *   sv / dv = source / destination channels, vm = mask in every lane, xv / yv = scratch registers like in the JIT. */
typedef struct { uint32_t l[4]; } v4;
static v4 v_load(uint32_t p) { v4 r; for (int i = 0; i < 4; i++) r.l[i] = (p >> (8 * i)) & 0xFFu; return r; }
static v4 v_set1(uint32_t x) { v4 r; for (int i = 0; i < 4; i++) r.l[i] = x & 0xFFFFu; return r; }
static v4 v_mul(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] = jit_mul16(a.l[i], b.l[i]); return a; }
static v4 v_add(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] = jit_add16(a.l[i], b.l[i]); return a; }
static v4 v_sub(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] = (a.l[i] - b.l[i]) & 0xFFFFu; return a; }
static v4 v_subs(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] > b.l[i] ? a.l[i] - b.l[i] : 0u; return a; }   /* psubusw */
static v4 v_d255(v4 a) { for (int i = 0; i < 4; i++) a.l[i] = jit_div255(a.l[i]); return a; }
static v4 v_inv255(v4 a) { for (int i = 0; i < 4; i++) a.l[i] = (a.l[i] ^ 0x00FFu) & 0xFFFFu; return a; }                 /* pxor 0x00FF */
static v4 v_alpha(v4 a) { return v_set1(a.l[3]); }                                                                        /* v_expand_alpha_16 */
static v4 v_zero_alpha(v4 a) { a.l[3] = 0; return a; }                                                                    /* vZeroAlphaW */
static v4 v_minu16(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] < b.l[i] ? a.l[i] : b.l[i]; return a; }
static v4 v_minmax_u8(v4 a, v4 b, int take_min) {
  for (int i = 0; i < 4; i++) {
    uint32_t r = 0;
    for (int k = 0; k < 16; k += 8) {
      uint32_t x = (a.l[i] >> k) & 0xFFu, y = (b.l[i] >> k) & 0xFFu;
      r |= (take_min ? (x < y ? x : y) : (x > y ? x : y)) << k;
    }
    a.l[i] = r;
  }
  return a;
}
static uint32_t v_packus(v4 a) { uint32_t p = 0; for (int i = 0; i < 4; i++) p |= jit_packus(a.l[i]) << (8 * i); return p; }

enum { ORC_SRC_IN = 2, ORC_SRC_OUT = 3, ORC_SRC_ATOP = 4, ORC_DST_OVER = 5, ORC_DST_IN = 7, ORC_DST_OUT = 8, ORC_DST_ATOP = 9, ORC_XOR = 10,
       ORC_MINUS = 13, ORC_MODULATE = 14, ORC_OVERLAY = 17, ORC_DARKEN = 18, ORC_LIGHTEN = 19, ORC_COLOR_DODGE = 20, ORC_COLOR_BURN = 21,
       ORC_LINEAR_BURN = 22, ORC_LINEAR_LIGHT = 23, ORC_PIN_LIGHT = 24, ORC_HARD_LIGHT = 25, ORC_SOFT_LIGHT = 26, ORC_DIFFERENCE = 27, ORC_EXCLUSION = 28 };

static uint32_t orc_jit_ext(uint32_t op, uint32_t d, uint32_t s, uint32_t m) {
  v4 sv = v_load(s), dv = v_load(d), vm = v_set1(m), vn = v_inv255(vm), xv, yv, uv;
  switch (op) {
    case ORC_SRC_IN:                                                                 /* compoppart.cpp:3751-3774 */
      xv = v_alpha(dv); xv = v_mul(xv, sv); xv = v_d255(xv); xv = v_mul(xv, vm);
      dv = v_mul(dv, vn); dv = v_add(dv, xv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_SRC_OUT:                                                                /* :3801-3826 */
      xv = v_alpha(dv); xv = v_inv255(xv); xv = v_mul(xv, sv); xv = v_d255(xv); xv = v_mul(xv, vm);
      dv = v_mul(dv, vn); dv = v_add(dv, xv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_SRC_ATOP:                                                               /* :3859-3879 */
      sv = v_mul(sv, vm); sv = v_d255(sv);
      xv = v_alpha(sv); xv = v_inv255(xv); yv = v_alpha(dv);
      dv = v_mul(dv, xv); yv = v_mul(yv, sv); dv = v_add(dv, yv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_DST_OVER:                                                               /* :3919-3937: d.ui * S.m, packed, + d.pc */
      sv = v_mul(sv, vm); sv = v_d255(sv);
      xv = v_inv255(v_alpha(dv)); xv = v_mul(xv, sv); xv = v_d255(xv);
      return v_packus(xv) + d;
    case ORC_DST_IN:                                                                 /* :3969-3983: s.ui = 255 - Sa */
      uv = v_inv255(v_alpha(sv)); uv = v_mul(uv, vm); uv = v_d255(uv); uv = v_inv255(uv);
      dv = v_mul(dv, uv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_DST_OUT:                                                                /* :4012-4026: s.ua = Sa */
      uv = v_alpha(sv); uv = v_mul(uv, vm); uv = v_d255(uv); uv = v_inv255(uv);
      dv = v_mul(dv, uv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_DST_ATOP:                                                               /* :4061-4085 */
      uv = v_inv255(v_alpha(sv));
      xv = v_alpha(dv);
      sv = v_mul(sv, vm); uv = v_mul(uv, vm); sv = v_d255(sv); uv = v_d255(uv);
      xv = v_inv255(xv); uv = v_inv255(uv);
      xv = v_mul(xv, sv); dv = v_mul(dv, uv); dv = v_add(dv, xv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_XOR:                                                                    /* :4119-4140 */
      sv = v_mul(sv, vm); sv = v_d255(sv);
      xv = v_inv255(v_alpha(sv)); yv = v_inv255(v_alpha(dv));
      dv = v_mul(dv, xv); sv = v_mul(sv, yv); dv = v_add(dv, sv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_MINUS:                                                                  /* :4229-4256 (use_da) */
      xv = v_alpha(dv); yv = dv; xv = v_inv255(xv);
      dv = v_subs(dv, sv); sv = v_mul(sv, xv); dv = v_zero_alpha(dv); sv = v_d255(sv);
      dv = v_add(dv, sv); dv = v_mul(dv, vm);
      vm = v_zero_alpha(vm); vm = v_inv255(vm); yv = v_mul(yv, vm);
      dv = v_add(dv, yv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_MODULATE:                                                               /* :4303-4313 */
      sv = v_mul(sv, vm); sv = v_d255(sv); sv = v_add(sv, v_set1(0x00FF)); sv = v_sub(sv, vm);
      dv = v_mul(dv, sv); dv = v_d255(dv);
      return v_packus(dv);
    case ORC_DARKEN:
    case ORC_LIGHTEN:                                                                /* :4650-4678 (has_mask => use_sa) */
      sv = v_mul(sv, vm); sv = v_d255(sv);
      xv = v_inv255(v_alpha(dv)); yv = v_inv255(v_alpha(sv));
      xv = v_mul(xv, sv); yv = v_mul(yv, dv); xv = v_d255(xv); yv = v_d255(yv);
      dv = v_add(dv, xv); sv = v_add(sv, yv);
      dv = v_minmax_u8(dv, sv, op == ORC_DARKEN);
      return v_packus(dv);
    case ORC_LINEAR_BURN:                                                            /* :4853-4870 */
      sv = v_mul(sv, vm); sv = v_d255(sv);
      xv = v_alpha(sv); yv = v_alpha(dv); xv = v_mul(xv, yv); xv = v_d255(xv);
      dv = v_add(dv, sv); dv = v_subs(dv, xv);
      return v_packus(dv);
    case ORC_DIFFERENCE:                                                             /* :5287-5314 */
      sv = v_mul(sv, vm); sv = v_d255(sv);
      yv = v_alpha(sv); xv = v_alpha(dv);
      yv = v_mul(yv, dv); xv = v_mul(xv, sv); dv = v_add(dv, sv); yv = v_minu16(yv, xv);
      yv = v_d255(yv); dv = v_sub(dv, yv); yv = v_zero_alpha(yv); dv = v_sub(dv, yv);
      return v_packus(dv);
    case ORC_EXCLUSION:                                                              /* :5326-5346 */
      sv = v_mul(sv, vm); sv = v_d255(sv);
      xv = v_mul(dv, sv); dv = v_add(dv, sv); xv = v_d255(xv);
      dv = v_sub(dv, xv); xv = v_zero_alpha(xv); dv = v_sub(dv, xv);
      return v_packus(dv);
    default:
      return d;
  }
}

/* More of the JIT's vector vocabulary, for Overlay / HardLight / PinLight / LinearLight (16-bit integer lanes) and the
 * scalar-float operators ColorDodge / ColorBurn / SoftLight (one pixel per iteration: four 32-bit float lanes b, g, r, a). */
static v4 v_slli1(v4 a) { for (int i = 0; i < 4; i++) a.l[i] = (a.l[i] << 1) & 0xFFFFu; return a; }
static v4 v_cmpgt_i16(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] = (int16_t)a.l[i] > (int16_t)b.l[i] ? 0xFFFFu : 0u; return a; }   /* pcmpgtw */
static v4 v_xor(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] ^= b.l[i]; return a; }
static v4 v_and(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] &= b.l[i]; return a; }
static v4 v_bic(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] &= ~b.l[i] & 0xFFFFu; return a; }                                       /* a & ~b */
static v4 v_fill_alpha(v4 a) { a.l[3] = 0xFFFFu; return a; }                                                                              /* por p_FFFF000000000000 */
static v4 v_min_i16(v4 a, v4 b) { for (int i = 0; i < 4; i++) a.l[i] = (int16_t)a.l[i] < (int16_t)b.l[i] ? a.l[i] : b.l[i]; return a; } /* pminsw */

typedef struct { float l[4]; } f4;
static f4 f_from(v4 a) { f4 r; for (int i = 0; i < 4; i++) r.l[i] = (float)(int32_t)a.l[i]; return r; }                                  /* cvtdq2ps */
static f4 f_set1(float x) { f4 r; for (int i = 0; i < 4; i++) r.l[i] = x; return r; }
static f4 f_alpha(f4 a) { return f_set1(a.l[3]); }                                                                                        /* vExpandAlphaPS */
static f4 f_mul(f4 a, f4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] * b.l[i]; return a; }
static f4 f_add(f4 a, f4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] + b.l[i]; return a; }
static f4 f_sub(f4 a, f4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] - b.l[i]; return a; }
static f4 f_div(f4 a, f4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] / b.l[i]; return a; }
static f4 f_sqrt(f4 a) { for (int i = 0; i < 4; i++) a.l[i] = sqrtf(a.l[i]); return a; }
static f4 f_max(f4 a, f4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] > b.l[i] ? a.l[i] : b.l[i]; return a; }                        /* maxps: the second operand unless a > b */
static f4 f_min(f4 a, f4 b) { for (int i = 0; i < 4; i++) a.l[i] = a.l[i] < b.l[i] ? a.l[i] : b.l[i]; return a; }                        /* minps */
static f4 f_neg(f4 a) { for (int i = 0; i < 4; i++) a.l[i] = -a.l[i]; return a; }                                                        /* xorps sign bit */
static f4 f_zero_alpha(f4 a) { a.l[3] = 0.0f; return a; }                                                                                 /* andps p_FFFFFFFF_FFFFFFFF_FFFFFFFF_0 */
static f4 f_sel(const int* mask, f4 a, f4 b) { for (int i = 0; i < 4; i++) a.l[i] = mask[i] ? a.l[i] : b.l[i]; return a; }                /* andps / andnps / orps */
static int32_t cvtt_f32(float v) { return (v >= -2147483648.0f && v < 2147483648.0f) ? (int32_t)v : INT32_MIN; }                          /* cvttps2dq */
static int32_t cvt_f32(float v) { return (v >= -2147483648.0f && v < 2147483648.0f) ? (int32_t)lrintf(v) : INT32_MIN; }                   /* cvtps2dq, round to nearest even */
static v4 f_trunc_pack_u16(f4 a) {                                                                                                         /* cvttps2dq + packusdw */
  v4 r;
  for (int i = 0; i < 4; i++) { int32_t x = cvtt_f32(a.l[i]); r.l[i] = x < 0 ? 0u : x > 65535 ? 65535u : (uint32_t)x; }
  return r;
}

/* ColorDodge / ColorBurn share their integer half: Dca.(1 - Sa) + Sca.(1 - Da) on the packed [Dca | Sca] register
 * (compoppart.cpp:4760-4766, 4819-4825). */
static v4 dodge_burn_int_part(v4 dv, v4 sv) {
  v4 isa = v_inv255(v_alpha(sv)), ida = v_inv255(v_alpha(dv));
  return v_add(v_mul(dv, isa), v_mul(sv, ida));
}

static uint32_t orc_jit_light(uint32_t op, uint32_t d, uint32_t s, uint32_t m) {
  v4 sv = v_load(s), dv = v_load(d), vm = v_set1(m), xv, yv, zv;
  sv = v_mul(sv, vm); sv = v_d255(sv);                                               /* has_mask: S = S.m */
  switch (op) {
    case ORC_OVERLAY:                                                                /* compoppart.cpp:4466-4540 (use_sa, use_da) */
      xv = v_alpha(dv); yv = v_alpha(sv);
      xv = v_mul(xv, sv); yv = v_mul(yv, dv); zv = v_mul(dv, sv);
      sv = v_add(sv, dv); xv = v_sub(xv, zv); zv = v_zero_alpha(zv); xv = v_add(xv, yv);
      yv = v_alpha(dv); xv = v_sub(xv, zv);
      dv = v_slli1(dv); yv = v_cmpgt_i16(yv, dv); xv = v_d255(xv); yv = v_fill_alpha(yv);
      zv = v_alpha(xv);
      xv = v_xor(xv, yv); xv = v_sub(xv, yv);
      yv = v_bic(zv, yv);
      sv = v_add(sv, xv); sv = v_sub(sv, yv);
      return v_packus(sv);
    case ORC_HARD_LIGHT:                                                             /* :5091-5147 */
      xv = v_alpha(dv); yv = v_alpha(sv);
      xv = v_mul(xv, sv); yv = v_mul(yv, dv); zv = v_mul(dv, sv);
      dv = v_add(dv, sv); xv = v_sub(xv, zv); xv = v_add(xv, yv); xv = v_sub(xv, zv);
      yv = v_alpha(yv); zv = v_alpha(sv); xv = v_d255(xv); yv = v_d255(yv);
      sv = v_slli1(sv); zv = v_cmpgt_i16(zv, sv);
      xv = v_xor(xv, zv); xv = v_sub(xv, zv); zv = v_zero_alpha(zv); zv = v_bic(yv, zv);
      dv = v_add(dv, xv); dv = v_sub(dv, zv);
      return v_packus(dv);
    case ORC_PIN_LIGHT:                                                              /* :4954-5001 (use_sa && use_da) */
      yv = v_alpha(sv); xv = v_alpha(dv);
      yv = v_mul(yv, dv); xv = v_mul(xv, sv); dv = v_add(dv, sv); yv = v_d255(yv); xv = v_d255(xv);
      yv = v_sub(yv, dv); dv = v_sub(dv, xv); xv = v_sub(xv, yv);
      yv = v_alpha(sv); sv = v_slli1(sv); sv = v_cmpgt_i16(sv, yv);
      zv = v_sub(dv, xv); zv = v_alpha(zv); zv = v_and(zv, sv); xv = v_add(xv, zv);
      dv = v_xor(dv, sv); xv = v_xor(xv, sv); dv = v_min_i16(dv, xv); dv = v_xor(dv, sv);
      return v_packus(dv);
    case ORC_LINEAR_LIGHT: {                                                         /* :4889-4935: [Dca | Sca] in one register */
      v4 d_lo = dv, d_hi = sv, x_lo = v_alpha(sv), x_hi = v_alpha(dv), s_lo, s_hi, y_lo, t;
      s_lo = d_lo; s_hi = d_hi;
      d_lo = v_mul(d_lo, x_lo); d_hi = v_mul(d_hi, x_hi);                            /* [Dca.Sa | Sca.Da] */
      x_lo = v_inv255(x_lo); x_hi = v_inv255(x_hi);
      d_lo = v_d255(d_lo); d_hi = v_d255(d_hi);
      s_lo = v_mul(s_lo, x_lo); s_hi = v_mul(s_hi, x_hi);                            /* [Dca.(1 - Sa) | Sca.(1 - Da)] */
      t = s_hi; y_lo = d_hi;                                                         /* the swapped halves (low part is all that is kept) */
      s_lo = v_add(s_lo, t);
      d_lo = v_add(d_lo, y_lo);
      xv = v_alpha(y_lo);                                                            /* Sa.Da */
      d_lo = v_add(d_lo, y_lo);
      s_lo = v_d255(s_lo);
      d_lo = v_subs(d_lo, xv); d_lo = v_min_i16(d_lo, xv);
      d_lo = v_add(d_lo, s_lo);
      return v_packus(d_lo);
    }
    case ORC_COLOR_DODGE: {                                                          /* :4724-4778 */
      f4 y0 = f_from(sv), z0 = f_from(dv), x0;
      x0 = f_alpha(y0); y0 = f_neg(y0); z0 = f_mul(z0, x0); y0 = f_zero_alpha(y0); y0 = f_add(y0, x0);
      y0 = f_max(y0, f_set1(1e-3f)); z0 = f_div(z0, y0);
      xv = dodge_burn_int_part(dv, sv);
      z0 = f_mul(z0, x0); x0 = f_alpha(z0); z0 = f_min(z0, x0);
      xv = v_add(xv, f_trunc_pack_u16(z0)); xv = v_d255(xv);
      return v_packus(xv);
    }
    case ORC_COLOR_BURN: {                                                           /* :4784-4841 */
      f4 y0 = f_from(sv), z0 = f_from(dv), x0;
      x0 = f_alpha(y0); y0 = f_max(y0, f_set1(1e-3f)); z0 = f_mul(z0, x0);
      x0 = f_alpha(z0); z0 = f_neg(z0); z0 = f_zero_alpha(z0); z0 = f_add(z0, x0); z0 = f_div(z0, y0);
      xv = dodge_burn_int_part(dv, sv);
      x0 = f_alpha(y0); z0 = f_mul(z0, x0); x0 = f_alpha(z0); z0 = f_min(z0, x0); z0 = f_zero_alpha(z0); x0 = f_sub(x0, z0);
      xv = v_add(xv, f_trunc_pack_u16(x0)); xv = v_d255(xv);
      return v_packus(xv);
    }
    case ORC_SOFT_LIGHT: {                                                           /* :5152-5247; v_madd_f32 as mul + add (no FMA in the baseline) */
      f4 s0 = f_from(sv), d0 = f_from(dv), x0 = f_set1(1.0f / 255.0f), a0, b0, y0, z0;
      int le[4], gt[4];
      uint32_t out = 0;
      s0 = f_mul(s0, x0); d0 = f_mul(d0, x0);
      b0 = f_alpha(d0); x0 = f_mul(s0, b0); b0 = f_max(b0, f_set1(1e-3f));
      a0 = f_div(d0, b0); d0 = f_add(d0, s0);
      y0 = f_alpha(s0);
      d0 = f_sub(d0, x0); s0 = f_add(s0, s0); z0 = f_mul(a0, f_set1(4.0f));
      x0 = f_sqrt(a0); s0 = f_sub(s0, y0);
      y0 = z0; z0 = f_add(f_mul(z0, a0), a0); s0 = f_mul(s0, b0);
      z0 = f_sub(z0, y0); b0 = f_set1(1.0f);
      z0 = f_add(z0, b0); z0 = f_mul(z0, y0);
      for (int i = 0; i < 4; i++) le[i] = y0.l[i] <= b0.l[i];
      z0 = f_sel(le, z0, x0);                                                        /* 4.Dc <= 1 ? polynomial : sqrt(Dc) */
      for (int i = 0; i < 4; i++) gt[i] = 0.0f < s0.l[i];
      z0 = f_sub(z0, a0); b0 = f_sub(b0, a0);
      b0 = f_mul(b0, a0);
      z0 = f_sel(gt, z0, b0);                                                        /* (2.Sca - Sa).Da > 0 ? [..] - Dc : Dc.(1 - Dc) */
      s0 = f_zero_alpha(s0);
      s0 = f_mul(s0, z0); d0 = f_add(d0, s0); d0 = f_mul(d0, f_set1(255.0f));
      for (int i = 0; i < 4; i++) {
        int32_t v = cvt_f32(d0.l[i]);
        v = v < -32768 ? -32768 : v > 32767 ? 32767 : v;                             /* packssdw */
        out |= (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v) << (8 * i);                /* packuswb */
      }
      return out;
    }
    default:
      return d;
  }
}

/* A8 pixels (P8_Alpha / U8_Alpha, pixelgeneric_p.h:85-200): one 16-bit lane, packed adds wrap at 8 bits. */
static uint32_t a8_div255(uint32_t u) { u = (u + 0x80u) & 0xFFFFu; return ((u + ((u >> 8) & 0xFFu)) >> 8) & 0xFFu; }
static uint32_t orc_a8_src_copy(uint32_t d, uint32_t s, uint32_t m) { return a8_div255(d * (m ^ 0xFFu) + s * m); }
static uint32_t orc_a8_src_over(uint32_t d, uint32_t s, uint32_t m) {
  uint32_t sm = a8_div255(s * m);
  return (sm + a8_div255(d * (sm ^ 0xFFu))) & 0xFFu;
}

enum { ORC_SRC_OVER = 0, ORC_SRC_COPY = 1, ORC_PLUS = 12, ORC_MULTIPLY = 15, ORC_SCREEN = 16 };

ORC_API uint32_t orc_composite_prgb32(uint32_t op, uint32_t d, uint32_t s, uint32_t m) {
  if (m == 0) return d;                      /* the fillers never call the compositor with a zero mask ... except in  */
  switch (op) {                              /* VMask runs, where every operator below is the identity for m == 0.    */
    case ORC_SRC_OVER: return orc_src_over(d, s, m);
    case ORC_SRC_COPY: return orc_src_copy(d, s, m);
    case ORC_PLUS: return orc_plus(d, s, m);
    case ORC_MULTIPLY: return orc_multiply(d, s, m);
    case ORC_SCREEN: return orc_screen(d, s, m);
    case ORC_OVERLAY: case ORC_COLOR_DODGE: case ORC_COLOR_BURN: case ORC_LINEAR_LIGHT: case ORC_PIN_LIGHT: case ORC_HARD_LIGHT: case ORC_SOFT_LIGHT:
      return orc_jit_light(op, d, s, m);
    default: return orc_jit_ext(op, d, s, m);
  }
}

/* dst[i] = op(dst[i], src[i], mask[i]) over a plane of n pixels; src_stride 0 means a solid source. */
ORC_API void orc_composite_plane_prgb32(uint32_t op, uint32_t* dst, const uint32_t* src, int src_is_solid, const uint8_t* mask, size_t n) {
  for (size_t i = 0; i < n; i++) dst[i] = orc_composite_prgb32(op, dst[i], src[src_is_solid ? 0 : i], mask[i]);
}

ORC_API void orc_composite_plane_a8(uint32_t op, uint8_t* dst, const uint8_t* src, int src_is_solid, const uint8_t* mask, size_t n) {
  for (size_t i = 0; i < n; i++) {
    uint32_t m = mask[i], s = src[src_is_solid ? 0 : i];
    if (!m) continue;
    dst[i] = (uint8_t)(op == ORC_SRC_COPY ? orc_a8_src_copy(dst[i], s, m) : orc_a8_src_over(dst[i], s, m));
  }
}

/* FillAnalytic_Base::calc_mask                                                      fillgeneric_p.h:381-387 */
ORC_API uint32_t orc_calc_mask(uint32_t cov, uint32_t fill_rule_mask, uint32_t global_alpha) {
  uint32_t c = 256;
  uint32_t m = ((uint32_t)(((int32_t)cov) >> 9) & fill_rule_mask) - c;
  int32_t mi = (int32_t)m;
  uint32_t a = (uint32_t)(mi < 0 ? -mi : mi);
  if (a > c) a = c;
  return (a * global_alpha) >> 8;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Analytic rasterizer, whole-edge form (rasterize<> without kOptionBandingMode).  Cells: u32 cells[h][stride],
 * stride >= w + 2.  The bit vectors of the reference only accelerate the later scan and are not modelled.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int ex0, ey0, ex1, ey1, fx0, fy0, fx1, fy1;
  int xErr, yErr, xDlt, yDlt, xRem, yRem, xLift, yLift, dx, dy, savedFy1;
  uint32_t flags, sign_mask;
} OrcRas;

enum { ORC_INITIAL = 1, ORC_VERT_OR_SINGLE = 2, ORC_RTL = 4 };

static void orc_err_step(int* acc, int* iter, int step, int correction) {                                 /* :71-82 */
  *iter -= step;
  if (*iter < 0) { (*acc)++; *iter += correction; }
}
static uint32_t orc_sign(const OrcRas* r, uint32_t v) { return (v ^ r->sign_mask) - r->sign_mask; }        /* :56-60 */
static void orc_merge(uint32_t* row, int x, uint32_t cover, uint32_t area) {                               /* :1202-1210 */
  row[x] += (cover << 9) - area;
  row[x + 1] += area;
}

static int orc_prepare(OrcRas* r, int x0, int y0, int x1, int y1) {                                        /* :289-360 */
  if (y0 == y1) return 0;
  r->dx = x1 - x0; r->dy = y1 - y0; r->flags = ORC_INITIAL;
  if (r->dx < 0) { r->flags |= ORC_RTL; r->dx = -r->dx; }
  r->ex0 = x0 >> 8; r->ey0 = y0 >> 8; r->ex1 = x1 >> 8; r->ey1 = (y1 - 1) >> 8;
  r->fx0 = x0 & 255; r->fy0 = y0 & 255; r->fx1 = x1 & 255; r->fy1 = ((y1 - 1) & 255) + 1;
  r->savedFy1 = r->fy1;
  if (r->ey0 != r->ey1) r->fy1 = 256;
  r->xErr = r->yErr = r->xDlt = r->yDlt = r->xRem = r->yRem = r->xLift = r->yLift = 0;
  if (r->ex0 == r->ex1 && (r->ey0 == r->ey1 || r->dx == 0)) { r->flags |= ORC_VERT_OR_SINGLE; return 1; }
  uint64_t x_base = (uint64_t)(uint32_t)r->dx * 256, y_base = (uint64_t)(uint32_t)r->dy * 256;
  r->xLift = (int)(x_base / (unsigned)r->dy); r->xRem = (int)(x_base % (unsigned)r->dy);
  r->yLift = (int)(y_base / (unsigned)r->dx); r->yRem = (int)(y_base % (unsigned)r->dx);
  r->xDlt = r->dx; r->yDlt = r->dy;
  r->xErr = (r->dy >> 1) - 1; r->yErr = (r->dx >> 1) - 1;
  if (r->ey0 != r->ey1) {
    uint64_t p = (uint64_t)(256 - (uint32_t)r->fy0) * (uint32_t)r->dx;
    r->xDlt = (int)(p / (unsigned)r->dy); r->xErr -= (int)(p % (unsigned)r->dy);
    orc_err_step(&r->xDlt, &r->xErr, 0, r->dy);
  }
  if (r->ex0 != r->ex1) {
    uint64_t p = (uint64_t)((r->flags & ORC_RTL) ? (uint32_t)r->fx0 : 256 - (uint32_t)r->fx0) * (uint32_t)r->dy;
    r->yDlt = (int)(p / (unsigned)r->dx); r->yErr -= (int)(p % (unsigned)r->dx);
    orc_err_step(&r->yDlt, &r->yErr, 0, r->dx);
  }
  r->yDlt += r->fy0;
  return 1;
}

/* One whole edge, top to bottom (rasterize<0>, :466-1165 with every kOptionBandingMode branch removed). */
static void orc_rasterize(OrcRas* r, uint32_t* cells, size_t stride) {
  size_t i = (size_t)(r->ey1 - r->ey0);
  uint32_t* row = cells + (size_t)r->ey0 * stride;
  const uint32_t full = orc_sign(r, 256);

  if (r->flags & ORC_VERT_OR_SINGLE) {                                                                      /* :493-571 */
    uint32_t area = (uint32_t)r->fx0 + (uint32_t)r->fx1;
    uint32_t cover = orc_sign(r, (uint32_t)(r->fy1 - r->fy0));
    orc_merge(row, r->ex0, cover, cover * area);
    if (!i) return;
    row += stride;
    cover = full;
    while (--i) { orc_merge(row, r->ex0, cover, cover * area); row += stride; }
    cover = orc_sign(r, (uint32_t)r->savedFy1);
    orc_merge(row, r->ex0, cover, cover * area);
    return;
  }

  if (r->dy >= r->dx) {                                                                                     /* :572-869 */
    for (;;) {
      uint32_t area = (uint32_t)r->fx0, cov;
      if (r->flags & ORC_RTL) {
        r->fx0 -= r->xDlt;
        if (r->fx0 < 0) {
          r->ex0--; r->fx0 += 256; r->yDlt &= 255;
          if (!area) {
            area = 256;
            orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); r->yDlt += r->yLift;
            cov = orc_sign(r, (uint32_t)(r->fy1 - r->fy0));
            orc_merge(row, r->ex0, cov, cov * (area + (uint32_t)r->fx0));
          }
          else {
            cov = orc_sign(r, (uint32_t)(r->yDlt - r->fy0));
            orc_merge(row, r->ex0 + 1, cov, cov * area);
            cov = orc_sign(r, (uint32_t)(r->fy1 - r->yDlt));
            orc_merge(row, r->ex0, cov, cov * ((uint32_t)r->fx0 + 256));
            orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); r->yDlt += r->yLift;
          }
        }
        else {
          cov = orc_sign(r, (uint32_t)(r->fy1 - r->fy0));
          orc_merge(row, r->ex0, cov, cov * (area + (uint32_t)r->fx0));
        }
      }
      else {
        r->fx0 += r->xDlt;
        if (r->fx0 <= 256) {
          cov = orc_sign(r, (uint32_t)(r->fy1 - r->fy0));
          orc_merge(row, r->ex0, cov, cov * (area + (uint32_t)r->fx0));
          if (r->fx0 == 256) { r->ex0++; r->fx0 = 0; r->yDlt += r->yLift; orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); }
        }
        else {
          r->ex0++; r->fx0 &= 255; r->yDlt &= 255;
          cov = orc_sign(r, (uint32_t)(r->yDlt - r->fy0));
          orc_merge(row, r->ex0 - 1, cov, cov * (area + 256));
          cov = orc_sign(r, (uint32_t)(r->fy1 - r->yDlt));
          orc_merge(row, r->ex0, cov, cov * (uint32_t)r->fx0);
          r->yDlt += r->yLift; orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx);
        }
      }
      r->fy0 = 0;
      row += stride;
      if (!i) return;

      /* scanlines strictly between the first and the last one: full cover, x advances by the DDA step */
      while (--i) {
        r->xDlt = r->xLift; orc_err_step(&r->xDlt, &r->xErr, r->xRem, r->dy);
        r->fy1 = 256;
        area = (uint32_t)r->fx0;
        if (r->flags & ORC_RTL) {
          r->fx0 -= r->xDlt;
          if (r->fx0 < 0) {
            r->ex0--; r->fx0 += 256; r->yDlt &= 255;
            if (!area) {
              area = 256;
              orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); r->yDlt += r->yLift;
              orc_merge(row, r->ex0, full, full * (area + (uint32_t)r->fx0));
            }
            else {
              uint32_t c1 = orc_sign(r, (uint32_t)r->yDlt);
              orc_merge(row, r->ex0 + 1, c1, c1 * area);
              uint32_t c0 = full - c1;
              orc_merge(row, r->ex0, c0, c0 * ((uint32_t)r->fx0 + 256));
              orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); r->yDlt += r->yLift;
            }
          }
          else orc_merge(row, r->ex0, full, full * (area + (uint32_t)r->fx0));
        }
        else {
          r->fx0 += r->xDlt;
          if (r->fx0 <= 256) {
            orc_merge(row, r->ex0, full, full * (area + (uint32_t)r->fx0));
            if (r->fx0 == 256) { r->ex0++; r->fx0 = 0; r->yDlt += r->yLift; orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); }
          }
          else {
            r->fx0 &= 255; r->yDlt &= 255;
            uint32_t c0 = orc_sign(r, (uint32_t)r->yDlt);
            orc_merge(row, r->ex0, c0, c0 * (area + 256));
            r->ex0++;
            uint32_t c1 = orc_sign(r, 256 - (uint32_t)r->yDlt);
            orc_merge(row, r->ex0, c1, c1 * (uint32_t)r->fx0);
            r->yDlt += r->yLift; orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx);
          }
        }
        row += stride;
      }

      /* prepare the last scanline (:862-866 / :717-722) and run the first-scanline code once more */
      r->fy1 = r->savedFy1;
      r->xDlt = (r->flags & ORC_RTL) ? ((r->ex0 - r->ex1) << 8) + r->fx0 - r->fx1 : ((r->ex1 - r->ex0) << 8) + r->fx1 - r->fx0;
      i = 0;
    }
  }

  /* shallow edge: a run of cells per scanline (:870-1164) */
  {
    size_t j = 1;
    int x_local = (r->ex0 << 8) + r->fx0;
    uint32_t cover, area;
    const int rtl = (r->flags & ORC_RTL) != 0;
    int skip_to_inside = 0;

    if (r->flags & ORC_INITIAL) {
      r->flags &= ~ORC_INITIAL;
      j = i; i = 1;
      cover = orc_sign(r, (uint32_t)(r->yDlt - r->fy0));
      if (rtl ? (r->fx0 - r->xDlt < 0) : (r->fx0 + r->xDlt > 256)) skip_to_inside = 1;
      else {
        if (rtl) x_local -= r->xDlt; else x_local += r->xDlt;
        cover = orc_sign(r, (uint32_t)(r->fy1 - r->fy0));
        area = rtl ? cover * (uint32_t)(r->fx0 * 2 - r->xDlt) : cover * ((uint32_t)r->fx0 * 2 + (uint32_t)r->xDlt);
        orc_merge(row, r->ex0, cover, area);
        if (rtl ? ((x_local & 255) == 0) : (r->fx0 + r->xDlt == 256)) { r->yDlt += r->yLift; orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); }
        r->xDlt = r->xLift; orc_err_step(&r->xDlt, &r->xErr, r->xRem, r->dy);
        row += stride;
        i--;
      }
    }
    else cover = 0;

    for (;;) {
      while (i) {
        if (!skip_to_inside) {
          if (rtl) { r->ex0 = (x_local - 1) >> 8; r->fx0 = ((x_local - 1) & 255) + 1; }
          else { r->ex0 = x_local >> 8; r->fx0 = x_local & 255; }
          r->yDlt -= 256;
          cover = orc_sign(r, (uint32_t)r->yDlt);
        }
        skip_to_inside = 0;

        if (rtl) {
          x_local -= r->xDlt;
          int ex_local = x_local >> 8, fx_local = x_local & 255;
          area = cover * (uint32_t)r->fx0;
          while (r->ex0 != ex_local) {
            orc_merge(row, r->ex0, cover, area);
            int cv = r->yLift; orc_err_step(&cv, &r->yErr, r->yRem, r->dx);
            r->yDlt += cv;
            cover = orc_sign(r, (uint32_t)cv);
            area = cover * 256;
            r->ex0--;
          }
          cover += orc_sign(r, (uint32_t)(r->fy1 - r->yDlt));
          area = cover * ((uint32_t)fx_local + 256);
          orc_merge(row, r->ex0, cover, area);
          if (fx_local == 0) { r->yDlt += r->yLift; orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); }
        }
        else {
          x_local += r->xDlt;
          int ex_local = (x_local - 1) >> 8, fx_local = ((x_local - 1) & 255) + 1;
          area = cover * ((uint32_t)r->fx0 + 256);
          while (r->ex0 != ex_local) {
            orc_merge(row, r->ex0, cover, area);
            int cv = r->yLift; orc_err_step(&cv, &r->yErr, r->yRem, r->dx);
            r->yDlt += cv;
            cover = orc_sign(r, (uint32_t)cv);
            area = cover * 256;
            r->ex0++;
          }
          cover += orc_sign(r, (uint32_t)(r->fy1 - r->yDlt));
          area = cover * (uint32_t)fx_local;
          orc_merge(row, r->ex0, cover, area);
          if (fx_local == 256) { r->yDlt += r->yLift; orc_err_step(&r->yDlt, &r->yErr, r->yRem, r->dx); }
        }
        r->xDlt = r->xLift; orc_err_step(&r->xDlt, &r->xErr, r->xRem, r->dy);
        row += stride;
        i--;
      }

      r->fy0 = 0; r->fy1 = 256;
      if (!j) return;
      i = j - 1; j = 1;
      if (!i) {
        /* last scanline (:993-1017 / :1136-1160) */
        i = 1; j = 0;
        r->fy1 = r->savedFy1;
        if (rtl) {
          r->xDlt = x_local - ((r->ex1 << 8) + r->fx1);
          r->ex0 = (x_local - 1) >> 8; r->fx0 = ((x_local - 1) & 255) + 1;
          if (r->fx0 - r->xDlt >= 0) {
            cover = orc_sign(r, (uint32_t)r->fy1);
            orc_merge(row, r->ex0, cover, cover * (uint32_t)(r->fx0 * 2 - r->xDlt));
            return;
          }
        }
        else {
          r->xDlt = ((r->ex1 << 8) + r->fx1) - x_local;
          r->ex0 = x_local >> 8; r->fx0 = x_local & 255;
          if (r->fx0 + r->xDlt <= 256) {
            cover = orc_sign(r, (uint32_t)r->fy1);
            orc_merge(row, r->ex0, cover, cover * ((uint32_t)r->fx0 * 2 + (uint32_t)r->xDlt));
            return;
          }
        }
        r->yDlt -= 256;
        cover = orc_sign(r, (uint32_t)r->yDlt);
        skip_to_inside = 1;
      }
    }
  }
}

/* Accumulates `n` edges (x0,y0,x1,y1 in 24.8, ORIGINAL direction: y0 > y1 means sign bit set) into zeroed cells. */
ORC_API void orc_rasterize_edges(const int32_t* edges, size_t n, uint32_t* cells, size_t stride) {
  for (size_t k = 0; k < n; k++) {
    int x0 = edges[k * 4], y0 = edges[k * 4 + 1], x1 = edges[k * 4 + 2], y1 = edges[k * 4 + 3];
    OrcRas r;
    r.sign_mask = 0;
    if (y0 > y1) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; r.sign_mask = 0xFFFFFFFFu; }
    if (!orc_prepare(&r, x0, y0, x1, y1)) continue;
    orc_rasterize(&r, cells, stride);
  }
}

/* Cells -> 8-bit masks for rows [0, h): mask[y][x] = calc_mask(256 << 9 + sum of cells[y][0..x]).  This is what
 * FillAnalytic_Base's VMask / CMask walk (fillgeneric_p.h:207-378) computes for every pixel it composites; pixels it
 * skips have mask 0 in this formulation as well. */
ORC_API void orc_cells_to_masks(const uint32_t* cells, size_t stride, int w, int h, uint32_t fill_rule_mask, uint32_t alpha, uint8_t* masks) {
  for (int y = 0; y < h; y++) {
    uint32_t cov = 256u << 9;
    for (int x = 0; x < w; x++) {
      cov += cells[(size_t)y * stride + (size_t)x];
      masks[(size_t)y * (size_t)w + (size_t)x] = (uint8_t)orc_calc_mask(cov, fill_rule_mask, alpha);
    }
  }
}

/* Unclipped polygon -> edges: every vertex is truncated to 24.8 (Math::trunc_to_int) and consecutive points form a line
 * when their fixed y differ (edgebuilder_p.h:1142-1152, 1282-1294); the polygon is implicitly closed (:1058-1062).
 * pts are already in 24.8 UNITS as doubles (i.e. multiplied by 256).  Returns the number of edges written. */
ORC_API size_t orc_polygon_edges(const double* pts, size_t n, int32_t* edges_out) {
  size_t ne = 0;
  for (size_t k = 0; k < n; k++) {
    size_t k1 = (k + 1) % n;
    int x0 = (int)pts[k * 2], y0 = (int)pts[k * 2 + 1], x1 = (int)pts[k1 * 2], y1 = (int)pts[k1 * 2 + 1];
    if (y0 == y1) continue;
    edges_out[ne * 4] = x0; edges_out[ne * 4 + 1] = y0; edges_out[ne * 4 + 2] = x1; edges_out[ne * 4 + 3] = y1;
    ne++;
  }
  return ne;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Axis-unaligned box: FillData::init_box_u_8bpc_24x8 builds a small program of mask commands (pipedefs_p.h:644-815)
 * and FillMask_Base interprets it with repeat counters (fillgeneric_p.h:67-162).  Both are restated; the result is the
 * mask of every pixel of the box's outer pixel rectangle.  Returns 0 when the reference would draw nothing.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { uint32_t x0, x1, type; uint32_t value; const uint8_t* ptr; } OrcMaskCmd;   /* type 0 end/repeat, 1 CMask, 2 VMask */

static void orc_cmask(OrcMaskCmd* c, uint32_t x0, uint32_t x1, uint32_t v) { c->x0 = x0; c->x1 = x1; c->type = 1; c->value = v; c->ptr = 0; }
static void orc_vmask(OrcMaskCmd* c, uint32_t x0, uint32_t x1, const uint8_t* p) { c->x0 = x0; c->x1 = x1; c->type = 2; c->value = 0; c->ptr = p; }
static void orc_end(OrcMaskCmd* c, uint32_t repeat) { c->x0 = repeat; c->x1 = 0; c->type = 0; c->value = 0; c->ptr = 0; }

static void orc_write_mask_row(uint8_t* dst, uint32_t m) { memset(dst, 0, 4); memset(dst + 4, (int)m, 28); }     /* :520-539 */

/* box = x0,y0,x1,y1 in 24.8; out_box = outer pixel box; masks = (out_box h) x (out_box w) bytes, row major. */
ORC_API int orc_box_u_masks(const int32_t* box, uint32_t alpha, int32_t* out_box, uint8_t* masks, size_t masks_capacity) {
  int x0 = box[0], y0 = box[1], x1 = box[2], y1 = box[3];
  uint32_t ax0 = (uint32_t)x0 >> 8, ay0 = (uint32_t)y0 >> 8, ax1 = (uint32_t)(x1 + 0xFF) >> 8, ay1 = (uint32_t)(y1 + 0xFF) >> 8;
  uint32_t fx0 = (uint32_t)x0 & 0xFF, fy0 = (uint32_t)y0 & 0xFF;
  uint32_t fx1 = ((uint32_t)(x1 - 1) & 0xFF) + 1, fy1 = ((uint32_t)(y1 - 1) & 0xFF) + 1;
  uint32_t w = ax1 - ax0, h = ay1 - ay0;
  out_box[0] = (int)ax0; out_box[1] = (int)ay0; out_box[2] = (int)ax1; out_box[3] = (int)ay1;
  if ((size_t)w * h > masks_capacity) return -1;
  memset(masks, 0, (size_t)w * h);

  fy0 = (h == 1 ? fy1 : 256u) - fy0;
  uint32_t fy0_a = fy0 * alpha, fy1_a = fy1 * alpha;

  OrcMaskCmd cmds[12];
  uint8_t mask_data[96];
  OrcMaskCmd* mc = cmds;
  uint8_t* mp = mask_data;
  int by0 = (int)ay0, by1 = (int)(ay0 + h);
  int draw;
  uint32_t span_x0 = ax0;                       /* left end of the spans (moves left when the mask run is padded) */

  if (w == 1) {
    fx0 = fx1 - fx0;
    uint32_t m0 = (fx0 * fy0_a) >> 16;
    orc_cmask(&mc[0], ax0, ax1, m0); orc_end(&mc[1], 1);
    if (h == 1) draw = m0 != 0;
    else {
      mc += m0 ? 2 : 0; by0 += (m0 == 0);
      uint32_t m1 = (fx0 * alpha) >> 8;
      orc_cmask(&mc[0], ax0, ax1, m1); orc_end(&mc[1], h - 2);
      mc += h > 2 ? 2 : 0;
      uint32_t m2 = (fx0 * fy1_a) >> 16;
      orc_cmask(&mc[0], ax0, ax1, m2); orc_end(&mc[1], 1);
      by1 -= (m2 == 0);
      draw = by0 < by1 && m1 != 0;
    }
  }
  else {
    uint32_t m0x1 = fy0_a >> 8, m1x1 = alpha, m2x1 = fy1_a >> 8;
    fx0 = 256 - fx0;
    if ((fx0 & fx1) == 256) {
      orc_cmask(&mc[0], ax0, ax1, m0x1); orc_end(&mc[1], 1); mc += m0x1 ? 2 : 0; by0 += (m0x1 == 0);
      orc_cmask(&mc[0], ax0, ax1, m1x1); orc_end(&mc[1], h - 2); mc += h > 2 ? 2 : 0;
      orc_cmask(&mc[0], ax0, ax1, m2x1); orc_end(&mc[1], 1); by1 -= (m2x1 == 0);
      draw = by0 < by1;
    }
    else {
      uint32_t m0x0 = (fx0 * fy0_a) >> 16, m0x2 = (fx1 * fy0_a) >> 16;
      uint32_t m1x0 = (fx0 * alpha) >> 8, m1x2 = (fx1 * alpha) >> 8;
      uint32_t m2x0 = (fx0 * fy1_a) >> 16, m2x2 = (fx1 * fy1_a) >> 16;
      orc_write_mask_row(mp + 0, m0x1); mp[0 + 4] = (uint8_t)m0x0;
      orc_write_mask_row(mp + 32, m1x1); mp[32 + 4] = (uint8_t)m1x0;
      orc_write_mask_row(mp + 64, m2x1); mp[64 + 4] = (uint8_t)m2x0;
      mp += 4;
      uint32_t w_align = (4u - (w & 3u)) & 3u;                                  /* IntOps::align_up_diff(w, 4) */
      if (w_align > ax0) w_align = 0;
      span_x0 = ax0 - w_align; w += w_align; mp -= w_align;
      if (w <= 20) {
        mp[0 + w - 1] = (uint8_t)m0x2; mp[32 + w - 1] = (uint8_t)m1x2; mp[64 + w - 1] = (uint8_t)m2x2;
        orc_vmask(&mc[0], span_x0, ax1, mp + 0); orc_end(&mc[1], 1); mc += m0x1 ? 2 : 0; by0 += (m0x1 == 0);
        orc_vmask(&mc[0], span_x0, ax1, mp + 32); orc_end(&mc[1], h - 2); mc += h > 2 ? 2 : 0;
        orc_vmask(&mc[0], span_x0, ax1, mp + 64); orc_end(&mc[1], 1); by1 -= (m2x1 == 0);
      }
      else {
        uint32_t inner_width = (w - 5) & ~7u;
        uint32_t inner_end = span_x0 + 4 + inner_width;
        uint32_t tail_width = ax1 - inner_end;
        const uint8_t* tail = mp + 16 - tail_width;
        mp[0 + 15] = (uint8_t)m0x2; mp[32 + 15] = (uint8_t)m1x2; mp[64 + 15] = (uint8_t)m2x2;
        for (int rowk = 0; rowk < 3; rowk++) {
          uint32_t inner = rowk == 0 ? m0x1 : rowk == 1 ? m1x1 : m2x1;
          orc_vmask(&mc[0], span_x0, span_x0 + 4, mp + 32 * rowk);
          orc_cmask(&mc[1], span_x0 + 4, inner_end, inner);
          orc_vmask(&mc[2], inner_end, ax1, tail + 32 * rowk);
          orc_end(&mc[3], rowk == 1 ? h - 2 : 1);
          if (rowk == 0) { mc += m0x1 ? 4 : 0; by0 += (m0x1 == 0); }
          else if (rowk == 1) mc += h > 2 ? 4 : 0;
          else by1 -= (m2x1 == 0);
        }
      }
      draw = by0 < by1;
    }
  }
  if (!draw) return 0;

  /* FillMask_Base::fill_func: walk the program for rows [by0, by1). */
  OrcMaskCmd* cmd = cmds;
  uint32_t rows_left = (uint32_t)(by1 - by0);
  int y = by0;
  for (;;) {
    OrcMaskCmd* begin = cmd;
    while (cmd->type != 0) {
      for (uint32_t x = cmd->x0; x < cmd->x1; x++) {
        uint32_t m = cmd->type == 1 ? cmd->value : cmd->ptr[x - cmd->x0];
        if (x >= ax0) masks[(size_t)(y - (int)ay0) * (ax1 - ax0) + (x - ax0)] = (uint8_t)m;
      }
      cmd++;
    }
    uint32_t repeat = cmd->x0;
    if (--rows_left == 0) break;
    cmd++; y++;
    repeat--;
    cmd[-1].x0 = repeat;
    if (repeat != 0) cmd = begin;
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Linear gradient, incremental form: spanInitY / spanStartX / fetch / advance_y (fetchgeneric_p.h:951-1010).
 * Fills out[h][w] for the pixel rectangle starting at (x0, y0).  lut: u32[lut_size].
 * ---------------------------------------------------------------------------------------------------------------- */
ORC_API void orc_linear_gradient_rect(uint64_t pt0, uint64_t dy, uint64_t dt, uint32_t maxi, uint32_t rori, int is_pad,
                                      const uint32_t* lut, int x0, int y0, int w, int h, uint32_t* out) {
  uint64_t py = pt0 + (uint64_t)(uint32_t)y0 * dy;
  for (int y = 0; y < h; y++) {
    uint64_t pt = py + (uint64_t)(uint32_t)x0 * dt;
    for (int x = 0; x < w; x++) {
      uint32_t idx = (uint32_t)(pt >> 32);
      if (is_pad) { int32_t v = (int32_t)idx; idx = (uint32_t)(v < 0 ? 0 : v > (int32_t)maxi ? (int32_t)maxi : v); }
      else { uint32_t a = idx & maxi, b = (idx & maxi) ^ rori; idx = b < a ? b : a; }
      pt += dt;
      out[(size_t)y * (size_t)w + (size_t)x] = lut[idx];
    }
    py += dy;
  }
}
