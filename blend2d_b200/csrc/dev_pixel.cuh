// dev_pixel.cuh - premultiplied-pixel arithmetic and the composition operators.
//
// Reference semantics restated here (nothing is copied; the reference packs 4x16-bit lanes into one u64, we keep two
// u32 words holding the lanes {.R.B} and {.A.G}, which is the natural width for the GPU's 32-bit integer pipes):
//   - U32_8888::div255/div256/addus8     blend2d/pipeline/reference/pixelgeneric_p.h:390-405
//   - CompOp_SrcCopy_Op / SrcOver / Plus blend2d/pipeline/reference/compopgeneric_p.h:24-81
//   - Multiply / Screen (JIT only)       blend2d/pipeline/jit/compoppart.cpp:4325-4461, 4591-4640
//   - udiv255                            blend2d/pixelops/scalar_p.h:34  (KAT: pixelops/scalar_test.cpp:19-31)
//   - FillAnalytic_Base::calc_mask       blend2d/pipeline/reference/fillgeneric_p.h:381-387
#pragma once
#include "dev_common.cuh"

namespace b2d {

// Two 16-bit lanes in one 32-bit word.
struct Lanes2 { uint32_t rb, ag; };

// (u >> 8) & 0x00FF00FF: the high byte of each 16-bit lane.  One PRMT on the GPU instead of shift + mask.
B2D_HD uint32_t lanes_hi(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(u, 0u, 0x4341);
#else
  return (u >> 8) & 0x00FF00FFu;
#endif
}

B2D_HD Lanes2 unpack(uint32_t p) { return Lanes2{ p & 0x00FF00FFu, lanes_hi(p) }; }
B2D_HD uint32_t pack(Lanes2 u) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(u.rb, u.ag, 0x6240);                 // bytes {rb.0, ag.0, rb.2, ag.2}; high lane bytes are dropped
#else
  return (u.rb & 0x00FF00FFu) | ((u.ag & 0x00FF00FFu) << 8);
#endif
}

// ((x + 128) + ((x + 128) >> 8)) >> 8 on each 16-bit lane; exact for lane values <= 65025 + 254.
B2D_HD uint32_t div255_2(uint32_t x) {
  uint32_t u = x + 0x00800080u;
  return lanes_hi(u + lanes_hi(u));
}

// pack(div255(a)) with the final ">> 8 & mask" of both lane pairs folded into the packing permute.
B2D_HD uint32_t pack_div255(Lanes2 a) {
  uint32_t u0 = a.rb + 0x00800080u, u1 = a.ag + 0x00800080u;
  uint32_t t0 = u0 + lanes_hi(u0), t1 = u1 + lanes_hi(u1);
#if defined(__CUDA_ARCH__)
  return __byte_perm(t0, t1, 0x7351);                     // bytes {t0.1, t1.1, t0.3, t1.3}
#else
  return ((t0 >> 8) & 0x00FF00FFu) | (t1 & 0xFF00FF00u);
#endif
}
B2D_HD uint32_t div256_2(uint32_t x) { return (x >> 8) & 0x00FF00FFu; }

B2D_HD Lanes2 mul(Lanes2 a, uint32_t m) { return Lanes2{ a.rb * m, a.ag * m }; }
B2D_HD Lanes2 add(Lanes2 a, Lanes2 b) { return Lanes2{ a.rb + b.rb, a.ag + b.ag }; }
B2D_HD Lanes2 div255(Lanes2 a) { return Lanes2{ div255_2(a.rb), div255_2(a.ag) }; }
B2D_HD Lanes2 div256(Lanes2 a) { return Lanes2{ div256_2(a.rb), div256_2(a.ag) }; }

// Saturating add of two {0..255} lane pairs (U32_8888::addus8).
B2D_HD uint32_t addus8_2(uint32_t a, uint32_t b) {
  uint32_t v = a + b;
  uint32_t msk = ((v >> 8) & 0x00010001u) * 0xFFu;
  return (v | msk) & 0x00FF00FFu;
}

B2D_HD uint32_t udiv255(uint32_t x) { return ((x + 0x80u) * 0x101u) >> 16; }

// 16-bit-lane helpers for the JIT-specified operators (pmullw / paddw wrap at 16 bits, pmulhuw for div255).
B2D_HD uint32_t jit_div255_u16(uint32_t x) { return (((x + 0x80u) & 0xFFFFu) * 0x101u) >> 16; }

// ---------------------------------------------------------------------------------------------------------------
// Composition operators: d' = op(d, s, m) with m in [1, 255].  m == 0 leaves d unchanged for every operator below
// (the callers skip such pixels), m == 255 reproduces the reference's "opaque" variants exactly.
// ---------------------------------------------------------------------------------------------------------------

// SrcCopy: (d*(255-m) + s*m).div255()                                      compopgeneric_p.h:38-40
B2D_HD uint32_t comp_src_copy(uint32_t d, uint32_t s, uint32_t m) {
  Lanes2 du = unpack(d), su = unpack(s);
  uint32_t im = m ^ 0xFFu;
  return pack_div255(add(mul(du, im), mul(su, m)));
}

// SrcOver: s' = (s*m).div255(); d' = s' + (d*(255 - s'.a)).div255()        compopgeneric_p.h:54-62
// The final "+" is a plain 32-bit add of packed pixels in the reference; kept as such.
B2D_HD uint32_t comp_src_over(uint32_t d, uint32_t s, uint32_t m) {
  uint32_t sm = pack_div255(mul(unpack(s), m));
  uint32_t ia = (sm >> 24) ^ 0xFFu;
  return sm + pack_div255(mul(unpack(d), ia));
}

// The same operator at m == 255: div255(s * 255) == s for every 8-bit s, so the first half is the identity.
B2D_HD uint32_t comp_src_over_opaque_mask(uint32_t d, uint32_t s) {
  uint32_t ia = (s >> 24) ^ 0xFFu;
  return s + pack_div255(mul(unpack(d), ia));
}

// Plus: addus8(d, (s*m).div255())                                          compopgeneric_p.h:74-80
B2D_HD uint32_t comp_plus(uint32_t d, uint32_t s, uint32_t m) {
  Lanes2 du = unpack(d);
  Lanes2 sm = div255(mul(unpack(s), m));
  return pack(Lanes2{ addus8_2(du.rb, sm.rb), addus8_2(du.ag, sm.ag) });
}

B2D_HD uint32_t sat8(uint32_t v) { return v > 255u ? 255u : v; }

// Multiply: S = div255(S*m); D' = div255(D*(S + 255 - Sa) + S*(255 - Da))   jit/compoppart.cpp:4408-4441
B2D_HD uint32_t comp_multiply(uint32_t d, uint32_t s, uint32_t m) {
  uint32_t out = 0;
  uint32_t sa = jit_div255_u16(((s >> 24) * m) & 0xFFFFu);
  uint32_t da = d >> 24;
  uint32_t isa = 255u - sa;
  uint32_t ida = 255u - da;
  #pragma unroll
  for (int sh = 0; sh < 32; sh += 8) {
    uint32_t sc = jit_div255_u16((((s >> sh) & 0xFFu) * m) & 0xFFFFu);
    uint32_t dc = (d >> sh) & 0xFFu;
    uint32_t y = (isa + sc) & 0xFFFFu;
    uint32_t v = ((dc * y) & 0xFFFFu) + ((ida * sc) & 0xFFFFu);
    out |= sat8(jit_div255_u16(v & 0xFFFFu)) << sh;
  }
  return out;
}

// Screen: S = div255(S*m); D' = div255(D*(255 - S)) + S                    jit/compoppart.cpp:4614-4630
B2D_HD uint32_t comp_screen(uint32_t d, uint32_t s, uint32_t m) {
  uint32_t out = 0;
  #pragma unroll
  for (int sh = 0; sh < 32; sh += 8) {
    uint32_t sc = jit_div255_u16((((s >> sh) & 0xFFu) * m) & 0xFFFFu);
    uint32_t dc = (d >> sh) & 0xFFu;
    uint32_t v = jit_div255_u16((dc * (255u - sc)) & 0xFFFFu) + sc;
    out |= sat8(v & 0xFFFFu) << sh;
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// The other operators the JIT specifies for a PRGB32 destination (jit/compoppart.cpp v_mask_proc_rgba32_vec, the
// `has_mask` variants :3731-5350; at m == 255 they reduce to the unmasked ones).  Restated channel by channel with the
// JIT's 16-bit lane semantics: products and sums wrap at 16 bits (pmullw / paddw), v_div255_u16 is
// ((x + 128) * 257) >> 16, the result is packed with unsigned saturation of the SIGNED 16-bit lane (packuswb).
// The reference build here has no JIT, so these are unpinned like Multiply and Screen.
// ---------------------------------------------------------------------------------------------------------------
enum CompOpId : uint32_t {
  kOpSrcOver = 0, kOpSrcCopy = 1, kOpSrcIn = 2, kOpSrcOut = 3, kOpSrcAtop = 4, kOpDstOver = 5, kOpDstCopy = 6, kOpDstIn = 7,
  kOpDstOut = 8, kOpDstAtop = 9, kOpXor = 10, kOpClear = 11, kOpPlus = 12, kOpMinus = 13, kOpModulate = 14, kOpMultiply = 15,
  kOpScreen = 16, kOpOverlay = 17, kOpDarken = 18, kOpLighten = 19, kOpColorDodge = 20, kOpColorBurn = 21, kOpLinearBurn = 22,
  kOpLinearLight = 23, kOpPinLight = 24, kOpHardLight = 25, kOpSoftLight = 26, kOpDifference = 27, kOpExclusion = 28
};

B2D_HD uint32_t w16(uint32_t v) { return v & 0xFFFFu; }
B2D_HD uint32_t subs_u16(uint32_t a, uint32_t b) { return a > b ? a - b : 0u; }
B2D_HD uint32_t packus_i16(uint32_t v) { int32_t x = int32_t(int16_t(uint16_t(v))); return uint32_t(x < 0 ? 0 : x > 255 ? 255 : x); }
B2D_HD uint32_t minmax_u8x2(uint32_t a, uint32_t b, bool take_min) {            // pminub / pmaxub on the two bytes of a lane
  uint32_t al = a & 0xFFu, ah = (a >> 8) & 0xFFu, bl_ = b & 0xFFu, bh = (b >> 8) & 0xFFu;
  uint32_t lo = take_min ? (al < bl_ ? al : bl_) : (al > bl_ ? al : bl_);
  uint32_t hi = take_min ? (ah < bh ? ah : bh) : (ah > bh ? ah : bh);
  return lo | (hi << 8);
}

B2D_HD uint32_t comp_jit_ext(uint32_t op, uint32_t d, uint32_t s, uint32_t m) {
  const uint32_t da = d >> 24, sa = s >> 24;
  const uint32_t n = 255u - m;                                                  // CompOpPart_negateMask
  const uint32_t sma = jit_div255_u16(w16(sa * m));                             // alpha of S.m
  uint32_t out = 0;
  #pragma unroll
  for (int sh = 0; sh < 32; sh += 8) {
    const bool is_alpha = sh == 24;
    const uint32_t dc = (d >> sh) & 0xFFu, sc = (s >> sh) & 0xFFu;
    const uint32_t sm = jit_div255_u16(w16(sc * m));                            // S.m
    uint32_t v;
    switch (op) {
      case kOpSrcIn: {        // Dca' = Sca.m.Da + Dca.(1 - m)                                          :3751-3774
        uint32_t x = w16(jit_div255_u16(w16(da * sc)) * m);
        v = jit_div255_u16(w16(w16(dc * n) + x));
        break;
      }
      case kOpSrcOut: {       // Dca' = Sca.(1 - Da).m + Dca.(1 - m)                                    :3801-3826
        uint32_t x = w16(jit_div255_u16(w16((255u - da) * sc)) * m);
        v = jit_div255_u16(w16(w16(dc * n) + x));
        break;
      }
      case kOpSrcAtop:        // Dca' = Sca.Da.m + Dca.(1 - Sa.m)                                       :3859-3879
        v = jit_div255_u16(w16(w16(dc * (255u - sma)) + w16(da * sm)));
        break;
      case kOpDstOver:        // Dca' = Dca + Sca.m.(1 - Da); the sum is a 32-bit add of packed pixels   :3919-3937
        v = packus_i16(jit_div255_u16(w16((255u - da) * sm)));
        break;
      case kOpDstIn:          // Dca' = Dca.(1 - m.(1 - Sa))                                            :3969-3983
        v = jit_div255_u16(w16(dc * (255u - jit_div255_u16(w16((255u - sa) * m)))));
        break;
      case kOpDstOut:         // Dca' = Dca.(1 - Sa.m)                                                  :4012-4026
        v = jit_div255_u16(w16(dc * (255u - sma)));
        break;
      case kOpDstAtop: {      // Dca' = Dca.(1 - m.(1 - Sa)) + Sca.m.(1 - Da)                           :4061-4085
        uint32_t u = 255u - jit_div255_u16(w16((255u - sa) * m));
        v = jit_div255_u16(w16(w16(dc * u) + w16((255u - da) * sm)));
        break;
      }
      case kOpXor:            // Dca' = Dca.(1 - Sa.m) + Sca.m.(1 - Da)                                 :4119-4140
        v = jit_div255_u16(w16(w16(dc * (255u - sma)) + w16(sm * (255u - da))));
        break;
      case kOpMinus: {        // Dca' = (Clamp(Dca - Sca) + Sca.(1 - Da)).m + Dca.(1 - m); Da' = Da + Sa.m.(1 - Da)  :4229-4256
        uint32_t t = is_alpha ? 0u : subs_u16(dc, sc);
        t = w16(t + jit_div255_u16(w16(sc * (255u - da))));
        v = jit_div255_u16(w16(w16(t * m) + w16(dc * (is_alpha ? 255u : n))));
        break;
      }
      case kOpModulate:       // Dca' = Dca.(Sca.m + 1 - m)                                             :4303-4313
        v = jit_div255_u16(w16(dc * w16(sm + 255u - m)));
        break;
      case kOpDarken:
      case kOpLighten: {      // Dca' = minmax(Dca + Sca.(1 - Da), Sca + Dca.(1 - Sa)) on S.m           :4654-4678
        uint32_t x = jit_div255_u16(w16((255u - da) * sm));
        uint32_t y = jit_div255_u16(w16((255u - sma) * dc));
        v = minmax_u8x2(w16(dc + x), w16(sm + y), op == kOpDarken);
        break;
      }
      case kOpLinearBurn:     // Dca' = Dca + Sca - Sa.Da on S.m                                        :4861-4870
        v = subs_u16(w16(dc + sm), jit_div255_u16(w16(sma * da)));
        break;
      case kOpDifference: {   // Dca' = Dca + Sca.m - 2.min(Sca.Da, Dca.Sa).m; alpha subtracts it once    :5287-5314
        uint32_t a = w16(sma * dc), b = w16(da * sm);
        uint32_t y = jit_div255_u16(a < b ? a : b);
        v = w16(w16(dc + sm) - y);
        if (!is_alpha) v = w16(v - y);
        break;
      }
      default: {              // kOpExclusion: Dca' = Dca + Sca - 2.Sca.Dca on S.m; alpha subtracts once  :5326-5346
        uint32_t x = jit_div255_u16(w16(dc * sm));
        v = w16(w16(dc + sm) - x);
        if (!is_alpha) v = w16(v - x);
        break;
      }
    }
    out |= packus_i16(v) << sh;
  }
  if (op == kOpDstOver) out += d;                                               // v_add_i32 on the packed pixels
  return out;
}

// Overlay, HardLight, PinLight, LinearLight (16-bit integer lanes) and ColorDodge, ColorBurn, SoftLight (the JIT's
// one-pixel float variants), masked forms with Da and Sa used (jit/compoppart.cpp:4466-4540, 5091-5147, 4954-5001,
// 4889-4935, 4724-4778, 4784-4841, 5152-5247).  Per channel; `sa` / `da` are the alphas of S.m and D, a value written
// `-x` is the 16-bit two's complement the JIT's psubw produces, comparisons are the signed pcmpgtw / pminsw.
// Float steps are single IEEE operations in the JIT's order (mulps, divps, sqrtps, no contraction; v_madd_f32 is
// mul + add on the baseline target), conversions are cvttps2dq / cvtps2dq.  Unpinned (no JIT in the reference build here).
B2D_HD int32_t s16(uint32_t v) { return int32_t(int16_t(uint16_t(v))); }
B2D_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
B2D_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
B2D_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
B2D_HD float f_sqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
B2D_HD float f_max_ps(float a, float b) { return a > b ? a : b; }               // maxps / minps: the second operand unless the test holds
B2D_HD float f_min_ps(float a, float b) { return a < b ? a : b; }
B2D_HD int f_trunc_i32(float v) {                                                // cvttps2dq
#if defined(__CUDA_ARCH__)
  return (v >= 2147483648.0f || v != v) ? int(0x80000000u) : __float2int_rz(v);
#else
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return int(0x80000000u);
  return int(v);
#endif
}
B2D_HD int f_round_i32(float v) {                                                // cvtps2dq (round to nearest even)
#if defined(__CUDA_ARCH__)
  return (v >= 2147483648.0f || v != v) ? int(0x80000000u) : __float2int_rn(v);
#else
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return int(0x80000000u);
  return int(lrintf(v));
#endif
}
B2D_HD uint32_t packus_dw(int v) { return uint32_t(v < 0 ? 0 : v > 65535 ? 65535 : v); }   // packusdw

B2D_HD uint32_t comp_jit_light(uint32_t op, uint32_t d, uint32_t s, uint32_t m) {
  const uint32_t da = d >> 24;
  const uint32_t sa = jit_div255_u16(w16((s >> 24) * m));
  const uint32_t sada = jit_div255_u16(w16(sa * da));
  uint32_t out = 0;

  if (op == kOpSoftLight) {
    // Dca' = Dca + Sca.(1 - Da) + (2.Sca - Sa).Da.F(Dc), Dc = Dca / max(Da, 0.001), everything in 0..1 floats;
    // F = Dc.(1 - Dc) for 2.Sca <= Sa, else 4.Dc.(4.Dc.Dc + Dc - 4.Dc + 1) - Dc for 4.Dc <= 1, else sqrt(Dc) - Dc.
    const float k = 1.0f / 255.0f;
    const float fsa = f_mul(float(int(sa)), k), fda = f_mul(float(int(da)), k);
    const float b0 = f_max_ps(fda, 1e-3f);
    #pragma unroll
    for (int sh = 0; sh < 32; sh += 8) {
      const float sc = f_mul(float(int(jit_div255_u16(w16(((s >> sh) & 0xFFu) * m)))), k);
      const float dc = f_mul(float(int((d >> sh) & 0xFFu)), k);
      const float a0 = f_div(dc, b0);
      float r = f_add(f_add(dc, sc), -f_mul(sc, fda));                          // Dca + Sca - Sca.Da
      if (sh != 24) {
        const float t = f_mul(f_add(f_add(sc, sc), -fsa), b0);                   // (2.Sca - Sa).Da
        const float q = f_mul(a0, 4.0f);
        const float poly = f_mul(f_add(f_add(f_add(f_mul(q, a0), a0), -q), 1.0f), q);
        const float hi = q <= 1.0f ? poly : f_sqrt(a0);
        const float f = 0.0f < t ? f_add(hi, -a0) : f_mul(f_add(1.0f, -a0), a0);
        r = f_add(r, f_mul(t, f));
      }
      else r = f_add(r, f_mul(0.0f, 0.0f));                                     // the alpha lane of (2.Sca - Sa).Da is masked to +0
      int v = f_round_i32(f_mul(r, 255.0f));
      v = v < -32768 ? -32768 : v > 32767 ? 32767 : v;                           // packssdw, then packuswb
      out |= uint32_t(v < 0 ? 0 : v > 255 ? 255 : v) << sh;
    }
    return out;
  }

  if (op == kOpColorDodge || op == kOpColorBurn) {
    // Dodge: Dca' = min(Dca.Sa.Sa / max(Sa - Sca, 0.001), Sa.Da) + Sca.(1 - Da) + Dca.(1 - Sa)
    // Burn:  Dca' = Sa.Da - min(Sa.Da, (Da - Dca).Sa.Sa / max(Sca, 0.001)) + Sca.(1 - Da) + Dca.(1 - Sa)
    // in 0..255 units: the float term is truncated and added to the 16-bit sum before the division by 255.
    const float fsa = float(int(sa)), fda = float(int(da));
    const float dasa = f_mul(fda, fsa);
    float lim;                                                                   // the alpha lane of the float term
    if (op == kOpColorDodge) lim = f_mul(f_div(dasa, f_max_ps(fsa, 1e-3f)), fsa);
    else { const float ya = f_max_ps(fsa, 1e-3f); lim = f_mul(f_div(dasa, ya), ya); }
    #pragma unroll
    for (int sh = 0; sh < 32; sh += 8) {
      const uint32_t dc = (d >> sh) & 0xFFu, sc = jit_div255_u16(w16(((s >> sh) & 0xFFu) * m));
      const float fd = float(int(dc)), fs = float(int(sc));
      float term;
      if (op == kOpColorDodge) {
        const float den = f_max_ps(sh == 24 ? fsa : f_add(-fs, fsa), 1e-3f);
        term = f_min_ps(f_mul(f_div(f_mul(fd, fsa), den), fsa), lim);
      }
      else {
        const float num = sh == 24 ? dasa : f_add(-f_mul(fd, fsa), dasa);
        const float z = f_min_ps(f_mul(f_div(num, f_max_ps(fs, 1e-3f)), f_max_ps(fsa, 1e-3f)), lim);
        term = f_add(lim, -(sh == 24 ? 0.0f : z));
      }
      const uint32_t ip = w16(w16(dc * (255u - sa)) + w16(sc * (255u - da)));
      out |= packus_i16(jit_div255_u16(w16(ip + packus_dw(f_trunc_i32(term))))) << sh;
    }
    return out;
  }

  #pragma unroll
  for (int sh = 0; sh < 32; sh += 8) {
    const bool is_alpha = sh == 24;
    const uint32_t dc = (d >> sh) & 0xFFu, sc = jit_div255_u16(w16(((s >> sh) & 0xFFu) * m));
    uint32_t v;
    switch (op) {
      case kOpOverlay:
      case kOpHardLight: {
        // X = Dca.Sa + Sca.Da - 2.Sca.Dca; Overlay tests 2.Dca < Da, HardLight 2.Sca < Sa:
        //   true:  Dca' = Dca + Sca - X          false: Dca' = Dca + Sca + X - Sa.Da          Da' = Da + Sa - Sa.Da
        uint32_t x = w16(w16(w16(sc * da) - w16(dc * sc)) + w16(sa * dc));
        x = jit_div255_u16(is_alpha ? (op == kOpOverlay ? w16(sa * da) : 0u) : w16(x - w16(dc * sc)));
        const bool lt = is_alpha ? op == kOpOverlay : (op == kOpOverlay ? s16(da) > s16(w16(dc << 1)) : s16(sa) > s16(w16(sc << 1)));
        const uint32_t y = (lt && !(is_alpha && op == kOpHardLight)) ? 0u : sada;
        v = w16(w16(w16(dc + sc) + (lt ? w16(0u - x) : x)) - (is_alpha && op == kOpOverlay ? 0u : y));
        break;
      }
      case kOpPinLight: {
        // 2.Sca <= Sa: min(Dca + Sca - Sca.Da, Dca + Sca + Sca.Da - Dca.Sa); else max(.., .. - Da.Sa)
        const uint32_t y = jit_div255_u16(w16(sa * dc)), x = jit_div255_u16(w16(da * sc)), sum = w16(dc + sc);
        const uint32_t a = w16(sum - x);
        uint32_t b = w16(x - w16(y - sum));
        const bool gt = s16(w16(sc << 1)) > s16(sa);
        if (gt) b = w16(b - sada);
        v = gt ? (s16(a) > s16(b) ? a : b) : (s16(a) < s16(b) ? a : b);
        break;
      }
      default: {              // kOpLinearLight: Dca' = min(max(Dca.Sa + 2.Sca.Da - Sa.Da, 0), Sa.Da) + Sca.(1 - Da) + Dca.(1 - Sa)
        uint32_t t = w16(jit_div255_u16(w16(dc * sa)) + 2u * jit_div255_u16(w16(sc * da)));
        t = subs_u16(t, sada);
        t = s16(t) < s16(sada) ? t : sada;
        v = w16(t + jit_div255_u16(w16(w16(dc * (255u - sa)) + w16(sc * (255u - da)))));
        break;
      }
    }
    out |= packus_i16(v) << sh;
  }
  return out;
}

B2D_HD uint32_t composite(uint32_t comp_op, uint32_t d, uint32_t s, uint32_t m) {
  switch (comp_op) {
    case kOpSrcOver:  return comp_src_over(d, s, m);
    case kOpSrcCopy:  return comp_src_copy(d, s, m);
    case kOpPlus:     return comp_plus(d, s, m);
    case kOpMultiply: return comp_multiply(d, s, m);
    case kOpScreen:   return comp_screen(d, s, m);
    case kOpOverlay: case kOpColorDodge: case kOpColorBurn: case kOpLinearLight: case kOpPinLight: case kOpHardLight: case kOpSoftLight:
                      return comp_jit_light(comp_op, d, s, m);
    default:          return comp_jit_ext(comp_op, d, s, m);
  }
}

// The operators outside the hot loop's instruction footprint (see B2D_HD_COLD).
B2D_HD_COLD uint32_t composite_cold(uint32_t comp_op, uint32_t d, uint32_t s, uint32_t m) {
  switch (comp_op) {
    case kOpPlus:     return comp_plus(d, s, m);
    case kOpMultiply: return comp_multiply(d, s, m);
    case kOpScreen:   return comp_screen(d, s, m);
    case kOpOverlay: case kOpColorDodge: case kOpColorBurn: case kOpLinearLight: case kOpPinLight: case kOpHardLight: case kOpSoftLight:
                      return comp_jit_light(comp_op, d, s, m);
    default:          return comp_jit_ext(comp_op, d, s, m);
  }
}

// d[i] = op(d[i], s[i], m[i]) for the pixels whose mask is non-zero; operator dispatch hoisted out of the loop.
// `all_opaque`: every non-zero m[i] the caller passes is 255 (decided per warp, so the branch never diverges).
B2D_HD void composite4(uint32_t comp_op, uint32_t* d, const uint32_t* s, const uint32_t* m, bool all_opaque = false) {
  // SrcOver and SrcCopy are the identity at m == 0, so the general variants run unguarded (no branch per pixel); the
  // m == 255 variants select.
  if (comp_op == kOpSrcOver) {
    if (all_opaque) {
      #pragma unroll
      for (int i = 0; i < 4; i++) { uint32_t v = comp_src_over_opaque_mask(d[i], s[i]); d[i] = m[i] ? v : d[i]; }
    }
    else {
      #pragma unroll
      for (int i = 0; i < 4; i++) d[i] = comp_src_over(d[i], s[i], m[i]);
    }
  }
  else if (comp_op == kOpSrcCopy) {
    if (all_opaque) {
      #pragma unroll
      for (int i = 0; i < 4; i++) d[i] = m[i] ? s[i] : d[i];              // div255(s * 255) == s
    }
    else {
      #pragma unroll
      for (int i = 0; i < 4; i++) d[i] = comp_src_copy(d[i], s[i], m[i]);
    }
  }
  else {
    #pragma unroll
    for (int i = 0; i < 4; i++) if (m[i]) d[i] = composite_cold(comp_op, d[i], s[i], m[i]);
  }
}

// Coverage accumulator -> 8-bit mask.  `cov` is the running u32 sum that starts at 256 << 9 on every scanline.
// fillgeneric_p.h:381-387: m = min(abs((sar(cov, 9) & rule) - 256), 256) * alpha >> 8.
B2D_HD uint32_t calc_mask(uint32_t cov, uint32_t fill_rule_mask, uint32_t alpha) {
  int32_t c = int32_t(cov) >> 9;                          // IntOps::sar on the reinterpreted value
  uint32_t m = (uint32_t(c) & fill_rule_mask) - 256u;
  int32_t mi = int32_t(m);
  uint32_t a = uint32_t(mi < 0 ? -mi : mi);
  a = a < 256u ? a : 256u;
  return (a * alpha) >> 8;
}

// Source pixel format adaptation (PixelIO<P32_A8R8G8B8, fmt>::fetch, pixelgeneric_p.h:632-676).
B2D_HD uint32_t adapt_src_xrgb32(uint32_t p) { return p | 0xFF000000u; }
B2D_HD uint32_t adapt_src_a8(uint32_t a) { return a * 0x01010101u; }

} // namespace b2d
