// stream.cu - k_stream_one: the streaming compositor for a batch that is ONE large FillBoxA with a non-solid source
// (fill_all / fill_rect with a gradient or a pattern, blit_image of a large image) on a 32-bit target.
//
// This is FillBoxA_Base (blend2d/pipeline/reference/fillgeneric_p.h:22-65) + one fetcher of fetchgeneric_p.h + one
// operator of compopgeneric_p.h over a large box: pure HBM streaming, 8 B per pixel (+ 4 B per pixel of pattern source).
// The generic k_box_stream re-decodes its command list for every 4-pixel chunk; here everything that depends only on
// the command is folded into registers before the loop, everything that depends only on the row (the FP64 row origins
// of the radial / conic gradients, the source row of a pattern) is computed by one lane per row and shuffled, and
// everything that depends only on the column is shared by the four rows a lane holds.
//
// Work item of a WARP: 128 columns x 4 rows (one 16-byte vector per lane and row, four independent loads in flight).
// Consecutive warps take consecutive column blocks of the same four rows, so a CTA streams 4 KB runs of each row.
// The grid is persistent (one wave); items are handed out round robin.
#include "kernels.h"
#include "dev_pixel.cuh"
#include "dev_fetch.cuh"

#include <cuda_runtime.h>

namespace b2d {

enum : int { kOneLinear = 0, kOneRadial = 1, kOneConic = 2, kOnePattern32 = 3, kOneGeneric = 4 };

// Source columns of a lane's four pixels for the aligned pattern fetchers (FetchPatternAligned*, horizontal extend
// contexts of fetchgeneric_p.h:186-352): the first one through the closed form, the next three by stepping.
struct PatCols { uint32_t c[4]; bool vec; };

__device__ __forceinline__ PatCols pattern_cols4(const b2dgpu_fetch_pattern& p, uint32_t ft, uint32_t x) {
  PatCols o;
  const uint32_t w = uint32_t(p.src.w);
  if (ft == B2DGPU_FETCH_PATTERN_ALIGNED_BLIT) {
    const uint32_t c0 = x - uint32_t(p.simple.tx);
    o.c[0] = c0; o.c[1] = c0 + 1u; o.c[2] = c0 + 2u; o.c[3] = c0 + 3u;
  }
  else if (ft == B2DGPU_FETCH_PATTERN_ALIGNED_PAD) {
    #pragma unroll
    for (int i = 0; i < 4; i++) o.c[i] = pattern_col_pad(p, x + uint32_t(i));
  }
  else if (ft == B2DGPU_FETCH_PATTERN_ALIGNED_REPEAT) {
    uint32_t c = pattern_col_repeat(p, x);
    #pragma unroll
    for (int i = 0; i < 4; i++) { o.c[i] = c; c = c + 1u == w ? 0u : c + 1u; }
  }
  else {
    // reflect: v walks [0, rx) and folds at w (pattern_col_ror)
    const uint64_t rx = uint64_t(int64_t(p.simple.rx));
    uint64_t v = (uint64_t(x) + uint64_t(int64_t(p.simple.tx))) % rx;
    #pragma unroll
    for (int i = 0; i < 4; i++) {
      int64_t f = int64_t(v);
      if (f >= int64_t(w)) f -= int64_t(rx);
      o.c[i] = uint32_t(f ^ (f >> 63));
      v = v + 1u == rx ? 0u : v + 1u;
    }
  }
  o.vec = o.c[1] == o.c[0] + 1u && o.c[2] == o.c[0] + 2u && o.c[3] == o.c[0] + 3u;
  return o;
}

// Gradient table staged in shared memory by ONE bulk asynchronous copy (cp.async.bulk, the TMA engine's 1-D form; SASS
// UBLKCP + SYNCS.ARRIVE.TRANS64): thread 0 arms an mbarrier with the byte count and issues the copy, every thread waits
// on the barrier's phase before its first lookup.  Tables of up to kLutSmemEntries entries (the reference builds 256 to
// 1024 entries per gradient, core/gradient.cpp) that are 16-byte aligned are staged; anything else stays on LDG.
enum : uint32_t { kLutSmemEntries = 2048 };

__device__ __forceinline__ void lut_stage_begin(uint32_t* s_lut, uint64_t* s_bar, const uint32_t* lut, uint32_t bytes) {
  const uint32_t bar = uint32_t(__cvta_generic_to_shared(s_bar));
  const uint32_t dst = uint32_t(__cvta_generic_to_shared(s_lut));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(lut), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void lut_stage_wait(uint64_t* s_bar) {
  const uint32_t bar = uint32_t(__cvta_generic_to_shared(s_bar));
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar) : "memory");
  }
}

template<int FC, bool STAGED>
__global__ void __launch_bounds__(256) k_stream_one(StreamOneParams P) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  const uint32_t warp0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);

  // The FetchData is read-only for the whole launch: the few fields a specialisation needs are copied into registers
  // (by-value copies of the sub-structures; the compiler keeps what is used), so nothing is re-read after a store.
  const b2dgpu_fetch_data& fdg = *P.fd;
  b2dgpu_gradient_linear lin;
  b2dgpu_gradient_radial rad;
  b2dgpu_gradient_conic con;
  b2dgpu_fetch_pattern pat;
  const uint32_t* __restrict__ lut = nullptr;
  if (FC == kOneLinear) { lin = fdg.gradient.linear; lut = static_cast<const uint32_t*>(fdg.gradient.lut.data); }
  if (FC == kOneRadial) { rad = fdg.gradient.radial; lut = static_cast<const uint32_t*>(fdg.gradient.lut.data); }
  if (FC == kOneConic) { con = fdg.gradient.conic; lut = static_cast<const uint32_t*>(fdg.gradient.lut.data); }
  if (FC == kOnePattern32) pat = fdg.pattern;
  // Gradient table in shared memory (STAGED: the launcher saw a table of at most kLutSmemEntries entries).
  __shared__ __align__(128) uint32_t s_lut[STAGED ? kLutSmemEntries : 4];
  __shared__ __align__(8) uint64_t s_lut_bar;
  if (STAGED) {
    const uint32_t entries = min(fdg.gradient.lut.size, uint32_t(kLutSmemEntries));
    if ((entries & 3u) == 0u && (uintptr_t(lut) & 15u) == 0u) {          // block uniform
      if (threadIdx.x == 0) lut_stage_begin(s_lut, &s_lut_bar, lut, entries * 4u);
      __syncthreads();                                   // the barrier is initialised before anybody polls it
      lut_stage_wait(&s_lut_bar);
    }
    else {
      // a table the bulk copy cannot take (16-byte granularity): plain cooperative copy
      for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) s_lut[i] = __ldg(lut + i);
      __syncthreads();
    }
  }
  FetchEnv env;
  env.fd = P.fd;
  env.bayer = P.bayer;
  env.fetch_type = P.fetch_type;
  env.src_format = P.src_format;
  env.solid = 0;
  env.origin_x = P.origin_x; env.origin_y = P.origin_y;

  const uint32_t ft = P.fetch_type;
  const uint32_t alpha = P.alpha;
  const bool opaque = alpha == 255u;
  const uint32_t comp_op = P.comp_op;
  const int cb0 = P.x0 / 128;
  const uint32_t ncb = uint32_t((P.x1 + 127) / 128 - cb0);
  const uint32_t nrq = uint32_t((P.y1 - P.y0 + 3) / 4);
  const uint32_t items = ncb * nrq;

  for (uint32_t it = warp0; it < items; it += warps) {
    const uint32_t rq = it / ncb;
    const uint32_t cb = it - rq * ncb;
    const int x = (cb0 + int(cb)) * 128 + lane * 4;
    const int y = P.y0 + int(rq) * 4;
    const int rows = min(4, P.y1 - y);

    // ---- destination: four independent 16-byte loads ----
    uint8_t* ptr = P.dst + size_t(y - P.y_begin) * P.dst_stride + size_t(x) * 4;
    uint4 v[4];
    #pragma unroll
    for (int j = 0; j < 4; j++)
      if (j < rows) v[j] = *reinterpret_cast<const uint4*>(ptr + size_t(j) * P.dst_stride);

    // ---- masks (constant inside the box; the first / last column block may be cut) ----
    uint32_t m[4];
    #pragma unroll
    for (int i = 0; i < 4; i++) m[i] = (x + i >= P.x0 && x + i < P.x1) ? alpha : 0u;
    const bool any = (m[0] | m[1] | m[2] | m[3]) != 0u;

    // ---- per-row state, computed by lane j for row j ----
    RowCtx3 rc; rc.a = rc.b = rc.c = 0;
    int prow = 0;
    if (FC == kOneRadial || FC == kOneConic) {
      if (lane < 4) rc = fetch_row_ctx(ft, fdg, uint32_t(y + lane));
    }
    if (FC == kOnePattern32) {
      if (lane < 4) prow = ft == B2DGPU_FETCH_PATTERN_ALIGNED_BLIT ? int(uint32_t(y + lane) - uint32_t(pat.simple.ty)) : pattern_row(pat, uint32_t(y + lane));
    }

    // ---- per-column state, shared by the four rows ----
    uint64_t lin_x = 0;
    PatCols pc; pc.vec = false; pc.c[0] = pc.c[1] = pc.c[2] = pc.c[3] = 0;
    if (FC == kOneLinear) lin_x = lin.pt[0].u64 + uint64_t(uint32_t(x)) * lin.dt.u64;
    if (FC == kOnePattern32) pc = pattern_cols4(pat, ft, uint32_t(x));

    #pragma unroll
    for (int j = 0; j < 4; j++) {
      // shuffles are executed by the whole warp (rows is warp uniform)
      RowCtx3 rj;
      int prj = 0;
      if (FC == kOneRadial || FC == kOneConic) {
        rj.a = __shfl_sync(0xFFFFFFFFu, rc.a, j); rj.b = __shfl_sync(0xFFFFFFFFu, rc.b, j); rj.c = __shfl_sync(0xFFFFFFFFu, rc.c, j);
      }
      if (FC == kOnePattern32) prj = __shfl_sync(0xFFFFFFFFu, prow, j);
      if (j >= rows) break;
      if (!any) continue;

      uint32_t s[4] = { 0, 0, 0, 0 };
      const uint32_t yy = uint32_t(y + j);
      if (FC == kOneLinear) {
        const b2dgpu_gradient_linear& l = lin;
        const bool pad = ft == B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD;
        uint64_t pt = lin_x + uint64_t(yy) * l.dy.u64;
        #pragma unroll
        for (int i = 0; i < 4; i++) {
          uint32_t idx = uint32_t(pt >> 32);
          idx = pad ? grad_index_pad(idx, l.maxi) : grad_index_ror(idx, l.maxi, l.rori);
          s[i] = STAGED ? s_lut[idx] : __ldg(lut + idx);
          pt += l.dt.u64;
        }
      }
      else if (FC == kOneRadial) {
        const b2dgpu_gradient_radial& r = rad;
        const bool pad = ft == B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD;
        RadialRow row; row.b = f32_from_bits(rj.a); row.d = f32_from_bits(rj.b); row.dd = f32_from_bits(rj.c);
        #pragma unroll
        for (int i = 0; i < 4; i++) {
          uint32_t idx = radial_index(r, row, uint32_t(x + i));
          idx = pad ? grad_index_pad(idx, r.maxi) : grad_index_ror(idx, r.maxi, r.rori);
          s[i] = STAGED ? s_lut[idx] : __ldg(lut + idx);
        }
      }
      else if (FC == kOneConic) {
        ConicRow row; row.tx = f32_from_bits(rj.a); row.ay = f32_from_bits(rj.b); row.by = f32_from_bits(rj.c);
        #pragma unroll
        for (int i = 0; i < 4; i++) { const uint32_t idx = conic_index(con, row, uint32_t(x + i)); s[i] = STAGED ? s_lut[idx] : __ldg(lut + idx); }
      }
      else if (FC == kOnePattern32) {
        const uint8_t* srow = pat.src.pixel_data + intptr_t(prj) * pat.src.stride;
        const bool full = (m[0] && m[1] && m[2] && m[3]);
        if (pc.vec && full && ((uintptr_t(srow) + size_t(pc.c[0]) * 4) & 15u) == 0u) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(srow + size_t(pc.c[0]) * 4));
          s[0] = q.x; s[1] = q.y; s[2] = q.z; s[3] = q.w;
        }
        else {
          #pragma unroll
          for (int i = 0; i < 4; i++) if (m[i]) s[i] = __ldg(reinterpret_cast<const uint32_t*>(srow) + pc.c[i]);
        }
        if (P.src_format == B2DGPU_FORMAT_XRGB32) {
          #pragma unroll
          for (int i = 0; i < 4; i++) s[i] = adapt_src_xrgb32(s[i]);
        }
      }
      else {
        fetch4(env, uint32_t(x), yy, m, s);
      }

      uint32_t d[4] = { v[j].x, v[j].y, v[j].z, v[j].w };
      composite4(comp_op, d, s, m, opaque);
      *reinterpret_cast<uint4*>(ptr + size_t(j) * P.dst_stride) = make_uint4(d[0], d[1], d[2], d[3]);
    }
  }
  if (P.pixel_counter && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.pixel_counter, P.pixels);
}

template<int FC, bool STAGED>
static int launch_one(const StreamOneParams& P, int sm_count, cudaStream_t s) {
  static int per_sm = 0;
  if (!per_sm) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_stream_one<FC, STAGED>, 256, 0) != cudaSuccess || n < 1) n = 2;
    per_sm = n;
  }
  const long long ncb = (P.x1 + 127) / 128 - P.x0 / 128, nrq = (P.y1 - P.y0 + 3) / 4;
  const long long want = (ncb * nrq + 7) / 8;
  const long long cap = (long long)sm_count * per_sm;
  const int grid = int(want < cap ? (want < 1 ? 1 : want) : cap);
  k_stream_one<FC, STAGED><<<grid, 256, 0, s>>>(P);
  return 1;
}

// Which specialisation serves a (fetch type, source format) pair.
static int stream_one_class(uint32_t ft, uint32_t src_format) {
  if (ft == B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD || ft == B2DGPU_FETCH_GRADIENT_LINEAR_NN_ROR) return kOneLinear;
  if (ft == B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD || ft == B2DGPU_FETCH_GRADIENT_RADIAL_NN_ROR) return kOneRadial;
  if (ft == B2DGPU_FETCH_GRADIENT_CONIC_NN) return kOneConic;
  if (ft >= B2DGPU_FETCH_PATTERN_ALIGNED_BLIT && ft <= B2DGPU_FETCH_PATTERN_ALIGNED_ROR &&
      (src_format == B2DGPU_FORMAT_PRGB32 || src_format == B2DGPU_FORMAT_XRGB32)) return kOnePattern32;
  return kOneGeneric;
}

int launch_stream_one(const StreamOneParams& P, int sm_count, cudaStream_t s) {
  if (P.x0 >= P.x1 || P.y0 >= P.y1) return 0;
  // stage_lut: entries of the gradient table when it fits the shared-memory copy (0: look it up through LDG)
  const bool staged = P.stage_lut != 0u && P.stage_lut <= uint32_t(kLutSmemEntries);
  switch (stream_one_class(P.fetch_type, P.src_format)) {
    case kOneLinear:    return staged ? launch_one<kOneLinear, true>(P, sm_count, s) : launch_one<kOneLinear, false>(P, sm_count, s);
    case kOneRadial:    return staged ? launch_one<kOneRadial, true>(P, sm_count, s) : launch_one<kOneRadial, false>(P, sm_count, s);
    case kOneConic:     return staged ? launch_one<kOneConic, true>(P, sm_count, s) : launch_one<kOneConic, false>(P, sm_count, s);
    case kOnePattern32: return launch_one<kOnePattern32, false>(P, sm_count, s);
    default:            return launch_one<kOneGeneric, false>(P, sm_count, s);
  }
}

} // namespace b2d
