// ref_scene_driver.cpp - replays a b2d_scene (include/b2d_scene.h) through the UNMODIFIED reference's public C API.
//
// TEST / BASELINE INFRASTRUCTURE: built by oracle/Makefile.ref into oracle/_ref/libref_scene_driver.so against the
// headers under /root/reference and linked to oracle/_ref/libblend2d_ref.so.  Used by bench.py's `--impl reference` arm
// and `cpu_baseline` leg (the reference's own CPU implementation of the path, async multithreaded rendering with
// BLContextCreateInfo.thread_count - blend2d/core/context.h:325-368) and by tests as a fast way to draw big scenes.
// Timing follows bl_bench: the clock stops after flush(BL_CONTEXT_FLUSH_SYNC)
// (blend2d-testing/bench/bl_bench_backend.cpp:49-91).
#include <blend2d/blend2d.h>

#include <stdint.h>
#include <string.h>
#include <time.h>

#include "../include/b2d_scene.h"

#define REF_API extern "C" __attribute__((visibility("default")))

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return double(ts.tv_sec) + double(ts.tv_nsec) * 1e-9;
}

static BLResult replay(BLContextCore* ctx, const b2d_scene* sc, uint32_t first, uint32_t count, BLImageCore* tex) {
  BLResult r = BL_SUCCESS;
  uint32_t end = first + count < sc->fill_count ? first + count : sc->fill_count;
  for (uint32_t i = first; i < end && r == BL_SUCCESS; i++) {
    const b2d_scene_fill& f = sc->fills[i];
    bl_context_set_comp_op(ctx, BLCompOp(f.comp_op));
    bl_context_set_fill_rule(ctx, BLFillRule(f.fill_rule));

    if (f.style == B2D_SCENE_STYLE_SOLID) {
      r = bl_context_set_fill_style_rgba32(ctx, f.rgba32);
    }
    else if (f.style == B2D_SCENE_STYLE_PATTERN) {
      bl_context_set_hint(ctx, BL_CONTEXT_HINT_PATTERN_QUALITY, f.quality);
      BLPatternCore p;
      BLMatrix2D m(f.values[0], f.values[1], f.values[2], f.values[3], f.values[4], f.values[5]);
      r = bl_pattern_init_as(&p, tex, nullptr, BLExtendMode(f.extend), &m);
      if (r == BL_SUCCESS) r = bl_context_set_fill_style(ctx, &p);
      bl_pattern_destroy(&p);
    }
    else {
      bl_context_set_hint(ctx, BL_CONTEXT_HINT_GRADIENT_QUALITY, f.quality);
      BLGradientCore g;
      double values[6];
      memcpy(values, f.values, sizeof(values));
      r = bl_gradient_init_as(&g, BLGradientType(f.style - B2D_SCENE_STYLE_LINEAR), values, BLExtendMode(f.extend),
                              reinterpret_cast<const BLGradientStop*>(sc->stops + f.stop_offset), f.stop_count, nullptr);
      if (r == BL_SUCCESS) r = bl_context_set_fill_style(ctx, &g);
      bl_gradient_destroy(&g);
    }

    if (r == BL_SUCCESS && f.has_transform) {
      double rot[3] = { f.angle, f.cx, f.cy };
      r = bl_context_apply_transform_op(ctx, BL_TRANSFORM_OP_ROTATE_PT, rot);
    }

    if (r == BL_SUCCESS) {
      switch (f.geom) {
        case B2D_SCENE_GEOM_RECT_I: {
          BLRectI rc(int(f.rect[0]), int(f.rect[1]), int(f.rect[2]), int(f.rect[3]));
          r = bl_context_fill_rect_i(ctx, &rc);
          break;
        }
        case B2D_SCENE_GEOM_RECT_D: {
          BLRect rc(f.rect[0], f.rect[1], f.rect[2], f.rect[3]);
          r = bl_context_fill_rect_d(ctx, &rc);
          break;
        }
        case B2D_SCENE_GEOM_POLYGON: {
          BLArrayView<BLPoint> view;
          view.reset(reinterpret_cast<const BLPoint*>(sc->vertices + size_t(f.vtx_offset) * 2), f.vtx_count);
          r = bl_context_fill_geometry(ctx, BL_GEOMETRY_TYPE_POLYGOND, &view);
          break;
        }
        default: {
          // A BLPath is built per call from the command / vertex arrays, like an application would.
          BLPathCore path;
          bl_path_init(&path);
          const double* v = sc->vertices + size_t(f.vtx_offset) * 2;
          const uint8_t* c = sc->path_cmds + f.vtx_offset;
          for (uint32_t k = 0; k < f.vtx_count && r == BL_SUCCESS;) {
            switch (c[k]) {
              case BL_PATH_CMD_MOVE: r = bl_path_move_to(&path, v[k * 2], v[k * 2 + 1]); k += 1; break;
              case BL_PATH_CMD_ON: r = bl_path_line_to(&path, v[k * 2], v[k * 2 + 1]); k += 1; break;
              case BL_PATH_CMD_QUAD: r = bl_path_quad_to(&path, v[k * 2], v[k * 2 + 1], v[k * 2 + 2], v[k * 2 + 3]); k += 2; break;
              case BL_PATH_CMD_CUBIC: r = bl_path_cubic_to(&path, v[k * 2], v[k * 2 + 1], v[k * 2 + 2], v[k * 2 + 3], v[k * 2 + 4], v[k * 2 + 5]); k += 3; break;
              case BL_PATH_CMD_CLOSE: r = bl_path_close(&path); k += 1; break;
              default: k += 1; break;
            }
          }
          BLPoint origin(0.0, 0.0);
          if (r == BL_SUCCESS) r = bl_context_fill_path_d(ctx, &origin, &path);
          bl_path_destroy(&path);
          break;
        }
      }
    }
    if (f.has_transform) bl_context_apply_transform_op(ctx, BL_TRANSFORM_OP_RESET, nullptr);
  }
  return r;
}

// Renders fills [first, first + count) `steps` times into a w x h canvas with `thread_count` worker threads.
// pixels_out (may be null) receives the canvas after the LAST step.  seconds_out[s] = wall time of step s, measured
// from the first render call to the return of flush(SYNC).  The canvas is cleared (untimed) before every step.
REF_API uint32_t ref_scene_run(const b2d_scene* sc, uint32_t first, uint32_t count, int w, int h, uint32_t format,
                               uint32_t thread_count, uint32_t steps, double* seconds_out, void* pixels_out, intptr_t out_stride) {
  BLImageCore img, tex;
  BLResult r = bl_image_init_as(&img, w, h, BLFormat(format));
  if (r != BL_SUCCESS) return r;
  bool has_tex = false;
  if (sc->texture) {
    r = bl_image_init_as(&tex, sc->texture_w, sc->texture_h, BL_FORMAT_PRGB32);
    if (r == BL_SUCCESS) {
      BLImageData td;
      bl_image_make_mutable(&tex, &td);
      for (int y = 0; y < sc->texture_h; y++)
        memcpy(static_cast<uint8_t*>(td.pixel_data) + intptr_t(y) * td.stride, sc->texture + size_t(y) * sc->texture_w, size_t(sc->texture_w) * 4);
      has_tex = true;
    }
  }

  BLContextCreateInfo cci {};
  cci.flags = BL_CONTEXT_CREATE_FLAG_DISABLE_JIT;
  cci.thread_count = thread_count;
  BLContextCore ctx;
  if (r == BL_SUCCESS) r = bl_context_init_as(&ctx, &img, &cci);
  if (r == BL_SUCCESS) {
    for (uint32_t s = 0; s < steps && r == BL_SUCCESS; s++) {
      bl_context_set_comp_op(&ctx, BL_COMP_OP_SRC_OVER);
      bl_context_clear_all(&ctx);
      bl_context_flush(&ctx, BL_CONTEXT_FLUSH_SYNC);
      double t0 = now_s();
      r = replay(&ctx, sc, first, count, has_tex ? &tex : nullptr);
      bl_context_flush(&ctx, BL_CONTEXT_FLUSH_SYNC);
      double t1 = now_s();
      if (seconds_out) seconds_out[s] = t1 - t0;
    }
    bl_context_end(&ctx);
    bl_context_destroy(&ctx);
  }
  if (r == BL_SUCCESS && pixels_out) {
    BLImageData d;
    bl_image_get_data(&img, &d);
    size_t row = size_t(w) * (format == BL_FORMAT_A8 ? 1 : 4);
    for (int y = 0; y < h; y++)
      memcpy(static_cast<uint8_t*>(pixels_out) + intptr_t(y) * out_stride, static_cast<const uint8_t*>(d.pixel_data) + intptr_t(y) * d.stride, row);
  }
  if (has_tex) bl_image_destroy(&tex);
  bl_image_destroy(&img);
  return r;
}

// Number of pixels each fill writes (coverage != 0), summed over fills [first, first + count): every fill is drawn
// alone, opaque white with SRC_COPY, onto a cleared A8 canvas and the non-zero bytes are counted.  This is the
// pixel count behind the Mpix/s metric; it is never inside a timed region.
REF_API uint32_t ref_scene_count_pixels(const b2d_scene* sc, uint32_t first, uint32_t count, int w, int h, uint64_t* pixels_out) {
  BLImageCore img;
  BLResult r = bl_image_init_as(&img, w, h, BL_FORMAT_A8);
  if (r != BL_SUCCESS) return r;
  BLContextCreateInfo cci {};
  cci.flags = BL_CONTEXT_CREATE_FLAG_DISABLE_JIT;
  BLContextCore ctx;
  r = bl_context_init_as(&ctx, &img, &cci);
  uint64_t total = 0;
  if (r == BL_SUCCESS) {
    BLImageData d;
    bl_image_get_data(&img, &d);
    uint32_t end = first + count < sc->fill_count ? first + count : sc->fill_count;
    b2d_scene one = *sc;
    for (uint32_t i = first; i < end && r == BL_SUCCESS; i++) {
      b2d_scene_fill f = sc->fills[i];
      f.style = B2D_SCENE_STYLE_SOLID; f.rgba32 = 0xFFFFFFFFu; f.comp_op = BL_COMP_OP_SRC_OVER;
      one.fills = &f; one.fill_count = 1;
      bl_context_set_comp_op(&ctx, BL_COMP_OP_SRC_OVER);
      bl_context_clear_all(&ctx);
      r = replay(&ctx, &one, 0, 1, nullptr);
      bl_context_flush(&ctx, BL_CONTEXT_FLUSH_SYNC);
      for (int y = 0; y < h; y++) {
        const uint8_t* row = static_cast<const uint8_t*>(d.pixel_data) + intptr_t(y) * d.stride;
        for (int x = 0; x < w; x++) total += row[x] != 0;
      }
    }
    bl_context_end(&ctx);
    bl_context_destroy(&ctx);
  }
  bl_image_destroy(&img);
  if (pixels_out) *pixels_out = total;
  return r;
}
