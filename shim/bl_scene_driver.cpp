// bl_scene_driver.cpp - replays a b2d_scene (include/b2d_scene.h) through Blend2D's PUBLIC C API (blend2d/blend2d.h).
//
// It is a Blend2D application, nothing else: it knows BLImage / BLContext / BLPath / BLGradient / BLPattern / BLFont and
// a create-flags word.  It is built twice from this one source:
//   shim/_build/libgpu_scene_driver.so    against shim/_build/libblend2d_gpu.so  (bench.py's `e2e` and capture legs:
//                                         BLContextCreateInfo.flags = 0x10000000 selects the B200 pipeline runtime)
//   oracle/_ref/libref_scene_driver.so    against the UNMODIFIED reference (bench.py's `--impl reference` arm and
//                                         `cpu_baseline` leg, and the tests' way of drawing big scenes on the CPU)
// so both arms of every comparison execute the same calls.  bench.py generates the scene arrays with numpy; replaying
// them in C keeps Python out of every timed region, the way bl_bench drives Blend2D
// (blend2d-testing/bench/bl_bench_backend_blend2d.cpp).  Timing follows bl_bench: the clock stops after
// flush(BL_CONTEXT_FLUSH_SYNC) (blend2d-testing/bench/bl_bench_backend.cpp:49-91).
#include <blend2d/blend2d.h>

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <new>

#include "../include/b2d_scene.h"

#define DRV_API extern "C" __attribute__((visibility("default")))

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return double(ts.tv_sec) + double(ts.tv_nsec) * 1e-9;
}

struct bl_scene_session {
  const b2d_scene* scene;
  BLImageCore img, tex;
  BLFontFaceCore face;
  BLFontCore font;
  BLContextCore ctx;
  bool has_tex, has_font, has_ctx;
  int w, h;
  uint32_t format;
};

static BLResult set_style(BLContextCore* ctx, const b2d_scene* sc, const b2d_scene_fill& f, BLImageCore* tex, bool stroke) {
  BLResult r = BL_SUCCESS;
  if (f.style == B2D_SCENE_STYLE_SOLID) {
    return stroke ? bl_context_set_stroke_style_rgba32(ctx, f.rgba32) : bl_context_set_fill_style_rgba32(ctx, f.rgba32);
  }
  if (f.style == B2D_SCENE_STYLE_PATTERN) {
    if (!tex) return BL_ERROR_INVALID_VALUE;
    bl_context_set_hint(ctx, BL_CONTEXT_HINT_PATTERN_QUALITY, f.quality);
    BLPatternCore p;
    BLMatrix2D m(f.values[0], f.values[1], f.values[2], f.values[3], f.values[4], f.values[5]);
    r = bl_pattern_init_as(&p, tex, nullptr, BLExtendMode(f.extend), &m);
    if (r == BL_SUCCESS) r = stroke ? bl_context_set_stroke_style(ctx, &p) : bl_context_set_fill_style(ctx, &p);
    bl_pattern_destroy(&p);
    return r;
  }
  bl_context_set_hint(ctx, BL_CONTEXT_HINT_GRADIENT_QUALITY, f.quality);
  BLGradientCore g;
  double values[6];
  memcpy(values, f.values, sizeof(values));
  r = bl_gradient_init_as(&g, BLGradientType(f.style - B2D_SCENE_STYLE_LINEAR), values, BLExtendMode(f.extend),
                          reinterpret_cast<const BLGradientStop*>(sc->stops + f.stop_offset), f.stop_count, nullptr);
  if (r == BL_SUCCESS) r = stroke ? bl_context_set_stroke_style(ctx, &g) : bl_context_set_fill_style(ctx, &g);
  bl_gradient_destroy(&g);
  return r;
}

static BLResult replay(bl_scene_session* s, const b2d_scene* sc, uint32_t first, uint32_t count) {
  BLContextCore* ctx = &s->ctx;
  BLResult r = BL_SUCCESS;
  uint32_t end = first + count < sc->fill_count ? first + count : sc->fill_count;
  if (first > end) first = end;
  for (uint32_t i = first; i < end && r == BL_SUCCESS; i++) {
    const b2d_scene_fill& f = sc->fills[i];
    const bool stroke = f.stroke_width > 0.0;
    bl_context_set_comp_op(ctx, BLCompOp(f.comp_op));
    bl_context_set_fill_rule(ctx, BLFillRule(f.fill_rule));
    if (stroke) bl_context_set_stroke_width(ctx, f.stroke_width);
    r = set_style(ctx, sc, f, s->has_tex ? &s->tex : nullptr, stroke);

    if (r == BL_SUCCESS && f.has_transform) {
      double rot[3] = { f.angle, f.cx, f.cy };
      r = bl_context_apply_transform_op(ctx, BL_TRANSFORM_OP_ROTATE_PT, rot);
    }

    if (r == BL_SUCCESS) {
      switch (f.geom) {
        case B2D_SCENE_GEOM_RECT_I: {
          BLRectI rc(int(f.rect[0]), int(f.rect[1]), int(f.rect[2]), int(f.rect[3]));
          r = stroke ? bl_context_stroke_rect_i(ctx, &rc) : bl_context_fill_rect_i(ctx, &rc);
          break;
        }
        case B2D_SCENE_GEOM_RECT_D: {
          BLRect rc(f.rect[0], f.rect[1], f.rect[2], f.rect[3]);
          r = stroke ? bl_context_stroke_rect_d(ctx, &rc) : bl_context_fill_rect_d(ctx, &rc);
          break;
        }
        case B2D_SCENE_GEOM_POLYGON: {
          BLArrayView<BLPoint> view;
          view.reset(reinterpret_cast<const BLPoint*>(sc->vertices + size_t(f.vtx_offset) * 2), f.vtx_count);
          r = stroke ? bl_context_stroke_geometry(ctx, BL_GEOMETRY_TYPE_POLYGOND, &view) : bl_context_fill_geometry(ctx, BL_GEOMETRY_TYPE_POLYGOND, &view);
          break;
        }
        case B2D_SCENE_GEOM_TEXT: {
          if (!s->has_font || !sc->text || uint64_t(f.vtx_offset) + f.vtx_count > sc->text_size) { r = BL_ERROR_INVALID_VALUE; break; }
          BLPoint origin(f.rect[0], f.rect[1]);
          r = stroke ? bl_context_stroke_utf8_text_d(ctx, &origin, &s->font, sc->text + f.vtx_offset, f.vtx_count)
                     : bl_context_fill_utf8_text_d(ctx, &origin, &s->font, sc->text + f.vtx_offset, f.vtx_count);
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
          if (r == BL_SUCCESS) r = stroke ? bl_context_stroke_path_d(ctx, &origin, &path) : bl_context_fill_path_d(ctx, &origin, &path);
          bl_path_destroy(&path);
          break;
        }
      }
    }
    if (f.has_transform) bl_context_apply_transform_op(ctx, BL_TRANSFORM_OP_RESET, nullptr);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Sessions: one image + one context, kept open across steps (what an application that renders frames does).
// ---------------------------------------------------------------------------------------------------------------
DRV_API uint32_t bl_scene_close(bl_scene_session* s) {
  if (!s) return BL_ERROR_INVALID_VALUE;
  if (s->has_ctx) { bl_context_end(&s->ctx); bl_context_destroy(&s->ctx); }
  if (s->has_font) { bl_font_destroy(&s->font); bl_font_face_destroy(&s->face); }
  if (s->has_tex) bl_image_destroy(&s->tex);
  bl_image_destroy(&s->img);
  delete s;
  return BL_SUCCESS;
}

// create_flags: BLContextCreateFlags (0x1 = DISABLE_JIT; 0x10000000 = GPU pipeline runtime on builds that have the
// shim); reserved = BLContextCreateInfo::reserved[0] (CUDA device ordinal for the GPU runtime).
DRV_API uint32_t bl_scene_open(const b2d_scene* sc, int w, int h, uint32_t format, uint32_t create_flags, uint32_t thread_count,
                               uint32_t command_queue_limit, uint32_t reserved, bl_scene_session** out) {
  if (!sc || !out) return BL_ERROR_INVALID_VALUE;
  *out = nullptr;
  bl_scene_session* s = new (std::nothrow) bl_scene_session();
  if (!s) return BL_ERROR_OUT_OF_MEMORY;
  s->scene = sc; s->has_tex = s->has_font = s->has_ctx = false; s->w = w; s->h = h; s->format = format;
  BLResult r = bl_image_init_as(&s->img, w, h, BLFormat(format));
  if (r != BL_SUCCESS) { delete s; return r; }
  if (sc->texture) {
    r = bl_image_init_as(&s->tex, sc->texture_w, sc->texture_h, BL_FORMAT_PRGB32);
    if (r == BL_SUCCESS) {
      BLImageData td;
      bl_image_make_mutable(&s->tex, &td);
      for (int y = 0; y < sc->texture_h; y++)
        memcpy(static_cast<uint8_t*>(td.pixel_data) + intptr_t(y) * td.stride, sc->texture + size_t(y) * sc->texture_w, size_t(sc->texture_w) * 4);
      s->has_tex = true;
    }
  }
  if (r == BL_SUCCESS && sc->font_file) {
    bl_font_face_init(&s->face);
    bl_font_init(&s->font);
    s->has_font = true;
    r = bl_font_face_create_from_file(&s->face, sc->font_file, BL_FILE_READ_NO_FLAGS);
    if (r == BL_SUCCESS) r = bl_font_create_from_face(&s->font, &s->face, float(sc->font_size));
  }
  if (r == BL_SUCCESS) {
    BLContextCreateInfo cci {};
    cci.flags = create_flags;
    cci.thread_count = thread_count;
    cci.command_queue_limit = command_queue_limit;
    cci.reserved[0] = reserved;
    r = bl_context_init_as(&s->ctx, &s->img, &cci);
    s->has_ctx = r == BL_SUCCESS;
  }
  if (r != BL_SUCCESS) { bl_scene_close(s); return r; }
  *out = s;
  return BL_SUCCESS;
}

// clear_all + flush(SYNC): never inside a timed region.
DRV_API uint32_t bl_scene_clear(bl_scene_session* s) {
  bl_context_set_comp_op(&s->ctx, BL_COMP_OP_SRC_OVER);
  BLResult r = bl_context_clear_all(&s->ctx);
  if (r == BL_SUCCESS) r = bl_context_flush(&s->ctx, BL_CONTEXT_FLUSH_SYNC);
  return r;
}

DRV_API uint32_t bl_scene_draw(bl_scene_session* s, uint32_t first, uint32_t count) { return replay(s, s->scene, first, count); }
DRV_API uint32_t bl_scene_flush(bl_scene_session* s, int sync) { return bl_context_flush(&s->ctx, sync ? BL_CONTEXT_FLUSH_SYNC : BL_CONTEXT_FLUSH_NO_FLAGS); }

// draw + flush(SYNC), timed here (seconds) so that no Python runs between the first render call and the end of the flush.
DRV_API uint32_t bl_scene_step(bl_scene_session* s, uint32_t first, uint32_t count, double* seconds_out) {
  double t0 = now_s();
  BLResult r = replay(s, s->scene, first, count);
  BLResult r2 = bl_context_flush(&s->ctx, BL_CONTEXT_FLUSH_SYNC);
  double t1 = now_s();
  if (seconds_out) *seconds_out = t1 - t0;
  return r != BL_SUCCESS ? r : r2;
}

DRV_API uint32_t bl_scene_error_flags(bl_scene_session* s) {
  uint32_t v = 0;
  bl_object_get_property_uint32(&s->ctx, "accumulated_error_flags", 23, &v);
  return v;
}

// Host pixels of the session's image (coherent after a flush(SYNC)).
DRV_API uint32_t bl_scene_read_pixels(bl_scene_session* s, void* pixels_out, intptr_t out_stride) {
  BLImageData d;
  BLResult r = bl_image_get_data(&s->img, &d);
  if (r != BL_SUCCESS) return r;
  size_t row = size_t(s->w) * (s->format == BL_FORMAT_A8 ? 1 : 4);
  for (int y = 0; y < s->h; y++)
    memcpy(static_cast<uint8_t*>(pixels_out) + intptr_t(y) * out_stride, static_cast<const uint8_t*>(d.pixel_data) + intptr_t(y) * d.stride, row);
  return BL_SUCCESS;
}

// XOR of all pixel words: a cheap "the frame was read on the host" consumer for the many-frames workload.
DRV_API uint32_t bl_scene_checksum(bl_scene_session* s, uint32_t* out) {
  BLImageData d;
  BLResult r = bl_image_get_data(&s->img, &d);
  if (r != BL_SUCCESS) return r;
  uint32_t x = 0;
  const size_t words = size_t(s->w) * (s->format == BL_FORMAT_A8 ? 1 : 4) / 4;
  for (int y = 0; y < s->h; y++) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(d.pixel_data) + intptr_t(y) * d.stride);
    for (size_t i = 0; i < words; i++) x ^= p[i];
  }
  *out = x;
  return BL_SUCCESS;
}

// Frame-sharded workload (config 5-ii): fills [f * fills_per_frame, (f + 1) * fills_per_frame) are frame f.  Every
// frame is cleared, drawn and flushed (SYNC) - the frame is in the BLImage's host pixels when flush returns - one after
// the other on this session.  Every 32nd frame is also read back by the CPU (XOR checksum: the value both arms must
// agree on); reading all of them made the loop measure a scalar XOR over 8 MB (0.7 ms) rather than the renderer.
// seconds_out = wall time of the whole pass; checksum_out = XOR over the checked frames.
DRV_API uint32_t bl_scene_run_frames(bl_scene_session* s, uint32_t first_frame, uint32_t frame_count, uint32_t fills_per_frame,
                                     double* seconds_out, uint32_t* checksum_out) {
  BLResult r = BL_SUCCESS;
  uint32_t x = 0;
  double t0 = now_s();
  for (uint32_t f = first_frame; f < first_frame + frame_count && r == BL_SUCCESS; f++) {
    bl_context_set_comp_op(&s->ctx, BL_COMP_OP_SRC_OVER);
    bl_context_clear_all(&s->ctx);
    r = replay(s, s->scene, f * fills_per_frame, fills_per_frame);
    if (r == BL_SUCCESS) r = bl_context_flush(&s->ctx, BL_CONTEXT_FLUSH_SYNC);
    uint32_t c = 0;
    if (r == BL_SUCCESS && f % 32u == 0u) r = bl_scene_checksum(s, &c);
    x ^= c;
  }
  double t1 = now_s();
  if (seconds_out) *seconds_out = t1 - t0;
  if (checksum_out) *checksum_out = x;
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// One-call forms
// ---------------------------------------------------------------------------------------------------------------
// Renders fills [first, first + count) `steps` times into a w x h canvas.  pixels_out (may be null) receives the canvas
// after the LAST step.  seconds_out[s] = wall time of step s, measured from the first render call to the return of
// flush(SYNC).  The canvas is cleared (untimed) before every step.
DRV_API uint32_t bl_scene_run(const b2d_scene* sc, uint32_t first, uint32_t count, int w, int h, uint32_t format, uint32_t create_flags,
                              uint32_t thread_count, uint32_t steps, double* seconds_out, void* pixels_out, intptr_t out_stride) {
  bl_scene_session* s = nullptr;
  BLResult r = bl_scene_open(sc, w, h, format, create_flags, thread_count, 0, 0, &s);
  if (r != BL_SUCCESS) return r;
  for (uint32_t k = 0; k < steps && r == BL_SUCCESS; k++) {
    r = bl_scene_clear(s);
    double dt = 0.0;
    if (r == BL_SUCCESS) r = bl_scene_step(s, first, count, &dt);
    if (seconds_out) seconds_out[k] = dt;
  }
  if (r == BL_SUCCESS && pixels_out) r = bl_scene_read_pixels(s, pixels_out, out_stride);
  bl_scene_close(s);
  return r;
}

// Number of pixels each fill writes (coverage != 0), summed over fills [first, first + count): every fill is drawn
// alone, opaque white with SRC_OVER, onto a cleared A8 canvas and the non-zero bytes are counted.  This is the
// pixel count behind the Mpix/s metric on the CPU arm; it is never inside a timed region.
DRV_API uint32_t bl_scene_count_pixels(const b2d_scene* sc, uint32_t first, uint32_t count, int w, int h, uint64_t* pixels_out) {
  b2d_scene one = *sc;
  b2d_scene_fill f;
  one.fills = &f; one.fill_count = 1;
  bl_scene_session* s = nullptr;
  BLResult r = bl_scene_open(&one, w, h, BL_FORMAT_A8, BL_CONTEXT_CREATE_FLAG_DISABLE_JIT, 0, 0, 0, &s);
  if (r != BL_SUCCESS) return r;
  uint64_t total = 0;
  BLImageData d;
  bl_image_get_data(&s->img, &d);
  uint32_t end = first + count < sc->fill_count ? first + count : sc->fill_count;
  r = bl_scene_clear(s);
  for (uint32_t i = first; i < end && r == BL_SUCCESS; i++) {
    f = sc->fills[i];
    f.style = B2D_SCENE_STYLE_SOLID; f.rgba32 = 0xFFFFFFFFu; f.comp_op = BL_COMP_OP_SRC_OVER;
    // Region the fill can touch: the hull of its vertices for untransformed polygons / paths / rectangles (curves stay
    // inside the hull of their control points), the whole canvas otherwise.  Only that region is scanned and cleared.
    int x0 = 0, y0 = 0, x1 = w, y1 = h;
    if (!f.has_transform && !(f.stroke_width > 0.0) && f.geom != B2D_SCENE_GEOM_TEXT) {
      double mnx, mny, mxx, mxy;
      if (f.geom == B2D_SCENE_GEOM_POLYGON || f.geom == B2D_SCENE_GEOM_PATH) {
        mnx = mny = 1e300; mxx = mxy = -1e300;
        const double* v = sc->vertices + size_t(f.vtx_offset) * 2;
        for (uint32_t k = 0; k < f.vtx_count; k++) {
          if (v[k * 2] < mnx) mnx = v[k * 2]; if (v[k * 2] > mxx) mxx = v[k * 2];
          if (v[k * 2 + 1] < mny) mny = v[k * 2 + 1]; if (v[k * 2 + 1] > mxy) mxy = v[k * 2 + 1];
        }
      }
      else { mnx = f.rect[0]; mny = f.rect[1]; mxx = f.rect[0] + f.rect[2]; mxy = f.rect[1] + f.rect[3]; }
      if (mnx - 2 > 0) x0 = mnx - 2 < w ? int(mnx - 2) : w;
      if (mny - 2 > 0) y0 = mny - 2 < h ? int(mny - 2) : h;
      if (mxx + 3 < w) x1 = mxx + 3 > 0 ? int(mxx + 3) : 0;
      if (mxy + 3 < h) y1 = mxy + 3 > 0 ? int(mxy + 3) : 0;
    }
    if (x0 >= x1 || y0 >= y1) continue;
    r = replay(s, &one, 0, 1);
    bl_context_flush(&s->ctx, BL_CONTEXT_FLUSH_SYNC);
    for (int y = y0; y < y1; y++) {
      const uint8_t* row = static_cast<const uint8_t*>(d.pixel_data) + intptr_t(y) * d.stride;
      for (int x = x0; x < x1; x++) total += row[x] != 0;
    }
    BLRectI box(x0, y0, x1 - x0, y1 - y0);
    bl_context_clear_rect_i(&s->ctx, &box);
  }
  bl_scene_close(s);
  if (pixels_out) *pixels_out = total;
  return r;
}
