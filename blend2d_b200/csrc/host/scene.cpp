// scene.cpp - replays a b2d_scene through the host front end (include/b2d_host.h), one render call per fill, the way
// bl_bench issues them (a style object is created per call: bl_bench_backend_blend2d.cpp:92-158).
#include "../../../include/b2d_host.h"
#include "../../../include/b2d_scene.h"

#include <string.h>

extern "C" B2DGPU_API b2dgpu_result b2d_scene_replay(b2d_context* ctx, const b2d_scene* sc, uint32_t first, uint32_t count) {
  if (!ctx || !sc) return B2DGPU_ERROR_INVALID_VALUE;
  b2d_image* tex = nullptr;
  b2dgpu_result r = B2DGPU_SUCCESS;
  if (sc->texture) {
    r = b2d_image_create(sc->texture_w, sc->texture_h, B2DGPU_FORMAT_PRGB32, &tex);
    if (r) return r;
    b2dgpu_image_data d;
    b2d_image_get_data(tex, &d);
    memcpy(d.pixel_data, sc->texture, size_t(sc->texture_w) * size_t(sc->texture_h) * 4);
  }

  uint32_t end = first + count < sc->fill_count ? first + count : sc->fill_count;
  for (uint32_t i = first; i < end && !r; i++) {
    const b2d_scene_fill& f = sc->fills[i];
    // The host mirror has no font engine and no stroker: text and strokes only exist behind the real Blend2D frontend
    // (shim/), which feeds this same runtime.
    if (f.geom == B2D_SCENE_GEOM_TEXT || f.stroke_width > 0.0) { r = B2DGPU_ERROR_NOT_IMPLEMENTED; break; }
    b2d_gradient* g = nullptr;
    b2d_pattern* p = nullptr;
    b2d_context_set_comp_op(ctx, f.comp_op);
    b2d_context_set_fill_rule(ctx, f.fill_rule);

    if (f.style == B2D_SCENE_STYLE_SOLID) r = b2d_context_set_fill_style_rgba32(ctx, f.rgba32);
    else if (f.style == B2D_SCENE_STYLE_PATTERN) {
      b2d_context_set_hint(ctx, B2D_HINT_PATTERN_QUALITY, f.quality);
      r = tex ? b2d_pattern_create(tex, nullptr, f.extend, f.values, &p) : B2DGPU_ERROR_INVALID_VALUE;
      if (!r) r = b2d_context_set_fill_style_pattern(ctx, p);
    }
    else {
      b2d_context_set_hint(ctx, B2D_HINT_GRADIENT_QUALITY, f.quality);
      r = b2d_gradient_create(f.style - B2D_SCENE_STYLE_LINEAR, f.values, f.extend,
                              reinterpret_cast<const b2d_gradient_stop*>(sc->stops + f.stop_offset), f.stop_count, nullptr, &g);
      if (!r) r = b2d_context_set_fill_style_gradient(ctx, g);
    }

    if (!r && f.has_transform) {
      double rot[3] = { f.angle, f.cx, f.cy };
      r = b2d_context_apply_transform_op(ctx, 6, rot);
    }

    if (!r) {
      switch (f.geom) {
        case B2D_SCENE_GEOM_RECT_I: r = b2d_context_fill_rect_i(ctx, int32_t(f.rect[0]), int32_t(f.rect[1]), int32_t(f.rect[2]), int32_t(f.rect[3])); break;
        case B2D_SCENE_GEOM_RECT_D: r = b2d_context_fill_rect_d(ctx, f.rect[0], f.rect[1], f.rect[2], f.rect[3]); break;
        case B2D_SCENE_GEOM_POLYGON: r = b2d_context_fill_polygon_d(ctx, sc->vertices + size_t(f.vtx_offset) * 2, f.vtx_count); break;
        default: r = b2d_context_fill_path_d(ctx, 0.0, 0.0, sc->path_cmds + f.vtx_offset, sc->vertices + size_t(f.vtx_offset) * 2, f.vtx_count); break;
      }
    }
    if (f.has_transform) b2d_context_apply_transform_op(ctx, 0, nullptr);
    // The context retains what its queued FetchData points to (LUT, pixels) until the batch is flushed.
    if (g) b2d_gradient_destroy(g);
    if (p) b2d_pattern_destroy(p);
  }
  if (tex) b2d_image_destroy(tex);
  return r;
}
