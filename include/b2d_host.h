/*
 * b2d_host.h - host-side mirror of the slice of Blend2D's BLContext front end that feeds the rendering hot path.
 *
 * Blend2D itself stays the host product; this is the stand-alone equivalent of what its raster context does ABOVE the
 * drop-in boundary (include/b2dgpu.h) for the calls the parity tests and bench.py make: resolve style + comp-op +
 * alpha into a Signature / FetchData / geometry command, queue it, and hand the batch to the GPU runtime on flush.
 * Names, argument meaning and error behaviour follow the reference C API (blend2d/core/context.h:640-760,
 * image.h, gradient.h, pattern.h); results are BLResult-compatible.
 *
 * All pixels are produced by libb2dgpu's CUDA kernels.  There is no CPU rendering path here.
 */
#ifndef B2D_HOST_H_INCLUDED
#define B2D_HOST_H_INCLUDED

#include "b2dgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2d_image b2d_image;
typedef struct b2d_gradient b2d_gradient;
typedef struct b2d_pattern b2d_pattern;
typedef struct b2d_context b2d_context;

/* BLExtendMode (core/geometry.h). */
enum {
  B2D_EXTEND_PAD = 0, B2D_EXTEND_REPEAT = 1, B2D_EXTEND_REFLECT = 2,
  B2D_EXTEND_PAD_X_REPEAT_Y = 3, B2D_EXTEND_PAD_X_REFLECT_Y = 4, B2D_EXTEND_REPEAT_X_PAD_Y = 5,
  B2D_EXTEND_REPEAT_X_REFLECT_Y = 6, B2D_EXTEND_REFLECT_X_PAD_Y = 7, B2D_EXTEND_REFLECT_X_REPEAT_Y = 8
};
/* BLGradientType / BLGradientQuality / BLPatternQuality / BLFillRule / BLContextHint. */
enum { B2D_GRADIENT_LINEAR = 0, B2D_GRADIENT_RADIAL = 1, B2D_GRADIENT_CONIC = 2 };
enum { B2D_GRADIENT_QUALITY_NEAREST = 0, B2D_GRADIENT_QUALITY_SMOOTH = 1, B2D_GRADIENT_QUALITY_DITHER = 2 };
enum { B2D_PATTERN_QUALITY_NEAREST = 0, B2D_PATTERN_QUALITY_BILINEAR = 1 };
enum { B2D_FILL_RULE_NON_ZERO = 0, B2D_FILL_RULE_EVEN_ODD = 1 };
enum { B2D_HINT_RENDERING_QUALITY = 0, B2D_HINT_GRADIENT_QUALITY = 1, B2D_HINT_PATTERN_QUALITY = 2 };
/* BLPathCmd (core/path.h:22-39). */
enum { B2D_PATH_CMD_MOVE = 0, B2D_PATH_CMD_ON = 1, B2D_PATH_CMD_QUAD = 2, B2D_PATH_CMD_CONIC = 3, B2D_PATH_CMD_CUBIC = 4, B2D_PATH_CMD_CLOSE = 5, B2D_PATH_CMD_WEIGHT = 6 };

/* ---- BLImage (host pixel storage; stride == w * bpp like core/image.cpp:39-41) -------------------------------- */
B2DGPU_API b2dgpu_result b2d_image_create(int32_t w, int32_t h, uint32_t format, b2d_image** out);
B2DGPU_API b2dgpu_result b2d_image_destroy(b2d_image* img);
B2DGPU_API b2dgpu_result b2d_image_get_data(b2d_image* img, b2dgpu_image_data* out);

/* ---- BLGradient ----------------------------------------------------------------------------------------------- */
typedef struct b2d_gradient_stop { double offset; uint64_t rgba64; } b2d_gradient_stop;   /* BLGradientStop */
/* values: linear {x0,y0,x1,y1}, radial {x0,y0,x1,y1,r0,r1}, conic {x0,y0,angle,repeat}; matrix = 6 doubles or NULL.
 * Stops must be sorted by offset (what BLGradient keeps internally). */
B2DGPU_API b2dgpu_result b2d_gradient_create(uint32_t type, const double* values, uint32_t extend_mode,
                                             const b2d_gradient_stop* stops, uint32_t stop_count, const double* matrix,
                                             b2d_gradient** out);
B2DGPU_API b2dgpu_result b2d_gradient_destroy(b2d_gradient* g);

/* ---- BLPattern ------------------------------------------------------------------------------------------------ */
/* area = {x,y,w,h} or NULL for the whole image; matrix = 6 doubles or NULL.  The image must outlive the pattern. */
B2DGPU_API b2dgpu_result b2d_pattern_create(b2d_image* image, const int32_t* area, uint32_t extend_mode, const double* matrix, b2d_pattern** out);
B2DGPU_API b2dgpu_result b2d_pattern_destroy(b2d_pattern* p);

/* ---- BLContext ------------------------------------------------------------------------------------------------ */
/* Context create flag: only record commands (b2d_context_peek_batch); no runtime, target or rendering. */
#define B2D_CONTEXT_CREATE_FLAG_RECORD_ONLY 0x40000000u

typedef struct b2d_context_create_info {      /* BLContextCreateInfo (core/context.h:325-368) + device selection */
  uint32_t flags;
  uint32_t thread_count;                      /* ignored: the GPU grid replaces the worker pool                  */
  int32_t  pixel_origin_x, pixel_origin_y;
  int32_t  device;                            /* CUDA device ordinal                                             */
  uint32_t command_queue_limit;               /* commands per batch before an implicit flush; 0 = adaptive: 512,  */
                                              /* doubling up to 8192 within a frame (BLContextCreateInfo has the  */
                                              /* same field, core/context.h:338-346)                              */
  void*    runtime;                           /* optional shared b2dgpu_runtime*                                  */
  void*    stream;                            /* optional cudaStream_t for a runtime created by this context     */
  int32_t  slab_y0, slab_y1;                  /* band sharding: this context owns image rows [slab_y0, slab_y1) only;
                                                 0,0 = the whole image.  Mirrors the reference's worker/band model,
                                                 where every worker replays all commands clipped to its own bands
                                                 (raster/workerproc.cpp:260-299).                                   */
} b2d_context_create_info;

B2DGPU_API b2dgpu_result b2d_context_create(b2d_image* target, const b2d_context_create_info* info, b2d_context** out);
B2DGPU_API b2dgpu_result b2d_context_destroy(b2d_context* ctx);              /* implies end() */
B2DGPU_API b2dgpu_result b2d_context_end(b2d_context* ctx);                  /* flush + copy the canvas back to the image */
B2DGPU_API b2dgpu_result b2d_context_flush(b2d_context* ctx, uint32_t flags);/* BL_CONTEXT_FLUSH_SYNC = 0x80000000 */

B2DGPU_API b2dgpu_result b2d_context_set_comp_op(b2d_context* ctx, uint32_t comp_op);
B2DGPU_API b2dgpu_result b2d_context_set_global_alpha(b2d_context* ctx, double alpha);
B2DGPU_API b2dgpu_result b2d_context_set_fill_alpha(b2d_context* ctx, double alpha);
B2DGPU_API b2dgpu_result b2d_context_set_fill_rule(b2d_context* ctx, uint32_t fill_rule);
B2DGPU_API b2dgpu_result b2d_context_set_hint(b2d_context* ctx, uint32_t hint, uint32_t value);
B2DGPU_API b2dgpu_result b2d_context_set_flatten_tolerance(b2d_context* ctx, double tolerance);

B2DGPU_API b2dgpu_result b2d_context_set_fill_style_rgba32(b2d_context* ctx, uint32_t rgba32);
B2DGPU_API b2dgpu_result b2d_context_set_fill_style_gradient(b2d_context* ctx, const b2d_gradient* g);
B2DGPU_API b2dgpu_result b2d_context_set_fill_style_pattern(b2d_context* ctx, const b2d_pattern* p);

/* bl_context_apply_transform_op (op = BLTransformOp value, data as in core/matrix.h). */
B2DGPU_API b2dgpu_result b2d_context_apply_transform_op(b2d_context* ctx, uint32_t op, const double* data);

B2DGPU_API b2dgpu_result b2d_context_clear_all(b2d_context* ctx);
B2DGPU_API b2dgpu_result b2d_context_fill_all(b2d_context* ctx);
B2DGPU_API b2dgpu_result b2d_context_fill_rect_i(b2d_context* ctx, int32_t x, int32_t y, int32_t w, int32_t h);
/* bl_context_fill_mask_i (core/context.h): the fill style through an A8 mask image placed at (x, y); `mask_area` =
 * {x, y, w, h} inside the mask or NULL for all of it.  Like the reference (rastercontext.cpp:3594-3650) only pixel
 * aligned placements under a translation are implemented; anything else returns BL_ERROR_NOT_IMPLEMENTED. */
B2DGPU_API b2dgpu_result b2d_context_fill_mask_i(b2d_context* ctx, int32_t x, int32_t y, const b2d_image* mask, const int32_t* mask_area);
B2DGPU_API b2dgpu_result b2d_context_fill_rect_d(b2d_context* ctx, double x, double y, double w, double h);
/* bl_context_fill_path_d(origin, path): BLPathView given as command bytes + vertices (x,y pairs). */
B2DGPU_API b2dgpu_result b2d_context_fill_path_d(b2d_context* ctx, double ox, double oy, const uint8_t* cmd, const double* vtx, uint32_t count);
/* bl_context_fill_geometry(BL_GEOMETRY_TYPE_POLYGOND, ...). */
B2DGPU_API b2dgpu_result b2d_context_fill_polygon_d(b2d_context* ctx, const double* pts, uint32_t count);

/* Accessors used by bench.py / tests. */
B2DGPU_API b2dgpu_runtime* b2d_context_runtime(b2d_context* ctx);
B2DGPU_API b2dgpu_target* b2d_context_target(b2d_context* ctx);
/* The batch accumulated so far (valid until the next call on the context); lets callers upload it as a resident batch. */
B2DGPU_API b2dgpu_result b2d_context_peek_batch(b2d_context* ctx, b2dgpu_batch_view* out);
/* Drops the queued commands without rendering them. */
B2DGPU_API b2dgpu_result b2d_context_discard_batch(b2d_context* ctx);

/* Replays fills [first, first + count) of a flat scene description (include/b2d_scene.h) through the calls above. */
struct b2d_scene;
B2DGPU_API b2dgpu_result b2d_scene_replay(b2d_context* ctx, const struct b2d_scene* scene, uint32_t first, uint32_t count);

#ifdef __cplusplus
}
#endif
#endif /* B2D_HOST_H_INCLUDED */
