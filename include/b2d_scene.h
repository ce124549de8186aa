/*
 * b2d_scene.h - flat description of a synthetic vector scene (a list of fills) that can be replayed natively through
 * either front end: blend2d_b200's host API (b2d_scene_replay, in libb2dgpu.so) or the unmodified reference's public C
 * API (shim/bl_scene_driver.cpp, built against the GPU-enabled Blend2D and against the unmodified reference).  bench.py generates the arrays with numpy; replaying them in C
 * keeps Python call overhead out of every timed region on BOTH arms, the way bl_bench drives Blend2D
 * (blend2d-testing/bench/bl_bench_backend_blend2d.cpp).
 */
#ifndef B2D_SCENE_H_INCLUDED
#define B2D_SCENE_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { B2D_SCENE_GEOM_RECT_I = 0, B2D_SCENE_GEOM_RECT_D = 1, B2D_SCENE_GEOM_POLYGON = 2, B2D_SCENE_GEOM_PATH = 3,
       B2D_SCENE_GEOM_TEXT = 4 /* UTF-8 text at (rect[0], rect[1]): bytes [vtx_offset, vtx_offset + vtx_count) of scene->text */ };
enum { B2D_SCENE_STYLE_SOLID = 0, B2D_SCENE_STYLE_LINEAR = 1, B2D_SCENE_STYLE_RADIAL = 2, B2D_SCENE_STYLE_CONIC = 3, B2D_SCENE_STYLE_PATTERN = 4 };

typedef struct b2d_scene_stop { double offset; uint64_t rgba64; } b2d_scene_stop;

typedef struct b2d_scene_fill {       /* 160 bytes */
  uint32_t geom;                      /* B2D_SCENE_GEOM_*                                                  */
  uint32_t vtx_offset, vtx_count;     /* POLYGON / PATH: range in scene->vertices (and scene->path_cmds)   */
  uint32_t fill_rule;                 /* BLFillRule                                                        */
  uint32_t comp_op;                   /* BLCompOp                                                          */
  uint32_t style;                     /* B2D_SCENE_STYLE_*                                                 */
  uint32_t extend;                    /* BLExtendMode of the gradient / pattern                            */
  uint32_t stop_offset, stop_count;   /* gradients: range in scene->stops                                  */
  uint32_t rgba32;                    /* solid colour                                                      */
  uint32_t quality;                   /* gradient or pattern quality hint                                  */
  uint32_t has_transform;             /* user transform = rotate(angle) about (cx, cy) for this fill only  */
  double rect[4];                     /* RECT_I / RECT_D: x, y, w, h                                       */
  double values[6];                   /* gradient values, or pattern matrix                                */
  double angle, cx, cy;               /* see has_transform                                                 */
  double stroke_width;                /* > 0: the geometry is stroked with this width instead of filled    */
} b2d_scene_fill;

typedef struct b2d_scene {
  const b2d_scene_fill* fills;  uint32_t fill_count;  uint32_t _pad0;
  const double* vertices;       uint32_t vertex_count; uint32_t _pad1;   /* x,y pairs */
  const uint8_t* path_cmds;                                                 /* one BLPathCmd per vertex (PATH only) */
  const b2d_scene_stop* stops;  uint32_t stop_count;   uint32_t _pad2;
  /* optional pattern texture (PRGB32, tightly packed) */
  const uint32_t* texture;      int32_t texture_w, texture_h;
  /* optional text (GEOM_TEXT): UTF-8 bytes, a font file and its size in pixels */
  const char* text;             uint32_t text_size;    uint32_t _pad3;
  const char* font_file;
  double font_size;
} b2d_scene;

#ifdef __cplusplus
}
#endif
#endif
