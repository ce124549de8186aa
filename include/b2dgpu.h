/*
 * b2dgpu.h - C-ABI of the B200-native Blend2D pipeline runtime (libb2dgpu.so).
 *
 * This header is the DROP-IN BOUNDARY.  It replaces, for the rendering hot path only,
 * what Blend2D's raster engine gets today from its CPU pipeline runtimes:
 *
 *   seam B  (pipeline provider)   blend2d/pipeline/piperuntime_p.h:39-62   PipeRuntime{_funcs.test,_funcs.get}
 *                                 blend2d/pipeline/pipedefs_p.h:229        FillFunc typedef
 *                                 blend2d/pipeline/pipedefs_p.h:235-368    Signature (32-bit key)
 *                                 blend2d/pipeline/pipedefs_p.h:838-1060   FetchData (176 bytes, 16-byte aligned)
 *   seam A  (batch consumer)      blend2d/raster/rastercontext.cpp:1021-1072  flush_render_batch()
 *                                 blend2d/raster/rendercommand_p.h:85-289     RenderCommand (64 bytes)
 *                                 blend2d/raster/workerproc.cpp:322-352       WorkerProc::process_work_data()
 *
 * Plain C, plain pointers and sizes, no C++/torch types.  All functions return a BLResult-compatible code
 * (uint32_t, 0 == success, errors >= 0x10000 - blend2d/core/api.h:1138-1152) and never throw.
 *
 * There is NO CPU fallback behind this interface: if no CUDA device (sm_100) is usable, b2dgpu_runtime_create()
 * fails with B2DGPU_ERROR_NOT_INITIALIZED and nothing renders.
 */
#ifndef B2DGPU_H_INCLUDED
#define B2DGPU_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#  define B2DGPU_API __declspec(dllexport)
#else
#  define B2DGPU_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------------------------------------------------
 * Result codes (values identical to BLResultCode, blend2d/core/api.h:1138-1208).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef uint32_t b2dgpu_result;

#define B2DGPU_SUCCESS                 0u
#define B2DGPU_ERROR_OUT_OF_MEMORY     0x00010000u
#define B2DGPU_ERROR_INVALID_VALUE     0x00010001u
#define B2DGPU_ERROR_INVALID_STATE     0x00010002u
#define B2DGPU_ERROR_INVALID_HANDLE    0x00010003u
#define B2DGPU_ERROR_NOT_INITIALIZED   0x00010006u
#define B2DGPU_ERROR_NOT_IMPLEMENTED   0x00010007u
#define B2DGPU_ERROR_NO_ENTRY          0x00010017u   /* PipeDynamicRuntime::test() miss, pipegenruntime.cpp:78-80 */
#define B2DGPU_ERROR_UNKNOWN           0x0001FFFFu   /* CUDA failure that has no BLResult equivalent            */

/* ------------------------------------------------------------------------------------------------------------------
 * Signature (blend2d/pipeline/pipedefs_p.h:244-251) - same bit layout, so the reference can pass its own value.
 * ---------------------------------------------------------------------------------------------------------------- */
#define B2DGPU_SIG_DST_FORMAT_MASK  0x0000000Fu
#define B2DGPU_SIG_SRC_FORMAT_MASK  0x000000F0u
#define B2DGPU_SIG_COMP_OP_MASK     0x00003F00u
#define B2DGPU_SIG_FILL_TYPE_MASK   0x0000C000u
#define B2DGPU_SIG_FETCH_TYPE_MASK  0x001F0000u
#define B2DGPU_SIG_PENDING_FLAG     0x80000000u

#define B2DGPU_SIG_DST_FORMAT(sig)  (((sig) >>  0) & 0xFu)
#define B2DGPU_SIG_SRC_FORMAT(sig)  (((sig) >>  4) & 0xFu)
#define B2DGPU_SIG_COMP_OP(sig)     (((sig) >>  8) & 0x3Fu)
#define B2DGPU_SIG_FILL_TYPE(sig)   (((sig) >> 14) & 0x3u)
#define B2DGPU_SIG_FETCH_TYPE(sig)  (((sig) >> 16) & 0x1Fu)
#define B2DGPU_MAKE_SIG(dst, src, op, fill, fetch) \
  (((uint32_t)(dst)) | ((uint32_t)(src) << 4) | ((uint32_t)(op) << 8) | ((uint32_t)(fill) << 14) | ((uint32_t)(fetch) << 16))

/* BLFormat / FormatExt (blend2d/core/format.h:33-39, core/format_p.h:20-44). */
enum {
  B2DGPU_FORMAT_NONE = 0, B2DGPU_FORMAT_PRGB32 = 1, B2DGPU_FORMAT_XRGB32 = 2, B2DGPU_FORMAT_A8 = 3,
  B2DGPU_FORMAT_FRGB32 = 4, B2DGPU_FORMAT_ZERO32 = 5
};

/* BLCompOp (blend2d/core/context.h:244-300) - the subset the GPU runtime implements. */
enum {
  B2DGPU_COMP_OP_SRC_OVER = 0, B2DGPU_COMP_OP_SRC_COPY = 1, B2DGPU_COMP_OP_SRC_IN = 2, B2DGPU_COMP_OP_SRC_OUT = 3,
  B2DGPU_COMP_OP_SRC_ATOP = 4, B2DGPU_COMP_OP_DST_OVER = 5, B2DGPU_COMP_OP_DST_COPY = 6, B2DGPU_COMP_OP_DST_IN = 7,
  B2DGPU_COMP_OP_DST_OUT = 8, B2DGPU_COMP_OP_DST_ATOP = 9, B2DGPU_COMP_OP_XOR = 10, B2DGPU_COMP_OP_CLEAR = 11,
  B2DGPU_COMP_OP_PLUS = 12, B2DGPU_COMP_OP_MINUS = 13, B2DGPU_COMP_OP_MODULATE = 14, B2DGPU_COMP_OP_MULTIPLY = 15,
  B2DGPU_COMP_OP_SCREEN = 16, B2DGPU_COMP_OP_OVERLAY = 17, B2DGPU_COMP_OP_DARKEN = 18, B2DGPU_COMP_OP_LIGHTEN = 19,
  B2DGPU_COMP_OP_COLOR_DODGE = 20, B2DGPU_COMP_OP_COLOR_BURN = 21, B2DGPU_COMP_OP_LINEAR_BURN = 22,
  B2DGPU_COMP_OP_LINEAR_LIGHT = 23, B2DGPU_COMP_OP_PIN_LIGHT = 24, B2DGPU_COMP_OP_HARD_LIGHT = 25,
  B2DGPU_COMP_OP_SOFT_LIGHT = 26, B2DGPU_COMP_OP_DIFFERENCE = 27, B2DGPU_COMP_OP_EXCLUSION = 28
};

/* FillType (pipedefs_p.h:64-76). */
enum { B2DGPU_FILL_NONE = 0, B2DGPU_FILL_BOX_A = 1, B2DGPU_FILL_MASK = 2, B2DGPU_FILL_ANALYTIC = 3 };

/* FetchType (pipedefs_p.h:123-184). */
enum {
  B2DGPU_FETCH_SOLID = 0,
  B2DGPU_FETCH_PATTERN_ALIGNED_BLIT = 1, B2DGPU_FETCH_PATTERN_ALIGNED_PAD = 2,
  B2DGPU_FETCH_PATTERN_ALIGNED_REPEAT = 3, B2DGPU_FETCH_PATTERN_ALIGNED_ROR = 4,
  B2DGPU_FETCH_PATTERN_FX_PAD = 5, B2DGPU_FETCH_PATTERN_FX_ROR = 6,
  B2DGPU_FETCH_PATTERN_FY_PAD = 7, B2DGPU_FETCH_PATTERN_FY_ROR = 8,
  B2DGPU_FETCH_PATTERN_FXFY_PAD = 9, B2DGPU_FETCH_PATTERN_FXFY_ROR = 10,
  B2DGPU_FETCH_PATTERN_AFFINE_NN_ANY = 11, B2DGPU_FETCH_PATTERN_AFFINE_NN_OPT = 12,
  B2DGPU_FETCH_PATTERN_AFFINE_BI_ANY = 13, B2DGPU_FETCH_PATTERN_AFFINE_BI_OPT = 14,
  B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD = 15, B2DGPU_FETCH_GRADIENT_LINEAR_NN_ROR = 16,
  B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_PAD = 17, B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_ROR = 18,
  B2DGPU_FETCH_GRADIENT_RADIAL_NN_PAD = 19, B2DGPU_FETCH_GRADIENT_RADIAL_NN_ROR = 20,
  B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_PAD = 21, B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_ROR = 22,
  B2DGPU_FETCH_GRADIENT_CONIC_NN = 23, B2DGPU_FETCH_GRADIENT_CONIC_DITHER = 24
};

/* FillRuleMask (pipedefs_p.h:112-115). */
#define B2DGPU_FILL_RULE_MASK_NON_ZERO 0xFFFFFFFFu
#define B2DGPU_FILL_RULE_MASK_EVEN_ODD 0x000001FFu

/* ------------------------------------------------------------------------------------------------------------------
 * FetchData - bit-for-bit the reference's `bl::Pipeline::FetchData` (pipedefs_p.h:838-1060; 176 bytes, align 16;
 * offsets checked against the reference headers with gcc 13 / x86-64 by tests/test_abi.py).  The reference's
 * host-side initialisers (pipeline/pipedefs.cpp:53-699) fill it; the GPU fetchers consume exactly these fields.
 * Pointers inside (`pixel_data`, `lut_data`) are HOST pointers at the boundary; b2dgpu_submit() uploads what they
 * reference and patches device addresses into its private copy.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef union b2dgpu_value64 { uint64_t u64; int64_t i64; double d; int32_t i32[2]; uint32_t u32[2]; } b2dgpu_value64;

typedef struct b2dgpu_fetch_solid { uint32_t prgb32; uint32_t reserved32; } b2dgpu_fetch_solid;

typedef struct b2dgpu_pattern_source {        /* FetchData::Pattern::SourceData, offset 0, 24 bytes (+8 pad) */
  const uint8_t* pixel_data;
  intptr_t stride;
  int32_t w, h;
} b2dgpu_pattern_source;

typedef struct b2dgpu_vert_extend {           /* FetchData::Pattern::VertExtendData, 48 bytes */
  intptr_t stride[2];
  uintptr_t y_stop[2];
  uintptr_t y_rewind_offset;
  intptr_t pixel_ptr_rewind_offset;
} b2dgpu_vert_extend;

typedef struct b2dgpu_pattern_simple {        /* FetchData::Pattern::Simple, 96 bytes, at offset 32 */
  int32_t tx, ty;
  int32_t rx, ry;
  uint8_t ix[16];                             /* ModuloTable (tables_p.h:25-27); unused by the GPU fetchers */
  uint32_t wa, wb, wc, wd;
  b2dgpu_vert_extend v_extend;
} b2dgpu_pattern_simple;

typedef struct b2dgpu_pattern_affine {        /* FetchData::Pattern::Affine, 144 bytes, at offset 32 */
  b2dgpu_value64 xx, xy;
  b2dgpu_value64 yx, yy;
  b2dgpu_value64 tx, ty;
  b2dgpu_value64 ox, oy;
  b2dgpu_value64 rx, ry;
  b2dgpu_value64 xx2, xy2;
  int32_t min_x, min_y;
  int32_t max_x, max_y;
  int32_t cor_x, cor_y;
  double tw, th;
  int32_t addr_mul32[2];
} b2dgpu_pattern_affine;

typedef struct b2dgpu_fetch_pattern {
  b2dgpu_pattern_source src;
  uint64_t _pad0;
  union { b2dgpu_pattern_simple simple; b2dgpu_pattern_affine affine; };
} b2dgpu_fetch_pattern;

typedef struct b2dgpu_gradient_lut { const void* data; uint32_t size; uint32_t _pad; } b2dgpu_gradient_lut;

typedef struct b2dgpu_gradient_linear {       /* FetchData::Gradient::Linear, 48 bytes, at offset 16 */
  b2dgpu_value64 pt[2];
  b2dgpu_value64 dy;
  b2dgpu_value64 dt;
  uint32_t maxi, rori;
  uint64_t _pad;
} b2dgpu_gradient_linear;

typedef struct b2dgpu_gradient_radial {       /* FetchData::Gradient::Radial, 112 bytes */
  double tx, ty;
  double yx, yy;
  double amul4, inv2a;
  double sq_fr, sq_inv2a;
  double b0, dd0;
  double by, ddy;
  float f32_ddd, f32_bd;
  uint32_t maxi, rori;
} b2dgpu_gradient_radial;

typedef struct b2dgpu_gradient_conic {        /* FetchData::Gradient::Conic, 80 bytes */
  double tx, ty;
  double yx, yy;
  float q_coeff[4];
  float n_div_1_2_4[3];
  float offset;
  float xx;
  uint32_t maxi, rori;
  uint32_t _pad;
} b2dgpu_gradient_conic;

typedef struct b2dgpu_fetch_gradient {
  b2dgpu_gradient_lut lut;
  union { b2dgpu_gradient_linear linear; b2dgpu_gradient_radial radial; b2dgpu_gradient_conic conic; };
} b2dgpu_fetch_gradient;

typedef union b2dgpu_fetch_data {
  b2dgpu_fetch_solid solid;
  b2dgpu_fetch_pattern pattern;
  b2dgpu_fetch_gradient gradient;
  uint8_t bytes[176];
#ifdef __cplusplus
} __attribute__((aligned(16))) b2dgpu_fetch_data;
#else
} __attribute__((aligned(16))) b2dgpu_fetch_data;
#endif

/* ------------------------------------------------------------------------------------------------------------------
 * Seam B - pipeline runtime.
 *
 * The first 32 bytes of the object b2dgpu_runtime_create() returns have the layout of the reference's
 * `bl::Pipeline::PipeRuntime` (piperuntime_p.h:39-62): {u8 type, u8 flags, u16 size, destroy(), test(), get()} so a
 * reference build can hand the pointer to `PipeProvider::init()` (rastercontext.cpp:4345) unchanged.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct b2dgpu_runtime b2dgpu_runtime;

typedef void (*b2dgpu_fill_func)(void* ctx_data, const void* fill_data, const void* fetch_data);

typedef struct b2dgpu_dispatch_data {         /* DispatchData, pipedefs_p.h:370-388 */
  b2dgpu_fill_func fill_func;                 /* non-null token; calling it is a programming error (aborts)    */
  b2dgpu_fill_func fetch_func;                /* not callable either: B2DGPU_DISPATCH_TAG | signature.  The     */
} b2dgpu_dispatch_data;                       /* frontend stores DispatchData over the command's signature      */
                                              /* (rendercommand_p.h:169-174); the batch consumer reads it back  */
#define B2DGPU_DISPATCH_TAG 0xB2D6000000000000ull   /* non-canonical x86-64 address: a call faults               */
#define B2DGPU_DISPATCH_SIGNATURE(dd) ((uint32_t)(uintptr_t)(dd)->fetch_func)
#define B2DGPU_DISPATCH_IS_GPU(dd) ((((uintptr_t)(dd)->fetch_func) >> 48) == (B2DGPU_DISPATCH_TAG >> 48))

typedef struct b2dgpu_create_info {
  uint32_t struct_size;                       /* sizeof(b2dgpu_create_info)                                     */
  int32_t  device;                            /* CUDA device ordinal (one process per GPU: LOCAL_RANK)          */
  void*    stream;                            /* optional cudaStream_t to run on (0 = runtime creates its own)  */
  uint32_t flags;                             /* reserved, 0                                                    */
} b2dgpu_create_info;

B2DGPU_API b2dgpu_result b2dgpu_runtime_create(const b2dgpu_create_info* info, b2dgpu_runtime** out);
B2DGPU_API b2dgpu_result b2dgpu_runtime_destroy(b2dgpu_runtime* rt);

/* PipeRuntime::_funcs.test / .get (piperuntime_p.h:53-56).  `cache` is the reference's PipeLookupCache* or NULL; it
 * is treated as opaque here (the shim stores into it on the reference side, fixedpiperuntime.cpp:319-320). */
B2DGPU_API b2dgpu_result b2dgpu_runtime_test(b2dgpu_runtime* rt, uint32_t signature, b2dgpu_dispatch_data* out, void* cache);
B2DGPU_API b2dgpu_result b2dgpu_runtime_get(b2dgpu_runtime* rt, uint32_t signature, b2dgpu_dispatch_data* out, void* cache);

/* ------------------------------------------------------------------------------------------------------------------
 * Render target: the device-resident canvas that stands for `ContextData::dst` (pipedefs_p.h:513-518).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct b2dgpu_target b2dgpu_target;

typedef struct b2dgpu_image_data {            /* BLImageData (core/image.h): pixel_data, stride, size, format */
  void*    pixel_data;
  intptr_t stride;
  int32_t  w, h;
  uint32_t format;
  uint32_t flags;
} b2dgpu_image_data;

/* Creates a canvas of w x h pixels in HBM owned by the runtime (rows padded to the tile width, 512-byte aligned). */
B2DGPU_API b2dgpu_result b2dgpu_target_create(b2dgpu_runtime* rt, int32_t w, int32_t h, uint32_t format, b2dgpu_target** out);
/* Band-sharded canvas: this GPU owns rows [y0, y1) of a `full_h`-row image. */
B2DGPU_API b2dgpu_result b2dgpu_target_create_slab(b2dgpu_runtime* rt, int32_t w, int32_t full_h, int32_t y0, int32_t y1, uint32_t format, b2dgpu_target** out);
B2DGPU_API b2dgpu_result b2dgpu_target_destroy(b2dgpu_target* t);
/* Host <-> device canvas copies (pinned staging inside; asynchronous on the runtime stream). */
B2DGPU_API b2dgpu_result b2dgpu_target_upload(b2dgpu_target* t, const b2dgpu_image_data* src);
B2DGPU_API b2dgpu_result b2dgpu_target_download(b2dgpu_target* t, const b2dgpu_image_data* dst);
B2DGPU_API b2dgpu_result b2dgpu_target_clear(b2dgpu_target* t);
/* Optional: page-lock the host pixels of an image that will be uploaded / downloaded repeatedly, so the copies are
 * direct DMA instead of going through the runtime's staging buffer.  The caller keeps the memory alive until
 * b2dgpu_host_unregister().  (BLImage pixel memory is plain malloc memory: core/image.cpp:39-41.) */
B2DGPU_API b2dgpu_result b2dgpu_host_register(b2dgpu_runtime* rt, void* pixels, size_t bytes);
B2DGPU_API b2dgpu_result b2dgpu_host_unregister(b2dgpu_runtime* rt, void* pixels);
/* Device view of the canvas (for torch / NCCL plumbing): base pointer of row y0, stride in bytes, padded extents. */
B2DGPU_API b2dgpu_result b2dgpu_target_device_view(b2dgpu_target* t, void** dev_ptr, intptr_t* stride, int32_t* padded_w, int32_t* padded_h);
/* Makes `stream` (a cudaStream_t) wait for the last render into `t` - and for nothing queued behind it: with
 * b2dgpu_batch_render_multi() a consumer (the NCCL send of a stripe) can start on stripe j while stripes
 * j + 1 .. are still being composited.  Returns immediately; no host synchronisation. */
B2DGPU_API b2dgpu_result b2dgpu_target_wait(b2dgpu_target* t, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Seam A - render batch.
 *
 * One `b2dgpu_command` per RenderCommand (rendercommand_p.h:85-289).  Commands are applied in array order, exactly
 * like the per-band replay in workerproc.cpp:200-252.
 * ---------------------------------------------------------------------------------------------------------------- */
enum {
  B2DGPU_CMD_FILL_BOX_A    = 1,               /* RenderCommandType::kFillBoxA   - box in pixels                  */
  B2DGPU_CMD_FILL_BOX_U    = 2,               /* RenderCommandType::kFillBoxU   - box in 24.8 fixed point        */
  B2DGPU_CMD_FILL_ANALYTIC = 3,               /* RenderCommandType::kFillAnalytic with CPU-built edges           */
  B2DGPU_CMD_FILL_GEOMETRY = 4,               /* kFillAnalytic whose edges come from a RenderJob_GeometryOp:     */
                                              /* the GPU edge builder flattens/clips the path itself             */
  B2DGPU_CMD_FILL_BOX_MASK_A = 5              /* RenderCommandType::kFillBoxMaskA (rendercommand_p.h:44,107):    */
};                                            /* box in pixels, per-pixel A8 mask; reserved[0] = index of a      */
                                              /* fetch_data entry whose pattern.src describes the mask rows that */
                                              /* start at the box's top-left pixel                                */

typedef struct b2dgpu_command {               /* 64 bytes */
  uint32_t type;                              /* B2DGPU_CMD_*                                                    */
  uint32_t signature;                         /* reference Signature value (dst|src|comp_op|fill|fetch)          */
  uint32_t alpha;                             /* 0..255 (FillData::*::alpha, pipedefs_p.h:541-605)                */
  uint32_t fill_rule_mask;                    /* FILL_ANALYTIC / FILL_GEOMETRY                                   */
  int32_t  box[4];                            /* BOX_A: pixels; BOX_U: 24.8 fixed; others: ignored               */
  uint32_t solid_prgb32;                      /* FETCH_SOLID: premultiplied colour (FetchData::Solid)            */
  uint32_t fetch_index;                       /* otherwise: index into batch->fetch_data                         */
  uint32_t data_offset;                       /* ANALYTIC: first edge in batch->edges; GEOMETRY: first segment   */
  uint32_t data_count;                        /* ANALYTIC: #edges; GEOMETRY: #segments                            */
  uint32_t state_index;                       /* GEOMETRY: index into batch->geometry_states                     */
  uint32_t reserved[3];
} b2dgpu_command;

/* A flattened, already clipped edge in 24.8 fixed point: what EdgeBuilder emits (edgestorage_p.h:38-60), stored as
 * one record per line in its ORIGINAL direction; y0 > y1 means the reference's sign bit is set. */
typedef struct b2dgpu_edge { int32_t x0, y0, x1, y1; } b2dgpu_edge;

/* Path segment kinds for B2DGPU_CMD_FILL_GEOMETRY (BLPathCmd, core/path.h): each segment names its start vertex and
 * its first following vertex; quads/cubics/conics read their remaining control points at p1+1.. . */
enum { B2DGPU_SEG_LINE = 0, B2DGPU_SEG_QUAD = 1, B2DGPU_SEG_CUBIC = 2, B2DGPU_SEG_CONIC = 3 };

typedef struct b2dgpu_segment {               /* 12 bytes */
  uint32_t p0;                                /* vertex index of the segment start (absolute, in batch->vertices) */
  uint32_t p1_kind;                           /* (vertex index of next point << 2) | B2DGPU_SEG_*                 */
  uint32_t command;                           /* index of the owning command                                      */
} b2dgpu_segment;

/* SharedFillState + final transform (raster/statedata_p.h:145-149, rastercontext.cpp:2339-2356): what a
 * RenderJob_GeometryOp needs to build edges. */
typedef struct b2dgpu_geometry_state {        /* 96 bytes */
  double m[6];                                /* final transform, already scaled by 256 (final_transform_fixed)  */
  double clip[4];                             /* final clip box in 24.8 units as doubles (x0,y0,x1,y1)           */
  double tolerance_sq;                        /* (flatten tolerance * 256)^2, rastercontext.cpp:208-211          */
  uint32_t transform_type;                    /* <= 2 (identity/translate/scale) => scale path, else affine      */
  uint32_t reserved;
} b2dgpu_geometry_state;

/* Glyph instancing.  The caller keeps an append-only cache of TrueType outlines (format:
 * blend2d_b200/csrc/dev_glyph.cuh; the runtime mirrors it on the device and uploads only what was appended since the last
 * submit) and describes every glyph of a text fill by ONE of these records instead of its decoded outline.  The device
 * replays the reference's decoder (opentype/otglyf.cpp:376-448) for the instance matrix and writes the vertices and path
 * segments the glyph contributes at `vertex_base` / `segment_base` of the batch's arrays, behind the uploaded ones. */
typedef struct b2dgpu_glyph_instance {
  double m[6];                 /* glyph matrix: font matrix x placement x final (fixed-point) transform, core/font.cpp:659-744 */
  uint32_t blob_offset;        /* word offset of the glyph inside the cache                                        */
  uint32_t vertex_base;        /* >= vertex_count: first vertex index this instance writes                          */
  uint32_t segment_base;       /* >= segment_count: first segment index this instance writes                        */
  uint32_t command;            /* the FILL_GEOMETRY command the segments belong to                                  */
} b2dgpu_glyph_instance;

/* Gradient table interpolation on the device.  A gradient FetchData whose `lut.data` is NULL (and whose
 * `lut.size` is set) gets its table built by the runtime from the gradient's stops - core/gradient.cpp:289-330 calling
 * interpolate_prgb32 (pixelops/interpolation_avx2.cpp:18-205) - instead of receiving it from the host: 16 bytes per stop
 * travel instead of 1-4 KB per gradient.  Nearest-neighbour (32-bit) tables only. */
typedef struct b2dgpu_gradient_stop { double offset; uint64_t rgba64; } b2dgpu_gradient_stop;   /* = BLGradientStop */
typedef struct b2dgpu_lut_request {
  uint32_t fetch_index;        /* the FetchData entry the table belongs to                                         */
  uint32_t stop_offset;        /* first stop in b2dgpu_batch_view::lut_stops                                       */
  uint32_t stop_count;         /* >= 1, offsets ascending in [0, 1]                                                */
  uint32_t lut_size;           /* entries; must equal fetch_data[fetch_index].gradient.lut.size                    */
} b2dgpu_lut_request;

typedef struct b2dgpu_batch_view {
  uint32_t struct_size;
  uint32_t command_count;
  const b2dgpu_command* commands;
  const b2dgpu_fetch_data* fetch_data;   uint32_t fetch_count;   uint32_t _pad0;
  const b2dgpu_edge* edges;              uint32_t edge_count;    uint32_t _pad1;
  const double* vertices;                uint32_t vertex_count;  uint32_t _pad2;   /* x,y pairs */
  const b2dgpu_segment* segments;        uint32_t segment_count; uint32_t _pad3;
  const b2dgpu_geometry_state* geometry_states; uint32_t geometry_state_count; uint32_t _pad4;
  int32_t pixel_origin_x, pixel_origin_y;                                          /* ContextData::pixel_origin */
  /* --- glyph instancing; absent (struct_size ends before it) or zero when unused --- */
  const uint32_t* glyph_cache;           uint32_t glyph_cache_words; uint32_t _pad5;
  uint64_t glyph_cache_id;               /* changes when the cache is not an extension of the previous submit's */
  const b2dgpu_glyph_instance* glyph_instances; uint32_t glyph_instance_count; uint32_t _pad6;
  uint32_t generated_vertex_count;       /* vertices / segments the instances write: commands may refer to segment */
  uint32_t generated_segment_count;      /* indices up to segment_count + generated_segment_count                 */
  /* --- device-built gradient tables; absent or zero when unused --- */
  const b2dgpu_lut_request* lut_requests;  uint32_t lut_request_count;  uint32_t _pad7;
  const b2dgpu_gradient_stop* lut_stops;   uint32_t lut_stop_count;     uint32_t _pad8;
} b2dgpu_batch_view;
#define B2DGPU_BATCH_VIEW_SIZE_V1 ((uint32_t)offsetof(b2dgpu_batch_view, glyph_cache))

/* Copies/serialises the batch (the caller may free it on return, cf. rastercontext.cpp:1060-1063), uploads it and
 * launches the edge builder + tile compositor asynchronously on the runtime stream. */
B2DGPU_API b2dgpu_result b2dgpu_submit(b2dgpu_runtime* rt, b2dgpu_target* target, const b2dgpu_batch_view* batch);

/* Device-resident batches (inputs already in HBM): upload once, replay many times. */
typedef struct b2dgpu_batch b2dgpu_batch;
B2DGPU_API b2dgpu_result b2dgpu_batch_upload(b2dgpu_runtime* rt, const b2dgpu_batch_view* view, b2dgpu_batch** out);
B2DGPU_API b2dgpu_result b2dgpu_batch_destroy(b2dgpu_batch* b);
B2DGPU_API b2dgpu_result b2dgpu_batch_render(b2dgpu_runtime* rt, b2dgpu_target* target, b2dgpu_batch* batch);
/* Same batch into several slab targets of one canvas (interleaved stripes of a band-sharded image): the geometry pass
 * runs once, clipping + compositing once per target. */
B2DGPU_API b2dgpu_result b2dgpu_batch_render_multi(b2dgpu_runtime* rt, b2dgpu_target* const* targets, uint32_t target_count, b2dgpu_batch* batch);

/* BL_CONTEXT_FLUSH_SYNC (core/context.h:105): blocks until everything submitted so far has executed. */
B2DGPU_API b2dgpu_result b2dgpu_sync(b2dgpu_runtime* rt);

/* Counters accumulated since the last reset - the numbers bench.py reports. */
typedef struct b2dgpu_stats {
  uint64_t kernel_launches;                   /* number of OUR kernels launched                                  */
  uint64_t pixels_composited;                 /* pixels whose mask was non-zero (written), summed over commands  */
  uint64_t commands;                          /* commands rendered                                               */
  uint64_t edges;                             /* edges produced by the edge builder + supplied                   */
  uint64_t h2d_bytes, d2h_bytes;
  /* Device time of OUR kernels, measured with CUDA events on the runtime stream while profiling is enabled
   * (b2dgpu_set_profiling): the tile compositor (K2+K3) and everything before it (K1 edge builder, scan, finalize). */
  double tile_kernel_ms, build_kernels_ms;
  uint64_t tile_kernel_launches;
} b2dgpu_stats;
B2DGPU_API b2dgpu_result b2dgpu_get_stats(b2dgpu_runtime* rt, b2dgpu_stats* out, int reset);
/* Enables (1) / disables (0) per-kernel CUDA-event timing; adds two event records per phase and render. */
B2DGPU_API b2dgpu_result b2dgpu_set_profiling(b2dgpu_runtime* rt, int enabled);
/* b2dgpu_stats::pixels_composited is maintained by the compositing kernels themselves (a vote and a few adds per
 * composited group of pixels).  On by default; a caller that does not read the statistic can switch it off. */
B2DGPU_API b2dgpu_result b2dgpu_set_pixel_counting(b2dgpu_runtime* rt, int enabled);

/* ------------------------------------------------------------------------------------------------------------------
 * Process-wide access.  A Blend2D application that selects this runtime through BLContextCreateInfo (shim/)
 * never holds a b2dgpu_runtime*: its contexts own them.  These entry points address every runtime of the process.
 * ---------------------------------------------------------------------------------------------------------------- */
/* Sum of b2dgpu_get_stats() over all runtimes created so far (destroyed ones included). */
B2DGPU_API b2dgpu_result b2dgpu_global_stats(b2dgpu_stats* out, int reset);
/* b2dgpu_set_profiling() for every live runtime and for the ones created later. */
B2DGPU_API b2dgpu_result b2dgpu_global_set_profiling(int enabled);
/* b2dgpu_set_pixel_counting() for every live runtime and for the ones created later. */
B2DGPU_API b2dgpu_result b2dgpu_global_set_pixel_counting(int enabled);
/* Capture: between begin and end every b2dgpu_submit() of the process also keeps a device-resident copy of its batch
 * (b2dgpu_batch_upload).  b2dgpu_capture_replay() renders the captured batches again, `times` times, into the targets
 * they were submitted to - inputs already in HBM, nothing crosses PCIe - and returns the device time of the replay
 * (CUDA events on the runtime's stream; synchronous).  The contexts that were captured must still be alive. */
typedef struct b2dgpu_capture b2dgpu_capture;
B2DGPU_API b2dgpu_result b2dgpu_capture_begin(void);
B2DGPU_API b2dgpu_result b2dgpu_capture_end(b2dgpu_capture** out);
B2DGPU_API b2dgpu_result b2dgpu_capture_info(const b2dgpu_capture* c, uint32_t* batches, uint64_t* commands);
B2DGPU_API b2dgpu_result b2dgpu_capture_replay(b2dgpu_capture* c, uint32_t times, float* ms_out);
B2DGPU_API b2dgpu_result b2dgpu_capture_destroy(b2dgpu_capture* c);

/* Debug/KAT access used by the parity tests: runs only the edge builder and returns the flattened edges. */
B2DGPU_API b2dgpu_result b2dgpu_debug_build_edges(b2dgpu_runtime* rt, const b2dgpu_batch_view* batch,
                                                  b2dgpu_edge* edges_out, uint32_t capacity, uint32_t* count_out,
                                                  uint32_t* per_command_begin_out /* command_count + 1 entries */);

B2DGPU_API const char* b2dgpu_last_error_message(void);
B2DGPU_API uint32_t b2dgpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B2DGPU_H_INCLUDED */
