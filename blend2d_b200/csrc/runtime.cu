// runtime.cu - implementation of the C-ABI declared in include/b2dgpu.h.
//
// Host-side plumbing only: device/stream ownership, the device-resident canvas, batch serialisation (one pinned staging
// block -> one H2D copy -> pointer patching), and the kernel sequence  K1 count -> scan -> K1 write -> finalize ->
// K2+K3 tile render.  No pixel is ever computed on the host: if CUDA is unavailable every entry point fails.
#include "kernels.h"
#include "dev_glyph.cuh"
#include "dev_common.cuh"

#include <cuda_runtime.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

using namespace b2d;

// ---------------------------------------------------------------------------------------------------------------
// Errors
// ---------------------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static b2dgpu_result fail(b2dgpu_result code, const char* what, const char* detail = nullptr) {
  g_last_error = what;
  if (detail) { g_last_error += ": "; g_last_error += detail; }
  return code;
}

static b2dgpu_result cuda_fail(cudaError_t e, const char* what) {
  b2dgpu_result code = (e == cudaErrorMemoryAllocation) ? B2DGPU_ERROR_OUT_OF_MEMORY : B2DGPU_ERROR_UNKNOWN;
  return fail(code, what, cudaGetErrorString(e));
}

#define CU_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return cuda_fail(e__, #expr); } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Small helpers
// ---------------------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DevBuffer {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

struct PinnedBuffer {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaEvent_t free_event = nullptr;       // recorded after the last async copy that reads this buffer
  bool in_flight = false;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr; cap = 0;
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMallocHost(&ptr, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; cap = 0; if (free_event) cudaEventDestroy(free_event); free_event = nullptr; }
};

// ---------------------------------------------------------------------------------------------------------------
// Objects
// ---------------------------------------------------------------------------------------------------------------
struct b2dgpu_runtime {
  // --- layout of bl::Pipeline::PipeRuntime (pipeline/piperuntime_p.h:39-62) ---
  uint8_t runtime_type;                     // PipeRuntimeType: 0 static, 1 JIT; 2 = GPU (new)
  uint8_t runtime_flags;                    // PipeRuntimeFlags::kIsolated = 1: destroyed on detach
  uint16_t runtime_size;
  uint32_t _pad;
  void (*destroy_fn)(b2dgpu_runtime*);
  b2dgpu_result (*test_fn)(b2dgpu_runtime*, uint32_t, b2dgpu_dispatch_data*, void*);
  b2dgpu_result (*get_fn)(b2dgpu_runtime*, uint32_t, b2dgpu_dispatch_data*, void*);
  // --- private ---
  uint32_t magic;
  int device;
  cudaStream_t stream;
  bool own_stream;
  std::mutex mutex;

  uint8_t* d_bayer;
  unsigned long long* d_pixel_counter;
  uint32_t* d_scalars;                      // [0] = total built edges, [1] = error flag
  uint32_t* h_scalars;                      // pinned mirror

  PinnedBuffer staging[2];
  int staging_next;
  // b2dgpu_submit() is pipelined: batch k + 1 is uploaded, counted and scanned on `prep_stream` (its edge total is
  // read back there) while batch k is still rendering on `stream`.  Two device blocks alternate.
  cudaStream_t prep_stream;
  DevBuffer oneshot_block[2];
  cudaEvent_t slot_done[2];                 // recorded on `stream` after the render that reads oneshot_block[i]
  bool slot_busy[2];
  int slot_next;
  cudaEvent_t prep_ready;
  DevBuffer oneshot_edges;
  DevBuffer glyph_cache;                    // device mirror of the caller's glyph cache (b2dgpu_batch_view::glyph_cache)
  uint32_t glyph_cache_words;               // words already uploaded
  uint64_t glyph_cache_id;
  cudaEvent_t glyph_cache_ready;            // recorded behind the last upload into the mirror
  DevBuffer bins;                           // per-band command lists (k_bin_*), rebuilt by every render
  uint32_t bin_capacity;                    // cells the lists can hold; grown when a render reports that it needed more
  uint32_t* h_bin_state;                    // pinned: [0] cells the last render needed, [1] whether its lists were built,
                                            // [2] whether its per-cell edge lists were built, [3] (edge, band) pairs it needed
  size_t edge_pair_capacity;                // grown when a render reports that its edge lists did not fit
  PinnedBuffer image_staging;

  b2dgpu_stats stats;
  bool profiling;
  bool count_pixels;                        // b2dgpu_stats::pixels_composited is maintained (a few instructions per composited group)
  int sm_count;
  std::vector<cudaEvent_t> prof_events;     // triples: start, after build kernels, after tile kernel
};

struct b2dgpu_target {
  b2dgpu_runtime* rt;
  int w, h;                                 // h = rows held here (slab height)
  int full_h, y0;
  uint32_t format;
  int bpp;
  uint8_t* d_pixels;
  size_t stride;
  int padded_w, padded_h;
  cudaEvent_t rendered;                     // recorded after the last compositing kernel that wrote this target
  bool rendered_valid;
};

// Offsets of the sections inside a device block.
struct BlockLayout {
  size_t commands, fetch_data, vertices, segments, states, supplied_edges, blobs, instances;
  size_t lut_requests, lut_stops, lut_offsets;      // uploaded: the requests, their stops, word offset of every table
  size_t gen_luts;                                  // device only: the tables k_build_luts writes
  size_t seg_counts, seg_offsets, scan_scratch, bbox_fixed, bbox_px, cmd_edges;
  // Host-filled data: [0, upload1_bytes) of the block - everything up to and including the uploaded vertices - and the
  // uploaded segments; the vertices / segments the glyph instances generate follow their uploaded ones on the device
  // only.  The pinned staging buffer holds part 1 at offset 0 and the segments at `staging_segments`.
  size_t upload1_bytes, seg_upload_bytes, staging_segments;
  size_t upload_bytes;                      // bytes copied host -> device per upload (statistics, staging size)
  size_t total_bytes;
};

// A batch that is a single solid box fill (fill_all, clear_all, one big FillRectA).
// `one`: the batch is a single FillBoxA with SrcOver / SrcCopy, any source (k_stream_one); `ok`: ... with a solid source.
struct SolidFill { bool ok, one; int box[4]; uint32_t comp_op, alpha, prgb32, fetch_type, src_format, fetch_index, lut_entries; };

struct b2dgpu_batch {
  b2dgpu_runtime* rt;
  DevBuffer block;
  DevBuffer edges;
  BlockLayout lay;
  uint32_t command_count, fetch_count, supplied_edges, vertex_count, segment_count, state_count;   // segment_count: uploaded + generated
  uint32_t instance_count;
  int origin_x, origin_y;
  bool has_analytic;
  uint32_t built_edges;                     // known after the first render
  bool built_known;
  bool edges_staged;                        // supplied edges already copied into `edges`
  bool stream_ok;
  int stream_box[4];
  SolidFill solid;
};

static const uint32_t kRuntimeMagic = 0xB2D09B00u;

// ---------------------------------------------------------------------------------------------------------------
// Process-wide registry: an application that reaches the runtime only through Blend2D (shim/) never sees the
// b2dgpu_runtime objects its contexts own; b2dgpu_global_* and b2dgpu_capture_* address all of them.
// ---------------------------------------------------------------------------------------------------------------
struct CaptureEntry { b2dgpu_runtime* rt; b2dgpu_target* target; b2dgpu_batch* batch; };
struct b2dgpu_capture { std::vector<CaptureEntry> entries; };

static std::mutex g_registry_mutex;
static std::vector<b2dgpu_runtime*> g_runtimes;
static std::vector<b2dgpu_target*> g_targets;
static b2dgpu_stats g_retired_stats;                 // totals of destroyed runtimes
static bool g_profiling_default = false;
static bool g_count_pixels_default = true;
static b2dgpu_capture* g_capture = nullptr;          // non-null while b2dgpu_capture_begin() .. _end()

template<typename T>
static void registry_remove(std::vector<T*>& v, T* p) {
  for (size_t i = 0; i < v.size(); i++) if (v[i] == p) { v[i] = v.back(); v.pop_back(); return; }
}
template<typename T>
static bool registry_has(const std::vector<T*>& v, T* p) {
  for (T* q : v) if (q == p) return true;
  return false;
}

static void stats_add(b2dgpu_stats& a, const b2dgpu_stats& b) {
  a.kernel_launches += b.kernel_launches; a.pixels_composited += b.pixels_composited; a.commands += b.commands;
  a.edges += b.edges; a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes;
  a.tile_kernel_ms += b.tile_kernel_ms; a.build_kernels_ms += b.build_kernels_ms; a.tile_kernel_launches += b.tile_kernel_launches;
}

// ---------------------------------------------------------------------------------------------------------------
// Seam B
// ---------------------------------------------------------------------------------------------------------------
static void fill_func_token(void*, const void*, const void*) {
  fprintf(stderr, "b2dgpu: a GPU FillFunc token was called on the CPU; GPU pipelines only run through b2dgpu_submit()\n");
  abort();
}

static bool signature_supported(uint32_t sig) {
  uint32_t dst = B2DGPU_SIG_DST_FORMAT(sig), src = B2DGPU_SIG_SRC_FORMAT(sig);
  uint32_t op = B2DGPU_SIG_COMP_OP(sig), fill = B2DGPU_SIG_FILL_TYPE(sig), fetch = B2DGPU_SIG_FETCH_TYPE(sig);
  if (sig & B2DGPU_SIG_PENDING_FLAG) return false;
  if (!(dst == B2DGPU_FORMAT_PRGB32 || dst == B2DGPU_FORMAT_XRGB32 || dst == B2DGPU_FORMAT_A8 ||
        dst == B2DGPU_FORMAT_FRGB32 || dst == B2DGPU_FORMAT_ZERO32)) return false;
  const bool base_op = op == B2DGPU_COMP_OP_SRC_OVER || op == B2DGPU_COMP_OP_SRC_COPY || op == B2DGPU_COMP_OP_PLUS ||
                       op == B2DGPU_COMP_OP_MULTIPLY || op == B2DGPU_COMP_OP_SCREEN;
  // The rest of BLCompOp (dev_pixel.cuh comp_jit_ext / comp_jit_light): only on a PRGB32 destination - the JIT has
  // separate code for A8 and for destinations without alpha.  Clear and DstCopy never arrive (the frontend simplifies
  // them, core/compopsimplifyimpl_p.h).
  const bool ext_op = dst == B2DGPU_FORMAT_PRGB32 && op >= B2DGPU_COMP_OP_SRC_IN && op <= B2DGPU_COMP_OP_EXCLUSION &&
                      op != B2DGPU_COMP_OP_DST_COPY && op != B2DGPU_COMP_OP_CLEAR;
  if (!base_op && !ext_op) return false;
  if (fill < B2DGPU_FILL_BOX_A || fill > B2DGPU_FILL_ANALYTIC) return false;
  if (fetch > B2DGPU_FETCH_GRADIENT_CONIC_DITHER) return false;
  if (fetch >= B2DGPU_FETCH_PATTERN_ALIGNED_BLIT && fetch <= B2DGPU_FETCH_PATTERN_AFFINE_BI_OPT) {
    if (!(src == B2DGPU_FORMAT_PRGB32 || src == B2DGPU_FORMAT_XRGB32 || src == B2DGPU_FORMAT_A8 ||
          src == B2DGPU_FORMAT_FRGB32 || src == B2DGPU_FORMAT_ZERO32)) return false;
  }
  return true;
}

static b2dgpu_result runtime_lookup(b2dgpu_runtime* rt, uint32_t signature, b2dgpu_dispatch_data* out, bool is_test) {
  if (!rt || rt->magic != kRuntimeMagic || !out) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_runtime_get: invalid argument");
  if (!signature_supported(signature))
    return fail(is_test ? B2DGPU_ERROR_NO_ENTRY : B2DGPU_ERROR_NOT_IMPLEMENTED, "signature not implemented by the GPU runtime");
  out->fill_func = fill_func_token;
  // Never called either: carries the signature, because RenderCommand::_dispatch_data overwrites the command's
  // _signature (rendercommand_p.h:169-174) and the batch consumer needs it back (B2DGPU_DISPATCH_SIGNATURE).
  out->fetch_func = reinterpret_cast<b2dgpu_fill_func>(uintptr_t(B2DGPU_DISPATCH_TAG) | uintptr_t(signature));
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_runtime_test(b2dgpu_runtime* rt, uint32_t signature, b2dgpu_dispatch_data* out, void*) {
  return runtime_lookup(rt, signature, out, true);
}
extern "C" b2dgpu_result b2dgpu_runtime_get(b2dgpu_runtime* rt, uint32_t signature, b2dgpu_dispatch_data* out, void*) {
  return runtime_lookup(rt, signature, out, false);
}

// ---------------------------------------------------------------------------------------------------------------
// Runtime
// ---------------------------------------------------------------------------------------------------------------
static void make_bayer_table(uint8_t* t) {
  // 16x16 ordered-dither matrix, each row stored twice (32 entries) - blend2d/tables/tables_p.h:452-473.  Generated:
  // rank = bit-interleave of (x ^ y, x) from the least significant coordinate bit to the most significant rank bits,
  // value = rank - (rank >= 128).
  for (uint32_t y = 0; y < 16; y++)
    for (uint32_t x = 0; x < 16; x++) {
      uint32_t r = 0;
      for (uint32_t i = 0; i < 4; i++) {
        uint32_t xb = (x >> i) & 1u, yb = (y >> i) & 1u;
        r |= ((xb ^ yb) << (2 * (3 - i) + 1)) | (xb << (2 * (3 - i)));
      }
      uint8_t v = uint8_t(r - (r >= 128u));
      t[y * 32 + x] = v;
      t[y * 32 + 16 + x] = v;
    }
}

static void runtime_destroy_thunk(b2dgpu_runtime* rt) { b2dgpu_runtime_destroy(rt); }

extern "C" b2dgpu_result b2dgpu_runtime_create(const b2dgpu_create_info* info, b2dgpu_runtime** out) {
  if (!out) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_runtime_create: out is null");
  *out = nullptr;
  int device = info ? info->device : 0;

  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(B2DGPU_ERROR_NOT_INITIALIZED, "no CUDA device available (there is no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : nullptr);
  if (device < 0 || device >= count) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_runtime_create: invalid device ordinal");
  CU_TRY(cudaSetDevice(device));

  b2dgpu_runtime* rt = new (std::nothrow) b2dgpu_runtime();
  if (!rt) return fail(B2DGPU_ERROR_OUT_OF_MEMORY, "b2dgpu_runtime_create: out of host memory");
  rt->runtime_type = 2;
  rt->runtime_flags = 1;
  rt->runtime_size = uint16_t(sizeof(b2dgpu_runtime));
  rt->_pad = 0;
  rt->destroy_fn = runtime_destroy_thunk;
  rt->test_fn = [](b2dgpu_runtime* r, uint32_t s, b2dgpu_dispatch_data* o, void* c) { return b2dgpu_runtime_test(r, s, o, c); };
  rt->get_fn = [](b2dgpu_runtime* r, uint32_t s, b2dgpu_dispatch_data* o, void* c) { return b2dgpu_runtime_get(r, s, o, c); };
  rt->magic = kRuntimeMagic;
  rt->device = device;
  rt->stream = nullptr;
  rt->own_stream = false;
  rt->prep_stream = nullptr; rt->prep_ready = nullptr; rt->slot_next = 0;
  rt->slot_done[0] = rt->slot_done[1] = nullptr; rt->slot_busy[0] = rt->slot_busy[1] = false;
  rt->d_bayer = nullptr; rt->d_pixel_counter = nullptr; rt->d_scalars = nullptr; rt->h_scalars = nullptr;
  rt->staging_next = 0;
  rt->bin_capacity = 0; rt->h_bin_state = nullptr; rt->edge_pair_capacity = 0;
  rt->glyph_cache_words = 0; rt->glyph_cache_id = 0; rt->glyph_cache_ready = nullptr;
  rt->profiling = false;
  rt->count_pixels = true;
  rt->sm_count = 148;
  { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) rt->sm_count = v; }
  memset(&rt->stats, 0, sizeof(rt->stats));

  if (info && info->stream) rt->stream = (cudaStream_t)info->stream;
  else {
    e = cudaStreamCreateWithFlags(&rt->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete rt; return cuda_fail(e, "cudaStreamCreate"); }
    rt->own_stream = true;
  }

  uint8_t bayer[512];
  make_bayer_table(bayer);
  if ((e = cudaMalloc((void**)&rt->d_bayer, 512)) != cudaSuccess ||
      (e = cudaMemcpy(rt->d_bayer, bayer, 512, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMalloc((void**)&rt->d_pixel_counter, 8)) != cudaSuccess ||
      (e = cudaMemset(rt->d_pixel_counter, 0, 8)) != cudaSuccess ||
      (e = cudaMalloc((void**)&rt->d_scalars, 64)) != cudaSuccess ||
      (e = cudaMemset(rt->d_scalars, 0, 64)) != cudaSuccess ||
      (e = cudaMallocHost((void**)&rt->h_scalars, 64)) != cudaSuccess ||
      (e = cudaMallocHost((void**)&rt->h_bin_state, 16)) != cudaSuccess) {
    b2dgpu_runtime_destroy(rt);
    return cuda_fail(e, "b2dgpu_runtime_create: device allocation");
  }
  rt->h_bin_state[0] = 0; rt->h_bin_state[1] = 1; rt->h_bin_state[2] = 1; rt->h_bin_state[3] = 0;
  e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&rt->staging[i].free_event, cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&rt->slot_done[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rt->prep_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rt->glyph_cache_ready, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    b2dgpu_runtime_destroy(rt);
    return cuda_fail(e, "b2dgpu_runtime_create: cudaEventCreate");
  }
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&rt->prep_stream, cudaStreamNonBlocking, hi)) != cudaSuccess) {
      b2dgpu_runtime_destroy(rt);
      return cuda_fail(e, "b2dgpu_runtime_create: cudaStreamCreate(prep)");
    }
  }
  {
    std::lock_guard<std::mutex> g(g_registry_mutex);
    g_runtimes.push_back(rt);
    rt->profiling = g_profiling_default;
    rt->count_pixels = g_count_pixels_default;
  }
  *out = rt;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_runtime_destroy(b2dgpu_runtime* rt) {
  if (!rt || rt->magic != kRuntimeMagic) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_runtime_destroy: invalid runtime");
  cudaSetDevice(rt->device);
  cudaStreamSynchronize(rt->stream);
  {
    bool registered;
    { std::lock_guard<std::mutex> g(g_registry_mutex); registered = registry_has(g_runtimes, rt); }
    if (registered) {
      b2dgpu_stats last;
      if (b2dgpu_get_stats(rt, &last, 0) == B2DGPU_SUCCESS) {
        std::lock_guard<std::mutex> g(g_registry_mutex);
        stats_add(g_retired_stats, last);
      }
      std::lock_guard<std::mutex> g(g_registry_mutex);
      registry_remove(g_runtimes, rt);
    }
  }
  if (rt->prep_stream) { cudaStreamSynchronize(rt->prep_stream); cudaStreamDestroy(rt->prep_stream); }
  for (int i = 0; i < 2; i++) if (rt->slot_done[i]) cudaEventDestroy(rt->slot_done[i]);
  if (rt->prep_ready) cudaEventDestroy(rt->prep_ready);
  if (rt->d_bayer) cudaFree(rt->d_bayer);
  if (rt->d_pixel_counter) cudaFree(rt->d_pixel_counter);
  if (rt->d_scalars) cudaFree(rt->d_scalars);
  if (rt->h_scalars) cudaFreeHost(rt->h_scalars);
  for (int i = 0; i < 2; i++) rt->staging[i].release();
  for (cudaEvent_t e : rt->prof_events) cudaEventDestroy(e);
  rt->image_staging.release();
  rt->oneshot_block[0].release();
  rt->oneshot_block[1].release();
  rt->oneshot_edges.release();
  rt->bins.release();
  rt->glyph_cache.release();
  if (rt->glyph_cache_ready) cudaEventDestroy(rt->glyph_cache_ready);
  if (rt->h_bin_state) cudaFreeHost(rt->h_bin_state);
  if (rt->own_stream) cudaStreamDestroy(rt->stream);
  rt->magic = 0;
  delete rt;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_sync(b2dgpu_runtime* rt) {
  if (!rt || rt->magic != kRuntimeMagic) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_sync: invalid runtime");
  CU_TRY(cudaStreamSynchronize(rt->stream));
  // The edge writer raises a flag instead of writing past the edge buffer (k_write_edges); surface it here.
  uint32_t flag = 0;
  CU_TRY(cudaMemcpy(&flag, rt->d_scalars + 1, 4, cudaMemcpyDeviceToHost));
  if (flag) {
    CU_TRY(cudaMemset(rt->d_scalars + 1, 0, 4));
    return fail(B2DGPU_ERROR_INVALID_STATE, "b2dgpu_sync: the device edge builder ran out of edge storage; the last render is incomplete");
  }
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_get_stats(b2dgpu_runtime* rt, b2dgpu_stats* out, int reset) {
  if (!rt || rt->magic != kRuntimeMagic || !out) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_get_stats: invalid argument");
  std::lock_guard<std::mutex> lock(rt->mutex);
  cudaSetDevice(rt->device);
  unsigned long long px = 0;
  CU_TRY(cudaMemcpyAsync(rt->h_scalars + 8, rt->d_pixel_counter, 8, cudaMemcpyDeviceToHost, rt->stream));
  CU_TRY(cudaStreamSynchronize(rt->stream));
  memcpy(&px, rt->h_scalars + 8, 8);
  rt->stats.pixels_composited = px;
  for (size_t i = 0; i + 2 < rt->prof_events.size(); i += 3) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, rt->prof_events[i], rt->prof_events[i + 1]);
    cudaEventElapsedTime(&b, rt->prof_events[i + 1], rt->prof_events[i + 2]);
    rt->stats.build_kernels_ms += a;
    rt->stats.tile_kernel_ms += b;
    rt->stats.tile_kernel_launches += 1;
  }
  for (cudaEvent_t e : rt->prof_events) cudaEventDestroy(e);
  rt->prof_events.clear();
  *out = rt->stats;
  if (reset) {
    memset(&rt->stats, 0, sizeof(rt->stats));
    CU_TRY(cudaMemsetAsync(rt->d_pixel_counter, 0, 8, rt->stream));
  }
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_set_profiling(b2dgpu_runtime* rt, int enabled) {
  if (!rt || rt->magic != kRuntimeMagic) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_set_profiling: invalid runtime");
  std::lock_guard<std::mutex> lock(rt->mutex);
  rt->profiling = enabled != 0;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_set_pixel_counting(b2dgpu_runtime* rt, int enabled) {
  if (!rt || rt->magic != kRuntimeMagic) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_set_pixel_counting: invalid runtime");
  std::lock_guard<std::mutex> lock(rt->mutex);
  rt->count_pixels = enabled != 0;
  return B2DGPU_SUCCESS;
}

extern "C" const char* b2dgpu_last_error_message(void) { return g_last_error.c_str(); }
extern "C" uint32_t b2dgpu_abi_version(void) { return 1; }

// ---------------------------------------------------------------------------------------------------------------
// Targets
// ---------------------------------------------------------------------------------------------------------------
static int format_bpp(uint32_t format) {
  switch (format) {
    case B2DGPU_FORMAT_PRGB32: case B2DGPU_FORMAT_XRGB32: case B2DGPU_FORMAT_FRGB32: case B2DGPU_FORMAT_ZERO32: return 4;
    case B2DGPU_FORMAT_A8: return 1;
    default: return 0;
  }
}

extern "C" b2dgpu_result b2dgpu_target_create_slab(b2dgpu_runtime* rt, int32_t w, int32_t full_h, int32_t y0, int32_t y1, uint32_t format, b2dgpu_target** out) {
  if (!rt || rt->magic != kRuntimeMagic || !out) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target_create: invalid argument");
  int bpp = format_bpp(format);
  if (!bpp || w <= 0 || full_h <= 0 || w > 65535 || full_h > 65535 || y0 < 0 || y1 > full_h || y0 >= y1)
    return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target_create: invalid size or format");
  cudaSetDevice(rt->device);
  b2dgpu_target* t = new (std::nothrow) b2dgpu_target();
  if (!t) return fail(B2DGPU_ERROR_OUT_OF_MEMORY, "b2dgpu_target_create: out of host memory");
  t->rt = rt; t->w = w; t->h = y1 - y0; t->full_h = full_h; t->y0 = y0; t->format = format; t->bpp = bpp;
  t->padded_w = int(align_up(size_t(w), kTileW));
  t->padded_h = int(align_up(size_t(t->h), kTileH));
  t->stride = size_t(t->padded_w) * bpp;
  t->d_pixels = nullptr;
  cudaError_t e = cudaMalloc((void**)&t->d_pixels, t->stride * t->padded_h);
  if (e != cudaSuccess) { delete t; return cuda_fail(e, "b2dgpu_target_create: cudaMalloc(canvas)"); }
  e = cudaMemsetAsync(t->d_pixels, 0, t->stride * t->padded_h, rt->stream);
  if (e != cudaSuccess) { cudaFree(t->d_pixels); delete t; return cuda_fail(e, "b2dgpu_target_create: cudaMemset"); }
  t->rendered = nullptr; t->rendered_valid = false;
  e = cudaEventCreateWithFlags(&t->rendered, cudaEventDisableTiming);
  if (e != cudaSuccess) { cudaFree(t->d_pixels); delete t; return cuda_fail(e, "b2dgpu_target_create: cudaEventCreate"); }
  { std::lock_guard<std::mutex> g(g_registry_mutex); g_targets.push_back(t); }
  *out = t;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_target_create(b2dgpu_runtime* rt, int32_t w, int32_t h, uint32_t format, b2dgpu_target** out) {
  return b2dgpu_target_create_slab(rt, w, h, 0, h, format, out);
}

extern "C" b2dgpu_result b2dgpu_target_destroy(b2dgpu_target* t) {
  if (!t) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target_destroy: null");
  { std::lock_guard<std::mutex> g(g_registry_mutex); registry_remove(g_targets, t); }
  cudaSetDevice(t->rt->device);
  cudaStreamSynchronize(t->rt->stream);
  cudaFree(t->d_pixels);
  if (t->rendered) cudaEventDestroy(t->rendered);
  delete t;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_target_clear(b2dgpu_target* t) {
  if (!t) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target_clear: null");
  cudaSetDevice(t->rt->device);
  CU_TRY(cudaMemsetAsync(t->d_pixels, 0, t->stride * t->padded_h, t->rt->stream));
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_target_wait(b2dgpu_target* t, void* stream) {
  if (!t) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target_wait: invalid argument");
  if (!t->rendered_valid) return B2DGPU_SUCCESS;           // nothing was rendered into it yet
  cudaSetDevice(t->rt->device);
  CU_TRY(cudaStreamWaitEvent((cudaStream_t)stream, t->rendered, 0));
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_target_device_view(b2dgpu_target* t, void** dev_ptr, intptr_t* stride, int32_t* padded_w, int32_t* padded_h) {
  if (!t) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target_device_view: null");
  if (dev_ptr) *dev_ptr = t->d_pixels;
  if (stride) *stride = intptr_t(t->stride);
  if (padded_w) *padded_w = t->padded_w;
  if (padded_h) *padded_h = t->padded_h;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_host_register(b2dgpu_runtime* rt, void* pixels, size_t bytes) {
  if (!rt || rt->magic != kRuntimeMagic || !pixels || !bytes) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_host_register: invalid argument");
  cudaSetDevice(rt->device);
  cudaError_t e = cudaHostRegister(pixels, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return B2DGPU_SUCCESS; }
  if (e != cudaSuccess) return cuda_fail(e, "b2dgpu_host_register");
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_host_unregister(b2dgpu_runtime* rt, void* pixels) {
  if (!rt || rt->magic != kRuntimeMagic || !pixels) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_host_unregister: invalid argument");
  cudaSetDevice(rt->device);
  cudaStreamSynchronize(rt->stream);
  cudaError_t e = cudaHostUnregister(pixels);
  if (e != cudaSuccess) { cudaGetLastError(); }
  return B2DGPU_SUCCESS;
}

static bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// The host image covers the FULL image; a slab target transfers only its own rows [y0, y0 + h).
static b2dgpu_result target_copy(b2dgpu_target* t, const b2dgpu_image_data* img, bool upload) {
  if (!t || !img || !img->pixel_data) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target copy: invalid argument");
  if (img->w != t->w || img->h != t->full_h || format_bpp(img->format) != t->bpp)
    return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_target copy: image size/format mismatch");
  b2dgpu_runtime* rt = t->rt;
  std::lock_guard<std::mutex> lock(rt->mutex);
  cudaSetDevice(rt->device);
  size_t row_bytes = size_t(t->w) * t->bpp;
  uint8_t* host = static_cast<uint8_t*>(img->pixel_data) + intptr_t(t->y0) * img->stride;
  size_t bytes = row_bytes * t->h;
  if (host_is_pinned(host) && host_is_pinned(host + intptr_t(t->h - 1) * img->stride + row_bytes - 1)) {
    // Page-locked image (b2dgpu_host_register): one direct 2-D DMA, no staging copy.
    if (upload) {
      CU_TRY(cudaMemcpy2DAsync(t->d_pixels, t->stride, host, size_t(img->stride), row_bytes, t->h, cudaMemcpyHostToDevice, rt->stream));
      CU_TRY(cudaStreamSynchronize(rt->stream));               // the caller may modify the image right after
      rt->stats.h2d_bytes += bytes;
    }
    else {
      CU_TRY(cudaMemcpy2DAsync(host, size_t(img->stride), t->d_pixels, t->stride, row_bytes, t->h, cudaMemcpyDeviceToHost, rt->stream));
      CU_TRY(cudaStreamSynchronize(rt->stream));
      rt->stats.d2h_bytes += bytes;
    }
    return B2DGPU_SUCCESS;
  }
  // Pageable host memory: stage through a pinned buffer so the copy runs at full PCIe rate and stays stream ordered.
  CU_TRY(rt->image_staging.ensure(bytes));
  uint8_t* stage = static_cast<uint8_t*>(rt->image_staging.ptr);
  if (upload) {
    CU_TRY(cudaStreamSynchronize(rt->stream));                 // previous use of the staging buffer
    for (int y = 0; y < t->h; y++) memcpy(stage + size_t(y) * row_bytes, host + intptr_t(y) * img->stride, row_bytes);
    CU_TRY(cudaMemcpy2DAsync(t->d_pixels, t->stride, stage, row_bytes, row_bytes, t->h, cudaMemcpyHostToDevice, rt->stream));
    rt->stats.h2d_bytes += bytes;
  }
  else {
    CU_TRY(cudaMemcpy2DAsync(stage, row_bytes, t->d_pixels, t->stride, row_bytes, t->h, cudaMemcpyDeviceToHost, rt->stream));
    CU_TRY(cudaStreamSynchronize(rt->stream));
    for (int y = 0; y < t->h; y++) memcpy(host + intptr_t(y) * img->stride, stage + size_t(y) * row_bytes, row_bytes);
    rt->stats.d2h_bytes += bytes;
  }
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_target_upload(b2dgpu_target* t, const b2dgpu_image_data* src) { return target_copy(t, src, true); }
extern "C" b2dgpu_result b2dgpu_target_download(b2dgpu_target* t, const b2dgpu_image_data* dst) { return target_copy(t, dst, false); }

// ---------------------------------------------------------------------------------------------------------------
// Batch serialisation
// ---------------------------------------------------------------------------------------------------------------
struct BlobRef { const void* host; size_t bytes; size_t offset; };


static SolidFill detect_solid_fill(const b2dgpu_batch_view* v) {
  SolidFill f; memset(&f, 0, sizeof(f));
  if (v->command_count != 1) return f;
  const b2dgpu_command& c = v->commands[0];
  const uint32_t op = B2DGPU_SIG_COMP_OP(c.signature);
  if (c.type != B2DGPU_CMD_FILL_BOX_A) return f;
  if (op != 0u /* SrcOver */ && op != 1u /* SrcCopy */) return f;
  if (c.alpha == 0) return f;
  memcpy(f.box, c.box, sizeof(f.box)); f.comp_op = op; f.alpha = c.alpha; f.prgb32 = c.solid_prgb32;
  f.fetch_type = B2DGPU_SIG_FETCH_TYPE(c.signature); f.src_format = B2DGPU_SIG_SRC_FORMAT(c.signature); f.fetch_index = c.fetch_index;
  f.ok = f.fetch_type == B2DGPU_FETCH_SOLID;
  f.one = !f.ok;
  if (f.fetch_type >= B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD && f.fetch_index < v->fetch_count && v->fetch_data)
    f.lut_entries = v->fetch_data[f.fetch_index].gradient.lut.size;
  return f;
}

struct FetchUse { uint32_t fetch_type; uint32_t src_format; bool used; };

// Callers built against the first layout of b2dgpu_batch_view pass a shorter struct: the missing tail reads as zero.
static bool normalize_view(const b2dgpu_batch_view* v, b2dgpu_batch_view* out) {
  if (!v || v->struct_size < B2DGPU_BATCH_VIEW_SIZE_V1) return false;
  memset(out, 0, sizeof(*out));
  memcpy(out, v, v->struct_size < sizeof(*out) ? v->struct_size : sizeof(*out));
  out->struct_size = uint32_t(sizeof(*out));
  return true;
}

static b2dgpu_result validate_glyph_instances(const b2dgpu_batch_view* v) {
  if (!v->glyph_instance_count) {
    if (v->generated_vertex_count || v->generated_segment_count) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: generated counts without glyph instances");
    return B2DGPU_SUCCESS;
  }
  if (!v->glyph_instances || !v->glyph_cache || !v->glyph_cache_words) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: glyph instances without a cache");
  // Instances must tile the generated ranges in order, so that every generated vertex / segment is written exactly once.
  uint64_t vtx = v->vertex_count, seg = v->segment_count;
  for (uint32_t i = 0; i < v->glyph_instance_count; i++) {
    const b2dgpu_glyph_instance& gi = v->glyph_instances[i];
    if (uint64_t(gi.blob_offset) + 3u > v->glyph_cache_words) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: glyph instance outside the cache");
    const b2d::GlyphBlobView g = b2d::glyph_blob_view(v->glyph_cache + gi.blob_offset);
    if (uint64_t(gi.blob_offset) + g.total_words() > v->glyph_cache_words) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: truncated glyph cache entry");
    if (gi.vertex_base != vtx || gi.segment_base != seg) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: glyph instances do not tile the generated ranges");
    if (gi.command >= v->command_count) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: glyph instance names no command");
    const b2dgpu_command& c = v->commands[gi.command];
    if (c.type != B2DGPU_CMD_FILL_GEOMETRY || seg < c.data_offset || seg + g.segments > uint64_t(c.data_offset) + c.data_count)
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: glyph instance outside the segment range of its command");
    for (uint32_t k = 0; k < g.segments; k++) {
      const uint32_t p0 = g.segment_words()[2 * k], p1 = g.segment_words()[2 * k + 1] >> 2, kind = g.segment_words()[2 * k + 1] & 3u;
      const uint32_t extra = kind == B2DGPU_SEG_LINE ? 0u : kind == B2DGPU_SEG_QUAD ? 1u : 2u;
      if (p0 >= g.vertices || uint64_t(p1) + extra >= g.vertices) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: glyph cache segment refers outside its glyph");
    }
    vtx += g.vertices; seg += g.segments;
  }
  if (vtx != uint64_t(v->vertex_count) + v->generated_vertex_count || seg != uint64_t(v->segment_count) + v->generated_segment_count)
    return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: generated counts do not match the glyph instances");
  if (vtx > 0x3FFFFFF0u || seg > 0xFFFFFFF0u) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: too many generated vertices");
  return B2DGPU_SUCCESS;
}

static b2dgpu_result validate_batch(const b2dgpu_batch_view* v) {
  if (!v || v->struct_size < sizeof(b2dgpu_batch_view)) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: bad struct_size");
  if (v->command_count && !v->commands) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: commands is null");
  if ((v->fetch_count && !v->fetch_data) || (v->edge_count && !v->edges) || (v->vertex_count && !v->vertices) ||
      (v->segment_count && !v->segments) || (v->geometry_state_count && !v->geometry_states))
    return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: a non-empty array is null");
  for (uint32_t i = 0; i < v->command_count; i++) {
    const b2dgpu_command& c = v->commands[i];
    if (c.type < B2DGPU_CMD_FILL_BOX_A || c.type > B2DGPU_CMD_FILL_BOX_MASK_A) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: unknown command type");
    if (c.type == B2DGPU_CMD_FILL_BOX_MASK_A) {
      if (c.reserved[0] >= v->fetch_count) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: mask index out of range");
      const b2dgpu_pattern_source& ms = v->fetch_data[c.reserved[0]].pattern.src;
      if (!(c.box[0] < c.box[2] && c.box[1] < c.box[3]) || ms.w < c.box[2] - c.box[0] || ms.h < c.box[3] - c.box[1])
        return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: mask smaller than its box");
    }
    if (!signature_supported(c.signature)) return fail(B2DGPU_ERROR_NOT_IMPLEMENTED, "batch view: command signature not implemented");
    if (c.alpha > 255) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: alpha out of range");
    if (B2DGPU_SIG_FETCH_TYPE(c.signature) != B2DGPU_FETCH_SOLID && c.fetch_index >= v->fetch_count)
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: fetch_index out of range");
    if (c.type == B2DGPU_CMD_FILL_ANALYTIC && (uint64_t(c.data_offset) + c.data_count > v->edge_count))
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: edge range out of bounds");
    if (c.type == B2DGPU_CMD_FILL_GEOMETRY) {
      if (uint64_t(c.data_offset) + c.data_count > uint64_t(v->segment_count) + v->generated_segment_count) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: segment range out of bounds");
      if (c.state_index >= v->geometry_state_count) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: state_index out of range");
    }
    if ((c.type == B2DGPU_CMD_FILL_BOX_A || c.type == B2DGPU_CMD_FILL_BOX_U) && !(c.box[0] < c.box[2] && c.box[1] < c.box[3]))
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: empty box");
  }
  for (uint32_t i = 0; i < v->segment_count; i++) {
    const b2dgpu_segment& s = v->segments[i];
    uint32_t kind = s.p1_kind & 3u, i1 = s.p1_kind >> 2;
    uint32_t extra = kind == B2DGPU_SEG_LINE ? 0u : kind == B2DGPU_SEG_QUAD ? 1u : 2u;
    if (s.p0 >= v->vertex_count || uint64_t(i1) + extra >= v->vertex_count || s.command >= v->command_count)
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: segment references out of range");
    // The device edge builder reads the geometry state of the segment's command and accumulates the command's bounding
    // box: the segment has to lie inside the range of a FILL_GEOMETRY command (whose state_index was checked above).
    const b2dgpu_command& c = v->commands[s.command];
    if (c.type != B2DGPU_CMD_FILL_GEOMETRY || i < c.data_offset || uint64_t(i) >= uint64_t(c.data_offset) + c.data_count)
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: segment does not belong to the geometry command it names");
  }
  return validate_glyph_instances(v);
}

static void plan_layout(const b2dgpu_batch_view* v, size_t blob_bytes, size_t gen_lut_bytes, BlockLayout& L) {
  size_t off = 0;
  auto add = [&](size_t bytes) { off = align_up(off, 256); size_t o = off; off += bytes; return o; };
  const size_t total_segments = size_t(v->segment_count) + v->generated_segment_count;
  L.commands = add(sizeof(b2dgpu_command) * v->command_count);
  L.fetch_data = add(sizeof(b2dgpu_fetch_data) * v->fetch_count);
  L.states = add(sizeof(b2dgpu_geometry_state) * v->geometry_state_count);
  L.supplied_edges = add(sizeof(b2dgpu_edge) * v->edge_count);
  L.blobs = add(blob_bytes);
  L.instances = add(sizeof(b2dgpu_glyph_instance) * v->glyph_instance_count);
  L.lut_requests = add(sizeof(b2dgpu_lut_request) * v->lut_request_count);
  L.lut_stops = add(sizeof(b2dgpu_gradient_stop) * v->lut_stop_count);
  L.lut_offsets = add(sizeof(uint32_t) * v->lut_request_count);
  L.vertices = add(sizeof(double) * 2 * (size_t(v->vertex_count) + v->generated_vertex_count));
  L.upload1_bytes = L.vertices + sizeof(double) * 2 * v->vertex_count;
  L.segments = add(sizeof(b2dgpu_segment) * total_segments);
  L.seg_upload_bytes = sizeof(b2dgpu_segment) * v->segment_count;
  L.staging_segments = align_up(L.upload1_bytes, 256);
  L.upload_bytes = align_up(L.staging_segments + L.seg_upload_bytes, 256);
  L.gen_luts = add(gen_lut_bytes);
  L.seg_counts = add(sizeof(uint32_t) * (total_segments + 1));
  L.seg_offsets = add(sizeof(uint32_t) * (total_segments + 1));
  L.scan_scratch = add(sizeof(uint32_t) * scan_scratch_items(uint32_t(total_segments)));
  L.bbox_fixed = add(sizeof(int4) * v->command_count);
  L.bbox_px = add(sizeof(int4) * v->command_count);
  L.cmd_edges = add(sizeof(uint2) * v->command_count);
  L.total_bytes = align_up(off, 256);
}

// Serialises `v` into `host_block` (size lay.upload_bytes) with fetch-data pointers patched to `dev_base + offset`.
static b2dgpu_result serialize_batch(const b2dgpu_batch_view* v, std::vector<BlobRef>& blobs, const BlockLayout& L, uint8_t* host_block, const uint8_t* dev_base) {
  memcpy(host_block + L.commands, v->commands, sizeof(b2dgpu_command) * v->command_count);
  if (v->vertex_count) memcpy(host_block + L.vertices, v->vertices, sizeof(double) * 2 * v->vertex_count);
  if (v->segment_count) memcpy(host_block + L.staging_segments, v->segments, sizeof(b2dgpu_segment) * v->segment_count);
  if (v->glyph_instance_count) memcpy(host_block + L.instances, v->glyph_instances, sizeof(b2dgpu_glyph_instance) * v->glyph_instance_count);
  if (v->geometry_state_count) memcpy(host_block + L.states, v->geometry_states, sizeof(b2dgpu_geometry_state) * v->geometry_state_count);
  if (v->edge_count) memcpy(host_block + L.supplied_edges, v->edges, sizeof(b2dgpu_edge) * v->edge_count);
  for (const BlobRef& b : blobs) memcpy(host_block + L.blobs + b.offset, b.host, b.bytes);

  b2dgpu_fetch_data* fd = reinterpret_cast<b2dgpu_fetch_data*>(host_block + L.fetch_data);
  if (v->fetch_count) memcpy(fd, v->fetch_data, sizeof(b2dgpu_fetch_data) * v->fetch_count);
  return B2DGPU_SUCCESS;
}

// Collects the host memory referenced by fetch data (gradient LUTs, pattern pixels), de-duplicated by address.
static b2dgpu_result collect_blobs(const b2dgpu_batch_view* v, std::vector<FetchUse>& uses, std::vector<BlobRef>& blobs,
                                   std::vector<size_t>& fetch_blob, size_t& blob_bytes, const std::vector<int32_t>& fetch_lut_request) {
  uses.assign(v->fetch_count, FetchUse{0, 0, false});
  for (uint32_t i = 0; i < v->command_count; i++) {
    const b2dgpu_command& c = v->commands[i];
    uint32_t ft = B2DGPU_SIG_FETCH_TYPE(c.signature);
    if (ft == B2DGPU_FETCH_SOLID) continue;
    FetchUse& u = uses[c.fetch_index];
    uint32_t sf = B2DGPU_SIG_SRC_FORMAT(c.signature);
    if (u.used && (u.fetch_type != ft || u.src_format != sf)) {
      // The same FetchData may legitimately be used with different fill types only; fetch type/format must agree.
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: one fetch_data entry used with two different fetch types");
    }
    u.fetch_type = ft; u.src_format = sf; u.used = true;
  }
  // Image masks travel like A8 pattern pixels: a fetch_data entry whose pattern.src points at the mask rows.
  for (uint32_t i = 0; i < v->command_count; i++) {
    const b2dgpu_command& c = v->commands[i];
    if (c.type != B2DGPU_CMD_FILL_BOX_MASK_A) continue;
    FetchUse& u = uses[c.reserved[0]];
    if (u.used && (u.fetch_type != B2DGPU_FETCH_PATTERN_ALIGNED_BLIT || u.src_format != B2DGPU_FORMAT_A8))
      return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: a mask entry is also used as a source");
    u.fetch_type = B2DGPU_FETCH_PATTERN_ALIGNED_BLIT; u.src_format = B2DGPU_FORMAT_A8; u.used = true;
  }

  std::unordered_map<const void*, size_t> seen;
  fetch_blob.assign(v->fetch_count, size_t(-1));
  blob_bytes = 0;
  for (uint32_t i = 0; i < v->fetch_count; i++) {
    if (!uses[i].used) continue;
    const b2dgpu_fetch_data& fd = v->fetch_data[i];
    const void* host = nullptr;
    size_t bytes = 0;
    if (uses[i].fetch_type >= B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD) {
      bool dither = uses[i].fetch_type == B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_PAD || uses[i].fetch_type == B2DGPU_FETCH_GRADIENT_LINEAR_DITHER_ROR ||
                    uses[i].fetch_type == B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_PAD || uses[i].fetch_type == B2DGPU_FETCH_GRADIENT_RADIAL_DITHER_ROR ||
                    uses[i].fetch_type == B2DGPU_FETCH_GRADIENT_CONIC_DITHER;
      host = fd.gradient.lut.data;
      bytes = size_t(fd.gradient.lut.size) * (dither ? 8 : 4);
      if (!host && bytes && !dither && fetch_lut_request[i] >= 0) continue;      // built on the device (k_build_luts)
      if (!host || !bytes) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: gradient without a LUT");
    }
    else {
      const b2dgpu_pattern_source& s = fd.pattern.src;
      if (!s.pixel_data || s.w <= 0 || s.h <= 0 || s.stride <= 0) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: invalid pattern source");
      host = s.pixel_data;
      bytes = size_t(s.stride) * size_t(s.h - 1) + size_t(s.w) * (uses[i].src_format == B2DGPU_FORMAT_A8 ? 1 : 4);
    }
    auto it = seen.find(host);
    if (it != seen.end() && blobs[it->second].bytes >= bytes) { fetch_blob[i] = it->second; continue; }
    BlobRef b; b.host = host; b.bytes = bytes; b.offset = align_up(blob_bytes, 256);
    blob_bytes = b.offset + bytes;
    seen[host] = blobs.size();
    fetch_blob[i] = blobs.size();
    blobs.push_back(b);
  }
  return B2DGPU_SUCCESS;
}

struct PreparedBatch {
  BlockLayout lay;
  std::vector<uint32_t> lut_table_offsets;          // per request, in words, inside the gen_luts section
  size_t gen_lut_bytes;
  std::vector<int32_t> fetch_lut_request;           // per fetch_data entry: its request or -1
  std::vector<BlobRef> blobs;
  std::vector<FetchUse> uses;
  std::vector<size_t> fetch_blob;
  bool has_analytic;
  bool stream_ok;
  int stream_box[4];
  SolidFill solid;
};

static b2dgpu_result prepare_batch(const b2dgpu_batch_view* v, PreparedBatch& pb) {
  b2dgpu_result r = validate_batch(v);
  if (r) return r;
  // device-built gradient tables
  pb.fetch_lut_request.assign(v->fetch_count, -1);
  pb.lut_table_offsets.assign(v->lut_request_count, 0u);
  pb.gen_lut_bytes = 0;
  if (v->lut_request_count) {
    if (!v->lut_requests || !v->lut_stops) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: table requests without stops");
    for (uint32_t i = 0; i < v->lut_request_count; i++) {
      const b2dgpu_lut_request& q = v->lut_requests[i];
      if (q.fetch_index >= v->fetch_count || !q.stop_count || uint64_t(q.stop_offset) + q.stop_count > v->lut_stop_count ||
          q.lut_size < 2u || q.lut_size > 65536u || q.lut_size != v->fetch_data[q.fetch_index].gradient.lut.size ||
          v->fetch_data[q.fetch_index].gradient.lut.data != nullptr || pb.fetch_lut_request[q.fetch_index] >= 0)
        return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: invalid gradient table request");
      for (uint32_t k = 0; k < q.stop_count; k++) {
        const double o = v->lut_stops[q.stop_offset + k].offset;
        if (!(o >= 0.0 && o <= 1.0) || (k && o < v->lut_stops[q.stop_offset + k - 1].offset))
          return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: gradient stops out of order");
      }
      pb.fetch_lut_request[q.fetch_index] = int32_t(i);
      pb.lut_table_offsets[i] = uint32_t(pb.gen_lut_bytes / 4);
      pb.gen_lut_bytes += align_up(size_t(q.lut_size) * 4, 256);
    }
  }
  size_t blob_bytes = 0;
  r = collect_blobs(v, pb.uses, pb.blobs, pb.fetch_blob, blob_bytes, pb.fetch_lut_request);
  if (r) return r;
  plan_layout(v, blob_bytes, pb.gen_lut_bytes, pb.lay);
  pb.has_analytic = false;
  for (uint32_t i = 0; i < v->command_count; i++) if (v->commands[i].type == B2DGPU_CMD_FILL_ANALYTIC) pb.has_analytic = true;
  pb.solid = detect_solid_fill(v);
  pb.stream_ok = v->command_count >= 1 && v->command_count <= 8;
  pb.stream_box[0] = pb.stream_box[1] = INT_MAX; pb.stream_box[2] = pb.stream_box[3] = INT_MIN;
  for (uint32_t i = 0; i < v->command_count && pb.stream_ok; i++) {
    const b2dgpu_command& c = v->commands[i];
    int b[4];
    if (c.type == B2DGPU_CMD_FILL_BOX_A) { b[0] = c.box[0]; b[1] = c.box[1]; b[2] = c.box[2]; b[3] = c.box[3]; }
    else if (c.type == B2DGPU_CMD_FILL_BOX_U) { b[0] = c.box[0] >> 8; b[1] = c.box[1] >> 8; b[2] = (c.box[2] + 0xFF) >> 8; b[3] = (c.box[3] + 0xFF) >> 8; }
    else { pb.stream_ok = false; break; }
    if (b[0] < pb.stream_box[0]) pb.stream_box[0] = b[0];
    if (b[1] < pb.stream_box[1]) pb.stream_box[1] = b[1];
    if (b[2] > pb.stream_box[2]) pb.stream_box[2] = b[2];
    if (b[3] > pb.stream_box[3]) pb.stream_box[3] = b[3];
  }
  return B2DGPU_SUCCESS;
}

// Brings the device mirror of the caller's glyph cache up to date (append-only: only the new tail travels).
static b2dgpu_result sync_glyph_cache(b2dgpu_runtime* rt, const b2dgpu_batch_view* v, cudaStream_t stream) {
  if (!v->glyph_instance_count) return B2DGPU_SUCCESS;
  if (v->glyph_cache_id != rt->glyph_cache_id || v->glyph_cache_words < rt->glyph_cache_words) {
    rt->glyph_cache_id = v->glyph_cache_id;
    rt->glyph_cache_words = 0;
  }
  const size_t need = size_t(v->glyph_cache_words) * 4;
  if (need > rt->glyph_cache.cap) {
    // the old mirror may still be read by kernels in flight on either stream
    CU_TRY(cudaStreamSynchronize(rt->stream));
    CU_TRY(cudaStreamSynchronize(rt->prep_stream));
    CU_TRY(rt->glyph_cache.ensure(need * 2));
    rt->glyph_cache_words = 0;
  }
  if (v->glyph_cache_words > rt->glyph_cache_words) {
    const size_t from = size_t(rt->glyph_cache_words) * 4;
    CU_TRY(cudaMemcpyAsync(static_cast<uint8_t*>(rt->glyph_cache.ptr) + from, reinterpret_cast<const uint8_t*>(v->glyph_cache) + from,
                           need - from, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaEventRecord(rt->glyph_cache_ready, stream));
    rt->stats.h2d_bytes += need - from;
    rt->glyph_cache_words = v->glyph_cache_words;
  }
  return B2DGPU_SUCCESS;
}

// Fills the pinned block and copies it to `dev_block`.
static b2dgpu_result upload_block(b2dgpu_runtime* rt, const b2dgpu_batch_view* v, PreparedBatch& pb, uint8_t* dev_block, cudaStream_t stream) {
  PinnedBuffer& st = rt->staging[rt->staging_next];
  rt->staging_next ^= 1;
  if (st.in_flight) { CU_TRY(cudaEventSynchronize(st.free_event)); st.in_flight = false; }
  CU_TRY(st.ensure(pb.lay.upload_bytes));
  uint8_t* host_block = static_cast<uint8_t*>(st.ptr);
  serialize_batch(v, pb.blobs, pb.lay, host_block, dev_block);

  // Patch host pointers inside FetchData to their device copies.
  b2dgpu_fetch_data* fd = reinterpret_cast<b2dgpu_fetch_data*>(host_block + pb.lay.fetch_data);
  if (v->lut_request_count) {
    memcpy(host_block + pb.lay.lut_requests, v->lut_requests, sizeof(b2dgpu_lut_request) * v->lut_request_count);
    memcpy(host_block + pb.lay.lut_stops, v->lut_stops, sizeof(b2dgpu_gradient_stop) * v->lut_stop_count);
    memcpy(host_block + pb.lay.lut_offsets, pb.lut_table_offsets.data(), sizeof(uint32_t) * v->lut_request_count);
  }
  for (uint32_t i = 0; i < v->fetch_count; i++) {
    if (!pb.uses[i].used) continue;
    if (pb.fetch_lut_request[i] >= 0 && pb.fetch_blob[i] == size_t(-1)) {
      fd[i].gradient.lut.data = dev_block + pb.lay.gen_luts + size_t(pb.lut_table_offsets[size_t(pb.fetch_lut_request[i])]) * 4;
      continue;
    }
    const uint8_t* dev = dev_block + pb.lay.blobs + pb.blobs[pb.fetch_blob[i]].offset;
    if (pb.uses[i].fetch_type >= B2DGPU_FETCH_GRADIENT_LINEAR_NN_PAD) fd[i].gradient.lut.data = dev;
    else fd[i].pattern.src.pixel_data = dev;
  }

  CU_TRY(cudaMemcpyAsync(dev_block, host_block, pb.lay.upload1_bytes, cudaMemcpyHostToDevice, stream));
  if (pb.lay.seg_upload_bytes)
    CU_TRY(cudaMemcpyAsync(dev_block + pb.lay.segments, host_block + pb.lay.staging_segments, pb.lay.seg_upload_bytes, cudaMemcpyHostToDevice, stream));
  CU_TRY(cudaEventRecord(st.free_event, stream));
  st.in_flight = true;
  rt->stats.h2d_bytes += pb.lay.upload1_bytes + pb.lay.seg_upload_bytes;
  if (v->lut_request_count) {
    LutParams LP;
    LP.requests = reinterpret_cast<const b2dgpu_lut_request*>(dev_block + pb.lay.lut_requests);
    LP.request_count = v->lut_request_count;
    LP.stops = reinterpret_cast<const b2dgpu_gradient_stop*>(dev_block + pb.lay.lut_stops);
    LP.table_offsets = reinterpret_cast<const uint32_t*>(dev_block + pb.lay.lut_offsets);
    LP.tables = reinterpret_cast<uint32_t*>(dev_block + pb.lay.gen_luts);
    rt->stats.kernel_launches += uint64_t(launch_build_luts(LP, stream));
  }
  return sync_glyph_cache(rt, v, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Rendering
// ---------------------------------------------------------------------------------------------------------------
struct RenderInput {
  SolidFill solid;                // the batch is ONE solid FillBoxA with SrcOver / SrcCopy: k_stream_solid
  uint8_t* block;
  const BlockLayout* lay;
  DevBuffer* edges;
  uint32_t command_count, supplied_edges, segment_count;      // segment_count: uploaded + generated
  uint32_t instance_count;
  int origin_x, origin_y;
  bool has_analytic;
  bool built_known;
  uint32_t built_edges;
  bool edges_staged;
  bool prepped;                   // init_bbox + count + scan already ran (on the prep stream) for this block
  bool stream_ok;                 // only box fills, few of them: eligible for the streaming compositor
  int stream_box[4];              // union of their boxes in pixels
};

static BuildParams make_build_params(b2dgpu_runtime* rt, const RenderInput& in) {
  const BlockLayout& L = *in.lay;
  uint8_t* blk = in.block;
  BuildParams B;
  B.vertices = reinterpret_cast<const double*>(blk + L.vertices);
  B.segments = reinterpret_cast<const b2dgpu_segment*>(blk + L.segments);
  B.segment_count = in.segment_count;
  B.commands = reinterpret_cast<const b2dgpu_command*>(blk + L.commands);
  B.states = reinterpret_cast<const b2dgpu_geometry_state*>(blk + L.states);
  B.seg_counts = reinterpret_cast<uint32_t*>(blk + L.seg_counts);
  B.seg_offsets = reinterpret_cast<uint32_t*>(blk + L.seg_offsets);
  B.edge_base = in.supplied_edges;
  B.cmd_bbox_fixed = reinterpret_cast<int4*>(blk + L.bbox_fixed);
  B.error_flag = rt->d_scalars + 1;
  B.edges = nullptr; B.edge_capacity = 0;
  return B;
}

// K1 pass 1 (edge counts per segment) + exclusive scan, on stream `s`.  These run on every render - they are part of
// the path even when the total is already known; the total is only read back (which synchronises `s`) the first
// time a geometry is seen, because the edge buffer has to be sized on the host.
static b2dgpu_result prep_block(b2dgpu_runtime* rt, RenderInput& in, cudaStream_t s, uint32_t* d_total, uint32_t* h_total, int& launches) {
  const BlockLayout& L = *in.lay;
  uint8_t* blk = in.block;
  BuildParams B = make_build_params(rt, in);
  launches += launch_init_bbox(B.cmd_bbox_fixed, in.command_count, s);
  if (in.instance_count) {
    // K0: the glyph instances write their vertices and segments behind the uploaded ones
    GlyphParams G;
    G.cache = static_cast<const uint32_t*>(rt->glyph_cache.ptr);
    G.instances = reinterpret_cast<const b2dgpu_glyph_instance*>(blk + L.instances);
    G.instance_count = in.instance_count;
    G.vertices = reinterpret_cast<double*>(blk + L.vertices);
    G.segments = reinterpret_cast<b2dgpu_segment*>(blk + L.segments);
    G.error_flag = rt->d_scalars + 1;
    CU_TRY(cudaStreamWaitEvent(s, rt->glyph_cache_ready, 0));
    launches += launch_glyph_instances(G, s);
  }
  if (in.segment_count) {
    launches += launch_count_edges(B, s);
    launches += launch_exclusive_scan(B.seg_counts, const_cast<uint32_t*>(B.seg_offsets), in.segment_count,
                                      reinterpret_cast<uint32_t*>(blk + L.scan_scratch), d_total, s);
    if (!in.built_known) {
      CU_TRY(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      in.built_edges = h_total[0];
      in.built_known = true;
      rt->stats.d2h_bytes += 4;
    }
  }
  else {
    in.built_edges = 0;
    in.built_known = true;
  }
  in.prepped = true;
  return B2DGPU_SUCCESS;
}

// Renders one prepared block into `ntargets` targets of the same image (slabs / stripes of one canvas): the geometry
// pass (K1: count, scan, write, bounding boxes) runs once, the per-target part (clip to the rows, band extents,
// compositing) once per target.
static b2dgpu_result render_block(b2dgpu_runtime* rt, b2dgpu_target* const* targets, uint32_t ntargets, RenderInput& in) {
  const BlockLayout& L = *in.lay;
  cudaStream_t s = rt->stream;
  uint8_t* blk = in.block;
  int launches = 0;

  const b2dgpu_command* d_cmds = reinterpret_cast<const b2dgpu_command*>(blk + L.commands);
  int4* d_bbox_fixed = reinterpret_cast<int4*>(blk + L.bbox_fixed);

  uint32_t* d_seg_offsets = reinterpret_cast<uint32_t*>(blk + L.seg_offsets);

  // Profiling events: owned by this guard until they are handed to rt->prof_events (an early error return frees them).
  struct EventTriple {
    cudaEvent_t e[3] = { nullptr, nullptr, nullptr };
    bool handed_over = false;
    ~EventTriple() { if (!handed_over) for (int i = 0; i < 3; i++) if (e[i]) cudaEventDestroy(e[i]); }
    cudaEvent_t& operator[](int i) { return e[i]; }
  } ev;
  if (rt->profiling) {
    for (int i = 0; i < 3; i++) CU_TRY(cudaEventCreate(&ev[i]));
    CU_TRY(cudaEventRecord(ev[0], s));
  }

  BuildParams B = make_build_params(rt, in);
  if (!in.prepped) {
    int prep_launches = 0;
    b2dgpu_result pr = prep_block(rt, in, s, rt->d_scalars, rt->h_scalars, prep_launches);
    if (pr) return pr;
    launches += prep_launches;
  }

  const size_t total_edges = size_t(in.supplied_edges) + in.built_edges;
  if (total_edges > 0xFFFFFFF0u) return fail(B2DGPU_ERROR_OUT_OF_MEMORY, "too many edges in one batch");
  {
    size_t need = total_edges * sizeof(b2dgpu_edge);
    if (need < 4096) need = 4096;
    if (need > in.edges->cap) {
      CU_TRY(cudaStreamSynchronize(s));
      CU_TRY(in.edges->ensure(need));
      in.edges_staged = false;
    }
    if (!in.edges_staged && in.supplied_edges) {
      CU_TRY(cudaMemcpyAsync(in.edges->ptr, blk + L.supplied_edges, sizeof(b2dgpu_edge) * in.supplied_edges, cudaMemcpyDeviceToDevice, s));
    }
    in.edges_staged = true;
  }
  B.edges = static_cast<b2dgpu_edge*>(in.edges->ptr);
  B.edge_capacity = uint32_t(in.edges->cap / sizeof(b2dgpu_edge));

  if (in.has_analytic)
    launches += launch_analytic_bbox(d_cmds, in.command_count, B.edges, d_bbox_fixed, s);
  if (in.segment_count && in.built_edges)
    launches += launch_write_edges(B, s);

  for (uint32_t ti = 0; ti < ntargets; ti++) {
  b2dgpu_target* t = targets[ti];
  FinalizeParams F;
  F.commands = d_cmds;
  F.command_count = in.command_count;
  F.seg_offsets = d_seg_offsets;
  F.edge_base = in.supplied_edges;
  F.cmd_bbox_fixed = d_bbox_fixed;
  F.cmd_bbox_px = reinterpret_cast<int4*>(blk + L.bbox_px);
  F.cmd_edges = reinterpret_cast<uint2*>(blk + L.cmd_edges);
  F.width = t->w;
  F.y_begin = t->y0;
  F.y_end = t->y0 + t->h;
  launches += launch_finalize_commands(F, s);

  TileParams T;
  T.dst = t->d_pixels;
  T.dst_stride = intptr_t(t->stride);
  T.tiles_x = t->padded_w / kTileW;
  const int tile_h = choose_tile_height(t->padded_w / kTileW, t->h, rt->sm_count);
  T.tiles_y = (t->h + tile_h - 1) / tile_h;           // padded_h is a multiple of the largest tile height
  T.y_begin = t->y0;
  T.commands = d_cmds;
  T.command_count = in.command_count;
  T.cmd_bbox_px = F.cmd_bbox_px;
  T.cmd_edges = F.cmd_edges;
  T.edges = B.edges;
  T.fetch_data = reinterpret_cast<const b2dgpu_fetch_data*>(blk + L.fetch_data);
  T.bayer = rt->d_bayer;
  T.origin_x = in.origin_x;
  T.origin_y = in.origin_y;
  T.pixel_counter = rt->count_pixels ? rt->d_pixel_counter : nullptr;
  T.band_off = nullptr; T.cell_cmd = nullptr; T.cell_ext = nullptr; T.bin_state = nullptr;
  T.cell_edge_off = nullptr; T.band_edges = nullptr;
  // Per-band command lists with x-extents (k_bin_*).  Their size is only known on the device: the buffer holds
  // `bin_capacity` cells (the dense bound tiles_y * commands when that is small); a render that needed more renders
  // without lists (every tile scans every command - slow but correct) and the buffer is grown for the next one.
  // A batch of boxes only on a small canvas (bl_bench rectangles on 512 x 600) is culled exactly by the pixel boxes;
  // when all tiles x commands tests are few the lists cost more than they save.
  const bool bin = in.has_analytic || in.segment_count ||
                   size_t(in.command_count) * size_t(T.tiles_x) * size_t(T.tiles_y) > (size_t(128) << 20);
  if (bin) {
    const size_t dense = size_t(in.command_count) * size_t(T.tiles_y);
    if (rt->h_bin_state[1] == 0u && rt->h_bin_state[0] > rt->bin_capacity)
      rt->bin_capacity = rt->h_bin_state[0] + rt->h_bin_state[0] / 4u;                    // reported by an earlier render
    size_t want = rt->bin_capacity;
    const size_t floor_cells = size_t(4) << 20;
    if (want < floor_cells) want = floor_cells;
    if (want < size_t(256) * in.command_count) want = size_t(256) * in.command_count;
    if (want > dense) want = dense;
    if (want > 0xFFFFFF00u) want = 0xFFFFFF00u;
    if (const char* e = getenv("B2DGPU_BIN_CAPACITY")) want = size_t(strtoull(e, nullptr, 10));   // test knob: force the fallback
    if (want < 1) want = 1;
    const uint32_t cap = uint32_t(want);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 256); const size_t o = off; off += bytes; return o; };
    const size_t o_state = take(16);
    const size_t o_cm_count = take(sizeof(uint32_t) * (size_t(in.command_count) + 1));
    const size_t o_cm_base = take(sizeof(uint32_t) * (size_t(in.command_count) + 1));
    const size_t o_band_count = take(sizeof(uint32_t) * (size_t(T.tiles_y) + 2));
    const size_t o_band_prefix = take(sizeof(uint32_t) * (size_t(T.tiles_y) + 2));
    const size_t o_band_off = take(sizeof(uint32_t) * (size_t(T.tiles_y) + 1));
    const size_t o_scratch = take(sizeof(uint32_t) * bin_scratch_items(in.command_count, T.tiles_y));
    const size_t o_cm_index = take(sizeof(uint32_t) * size_t(cap));
    const size_t o_cell_cmd = take(sizeof(uint32_t) * size_t(cap));
    const size_t o_cell_ext = take(sizeof(uint2) * size_t(cap));
    // per-cell edge lists: (edge, band) pairs; an edge of a flattened curve rarely spans more than two bands
    const bool with_edge_lists = total_edges != 0;
    if (rt->h_bin_state[1] != 0u && rt->h_bin_state[2] == 0u && rt->h_bin_state[3] > rt->edge_pair_capacity)
      rt->edge_pair_capacity = size_t(rt->h_bin_state[3]) + rt->h_bin_state[3] / 4u;      // reported by an earlier render
    size_t pair_cap = with_edge_lists ? total_edges * 3 + (size_t(1) << 16) : 0;
    if (with_edge_lists && pair_cap < rt->edge_pair_capacity) pair_cap = rt->edge_pair_capacity;
    if (pair_cap > 0xFFFFFF00u) pair_cap = 0xFFFFFF00u;
    if (const char* e = getenv("B2DGPU_EDGE_LIST_CAPACITY")) if (with_edge_lists) pair_cap = size_t(strtoull(e, nullptr, 10));   // test knob
    const size_t o_cell_edge_cnt = take(with_edge_lists ? sizeof(uint32_t) * size_t(cap) : 0);
    const size_t o_cell_edge_off = take(with_edge_lists ? sizeof(uint32_t) * (size_t(cap) + 1) : 0);
    const size_t o_scratch2 = take(with_edge_lists ? sizeof(uint32_t) * bin_cell_scan_scratch_items(cap) : 0);
    const size_t o_band_edges = take(sizeof(uint32_t) * pair_cap);
    const size_t need = align_up(off, 256);
    if (need > rt->bins.cap) {
      CU_TRY(cudaStreamSynchronize(s));
      CU_TRY(rt->bins.ensure(need + need / 4));
    }
    uint8_t* bp = static_cast<uint8_t*>(rt->bins.ptr);
    BinParams Bn;
    Bn.commands = d_cmds; Bn.command_count = in.command_count;
    Bn.cmd_bbox_px = F.cmd_bbox_px; Bn.cmd_edges = F.cmd_edges; Bn.edges = B.edges;
    Bn.y_begin = t->y0; Bn.tile_h = tile_h; Bn.tiles_y = T.tiles_y;
    Bn.state = reinterpret_cast<uint32_t*>(bp + o_state);
    Bn.cm_count = reinterpret_cast<uint32_t*>(bp + o_cm_count); Bn.cm_base = reinterpret_cast<uint32_t*>(bp + o_cm_base);
    Bn.band_count = reinterpret_cast<uint32_t*>(bp + o_band_count); Bn.band_off = reinterpret_cast<uint32_t*>(bp + o_band_off);
    Bn.band_prefix = reinterpret_cast<uint32_t*>(bp + o_band_prefix);
    Bn.scan_scratch = reinterpret_cast<uint32_t*>(bp + o_scratch);
    Bn.cm_index = reinterpret_cast<uint32_t*>(bp + o_cm_index); Bn.cell_cmd = reinterpret_cast<uint32_t*>(bp + o_cell_cmd);
    Bn.cell_ext = reinterpret_cast<uint2*>(bp + o_cell_ext);
    Bn.capacity = cap;
    Bn.cell_edge_cnt = reinterpret_cast<uint32_t*>(bp + o_cell_edge_cnt);
    Bn.cell_edge_off = reinterpret_cast<uint32_t*>(bp + o_cell_edge_off);
    Bn.scan_scratch2 = reinterpret_cast<uint32_t*>(bp + o_scratch2);
    Bn.band_edges = with_edge_lists ? reinterpret_cast<uint32_t*>(bp + o_band_edges) : nullptr;
    Bn.edge_list_capacity = uint32_t(pair_cap);
    launches += launch_binning(Bn, s);
    CU_TRY(cudaMemcpyAsync(rt->h_bin_state, Bn.state, 16, cudaMemcpyDeviceToHost, s));    // read by a LATER render, never waited for
    if (getenv("B2DGPU_DEBUG_BINS")) {
      uint32_t st[4] = {0, 0, 0, 0};
      cudaStreamSynchronize(s);
      cudaMemcpy(st, Bn.state, 16, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[bins] cmds %u tiles_y %d cap %u need %zu buf %zu pair_cap %zu edges %zu state %u %u %u %u err %s\n", in.command_count, T.tiles_y, cap, need, rt->bins.cap, pair_cap, total_edges, st[0], st[1], st[2], st[3], cudaGetErrorString(cudaGetLastError()));
    }
    T.band_off = Bn.band_off; T.cell_cmd = Bn.cell_cmd; T.cell_ext = Bn.cell_ext; T.bin_state = Bn.state;
    T.cell_edge_off = Bn.cell_edge_off; T.band_edges = Bn.band_edges;
  }
  if (rt->profiling && ti == 0) CU_TRY(cudaEventRecord(ev[1], s));
  bool streamed = false;
  if (in.stream_ok) {
    int box[4] = { in.stream_box[0] < 0 ? 0 : in.stream_box[0], in.stream_box[1] < t->y0 ? t->y0 : in.stream_box[1],
                   in.stream_box[2] > t->w ? t->w : in.stream_box[2], in.stream_box[3] > t->y0 + t->h ? t->y0 + t->h : in.stream_box[3] };
    // Large dirty regions only: small boxes are latency bound either way and the tile path culls them well.
    if (box[0] < box[2] && box[1] < box[3] && (long long)(box[2] - box[0]) * (box[3] - box[1]) >= (1 << 20)) {
      // A8 targets are streamed as words of four pixels: the box has to start on a word and end on one (or at the
      // right edge of the image, where the rest of the word is row padding).
      const bool words_ok = t->bpp == 4 || ((box[0] & 3) == 0 && ((box[2] & 3) == 0 || box[2] == t->w));
      if (in.solid.ok && words_ok) {
        SolidStreamParams S;
        S.dst = t->d_pixels; S.dst_stride = intptr_t(t->stride);
        S.x0w = t->bpp == 4 ? box[0] : box[0] / 4;
        S.x1w = t->bpp == 4 ? box[2] : (box[2] + 3) / 4;
        S.y0 = box[1] - t->y0; S.y1 = box[3] - t->y0;
        S.mode = in.solid.comp_op == 0u ? 0u : (in.solid.alpha == 255u ? 2u : 1u);
        S.src = t->bpp == 4 ? in.solid.prgb32 : (in.solid.prgb32 >> 24) * 0x01010101u;
        S.mask = in.solid.alpha;
        S.pixels = (unsigned long long)(box[2] - box[0]) * (unsigned long long)(box[3] - box[1]);
        S.pixel_counter = rt->count_pixels ? rt->d_pixel_counter : nullptr;
        launches += launch_stream_solid(S, rt->sm_count, s);
      }
      else if (in.solid.one && t->bpp == 4) {
        StreamOneParams S;
        S.dst = t->d_pixels; S.dst_stride = intptr_t(t->stride); S.y_begin = t->y0;
        S.x0 = box[0]; S.y0 = box[1]; S.x1 = box[2]; S.y1 = box[3];
        S.fetch_type = in.solid.fetch_type; S.src_format = in.solid.src_format; S.comp_op = in.solid.comp_op; S.alpha = in.solid.alpha;
        { static const int stage = [] { const char* e = getenv("B2DGPU_STREAM_LUT_SMEM"); return e ? atoi(e) : 1; }();   // experiment knob
          S.stage_lut = stage ? in.solid.lut_entries : 0u; }
        S.fd = T.fetch_data + in.solid.fetch_index;
        S.bayer = rt->d_bayer; S.origin_x = in.origin_x; S.origin_y = in.origin_y;
        S.pixels = (unsigned long long)(box[2] - box[0]) * (unsigned long long)(box[3] - box[1]);
        S.pixel_counter = rt->count_pixels ? rt->d_pixel_counter : nullptr;
        launches += launch_stream_one(S, rt->sm_count, s);
      }
      else launches += launch_box_stream(T, t->bpp, box, rt->sm_count, s);
      streamed = true;
    }
  }
  if (!streamed) launches += launch_tile_render(T, t->bpp, tile_h, s);
  CU_TRY(cudaEventRecord(t->rendered, s));              // b2dgpu_target_wait(): stream-ordered consumers of this target
  t->rendered_valid = true;
  }
  if (rt->profiling) {
    CU_TRY(cudaEventRecord(ev[2], s));
    for (int i = 0; i < 3; i++) rt->prof_events.push_back(ev[i]);
    ev.handed_over = true;
  }

  CU_TRY(cudaGetLastError());
  rt->stats.kernel_launches += uint64_t(launches);
  rt->stats.commands += in.command_count;
  rt->stats.edges += total_edges;
  return B2DGPU_SUCCESS;
}

static b2dgpu_result submit_impl(b2dgpu_runtime* rt, b2dgpu_target* target, const b2dgpu_batch_view* view);

extern "C" b2dgpu_result b2dgpu_submit(b2dgpu_runtime* rt, b2dgpu_target* target, const b2dgpu_batch_view* view_in) {
  if (!rt || rt->magic != kRuntimeMagic || !target || target->rt != rt) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_submit: invalid argument");
  if (!view_in) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_submit: view is null");
  b2dgpu_batch_view full;
  if (!normalize_view(view_in, &full)) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: bad struct_size");
  const b2dgpu_batch_view* view = &full;
  if (view->command_count == 0) return B2DGPU_SUCCESS;
  b2dgpu_result r = submit_impl(rt, target, view);
  if (r) return r;
  // Capture (b2dgpu_capture_begin): keep a device-resident copy of the batch so it can be replayed with its inputs in HBM.
  bool capturing;
  { std::lock_guard<std::mutex> g(g_registry_mutex); capturing = g_capture != nullptr; }
  if (capturing) {
    b2dgpu_batch* copy = nullptr;
    r = b2dgpu_batch_upload(rt, view, &copy);
    if (r) return r;
    std::lock_guard<std::mutex> g(g_registry_mutex);
    if (g_capture) g_capture->entries.push_back(CaptureEntry{ rt, target, copy });
    else b2dgpu_batch_destroy(copy);
  }
  return B2DGPU_SUCCESS;
}

static b2dgpu_result submit_impl(b2dgpu_runtime* rt, b2dgpu_target* target, const b2dgpu_batch_view* view) {
  std::lock_guard<std::mutex> lock(rt->mutex);
  cudaSetDevice(rt->device);

  PreparedBatch pb;
  b2dgpu_result r = prepare_batch(view, pb);
  if (r) return r;

  // Slot = one of two device blocks.  The slot's previous batch (two submits ago) must have finished rendering.
  const int slot = rt->slot_next;
  rt->slot_next ^= 1;
  if (rt->slot_busy[slot]) { CU_TRY(cudaEventSynchronize(rt->slot_done[slot])); rt->slot_busy[slot] = false; }
  CU_TRY(rt->oneshot_block[slot].ensure(pb.lay.total_bytes));
  uint8_t* blk = static_cast<uint8_t*>(rt->oneshot_block[slot].ptr);

  // Upload + K1 count + scan on the prep stream: none of it depends on the batch that may still be rendering on the
  // main stream, so the edge-total readback below only waits for these few small kernels.
  r = upload_block(rt, view, pb, blk, rt->prep_stream);
  if (r) return r;

  RenderInput in;
  in.block = blk; in.lay = &pb.lay; in.edges = &rt->oneshot_edges;
  in.command_count = view->command_count; in.supplied_edges = view->edge_count;
  in.segment_count = view->segment_count + view->generated_segment_count; in.instance_count = view->glyph_instance_count;
  in.origin_x = view->pixel_origin_x; in.origin_y = view->pixel_origin_y;
  in.has_analytic = pb.has_analytic;
  in.built_known = false; in.built_edges = 0; in.edges_staged = false; in.prepped = false;
  in.stream_ok = pb.stream_ok; memcpy(in.stream_box, pb.stream_box, sizeof(in.stream_box));
  in.solid = pb.solid;

  int prep_launches = 0;
  r = prep_block(rt, in, rt->prep_stream, rt->d_scalars + 2 + slot, rt->h_scalars + 2 + slot, prep_launches);
  if (r) return r;
  rt->stats.kernel_launches += uint64_t(prep_launches);
  CU_TRY(cudaEventRecord(rt->prep_ready, rt->prep_stream));
  CU_TRY(cudaStreamWaitEvent(rt->stream, rt->prep_ready, 0));

  r = render_block(rt, &target, 1, in);
  if (r) return r;
  CU_TRY(cudaEventRecord(rt->slot_done[slot], rt->stream));
  rt->slot_busy[slot] = true;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_batch_upload(b2dgpu_runtime* rt, const b2dgpu_batch_view* view_in, b2dgpu_batch** out) {
  if (!rt || rt->magic != kRuntimeMagic || !view_in || !out) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_batch_upload: invalid argument");
  *out = nullptr;
  b2dgpu_batch_view full;
  if (!normalize_view(view_in, &full)) return fail(B2DGPU_ERROR_INVALID_VALUE, "batch view: bad struct_size");
  const b2dgpu_batch_view* view = &full;
  std::lock_guard<std::mutex> lock(rt->mutex);
  cudaSetDevice(rt->device);

  PreparedBatch pb;
  b2dgpu_result r = prepare_batch(view, pb);
  if (r) return r;

  b2dgpu_batch* b = new (std::nothrow) b2dgpu_batch();
  if (!b) return fail(B2DGPU_ERROR_OUT_OF_MEMORY, "b2dgpu_batch_upload: out of host memory");
  b->rt = rt;
  b->lay = pb.lay;
  b->command_count = view->command_count; b->fetch_count = view->fetch_count; b->supplied_edges = view->edge_count;
  b->vertex_count = view->vertex_count; b->segment_count = view->segment_count + view->generated_segment_count; b->state_count = view->geometry_state_count;
  b->instance_count = view->glyph_instance_count;
  b->origin_x = view->pixel_origin_x; b->origin_y = view->pixel_origin_y;
  b->has_analytic = pb.has_analytic;
  b->built_known = false; b->built_edges = 0;

  cudaError_t e = b->block.ensure(pb.lay.total_bytes);
  if (e != cudaSuccess) { delete b; return cuda_fail(e, "b2dgpu_batch_upload: cudaMalloc(block)"); }
  b->edges_staged = false;
  b->stream_ok = pb.stream_ok; memcpy(b->stream_box, pb.stream_box, sizeof(b->stream_box));
  b->solid = pb.solid;
  r = upload_block(rt, view, pb, static_cast<uint8_t*>(b->block.ptr), rt->stream);
  if (r) { b->block.release(); b->edges.release(); delete b; return r; }
  *out = b;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_batch_destroy(b2dgpu_batch* b) {
  if (!b) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_batch_destroy: null");
  cudaSetDevice(b->rt->device);
  cudaStreamSynchronize(b->rt->stream);
  b->block.release();
  b->edges.release();
  delete b;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_batch_render_multi(b2dgpu_runtime* rt, b2dgpu_target* const* targets, uint32_t target_count, b2dgpu_batch* b) {
  if (!rt || rt->magic != kRuntimeMagic || !targets || !target_count || !b || b->rt != rt)
    return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_batch_render: invalid argument");
  for (uint32_t i = 0; i < target_count; i++)
    if (!targets[i] || targets[i]->rt != rt) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_batch_render: invalid target");
  if (!b->command_count) return B2DGPU_SUCCESS;
  std::lock_guard<std::mutex> lock(rt->mutex);
  cudaSetDevice(rt->device);
  RenderInput in;
  in.block = static_cast<uint8_t*>(b->block.ptr); in.lay = &b->lay; in.edges = &b->edges;
  in.command_count = b->command_count; in.supplied_edges = b->supplied_edges; in.segment_count = b->segment_count;
  in.instance_count = b->instance_count;
  in.origin_x = b->origin_x; in.origin_y = b->origin_y;
  in.has_analytic = b->has_analytic;
  in.built_known = b->built_known; in.built_edges = b->built_edges; in.edges_staged = b->edges_staged; in.prepped = false;
  in.stream_ok = b->stream_ok; memcpy(in.stream_box, b->stream_box, sizeof(in.stream_box));
  in.solid = b->solid;
  // The whole path (K1 count, scan, K1 write, finalize, K2+K3) re-runs on every render; only the host read-back of the
  // edge total is skipped after the first time because the geometry of a resident batch cannot change.
  b2dgpu_result r = render_block(rt, targets, target_count, in);
  b->built_known = in.built_known; b->built_edges = in.built_edges; b->edges_staged = in.edges_staged;
  return r;
}

extern "C" b2dgpu_result b2dgpu_batch_render(b2dgpu_runtime* rt, b2dgpu_target* target, b2dgpu_batch* b) {
  return b2dgpu_batch_render_multi(rt, &target, 1, b);
}

// ---------------------------------------------------------------------------------------------------------------
// Process-wide statistics / profiling switch / capture
// ---------------------------------------------------------------------------------------------------------------
extern "C" b2dgpu_result b2dgpu_global_stats(b2dgpu_stats* out, int reset) {
  if (!out) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_global_stats: out is null");
  std::vector<b2dgpu_runtime*> live;
  b2dgpu_stats total;
  {
    std::lock_guard<std::mutex> g(g_registry_mutex);
    live = g_runtimes;
    total = g_retired_stats;
    if (reset) memset(&g_retired_stats, 0, sizeof(g_retired_stats));
  }
  for (b2dgpu_runtime* rt : live) {
    b2dgpu_stats st;
    b2dgpu_result r = b2dgpu_get_stats(rt, &st, reset);
    if (r) return r;
    stats_add(total, st);
  }
  *out = total;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_global_set_profiling(int enabled) {
  std::vector<b2dgpu_runtime*> live;
  { std::lock_guard<std::mutex> g(g_registry_mutex); g_profiling_default = enabled != 0; live = g_runtimes; }
  for (b2dgpu_runtime* rt : live) b2dgpu_set_profiling(rt, enabled);
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_global_set_pixel_counting(int enabled) {
  std::vector<b2dgpu_runtime*> live;
  { std::lock_guard<std::mutex> g(g_registry_mutex); g_count_pixels_default = enabled != 0; live = g_runtimes; }
  for (b2dgpu_runtime* rt : live) b2dgpu_set_pixel_counting(rt, enabled);
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_capture_begin(void) {
  std::lock_guard<std::mutex> g(g_registry_mutex);
  if (g_capture) return fail(B2DGPU_ERROR_INVALID_STATE, "b2dgpu_capture_begin: a capture is already open");
  g_capture = new (std::nothrow) b2dgpu_capture();
  return g_capture ? B2DGPU_SUCCESS : fail(B2DGPU_ERROR_OUT_OF_MEMORY, "b2dgpu_capture_begin: out of host memory");
}

extern "C" b2dgpu_result b2dgpu_capture_end(b2dgpu_capture** out) {
  std::lock_guard<std::mutex> g(g_registry_mutex);
  if (!g_capture || !out) return fail(B2DGPU_ERROR_INVALID_STATE, "b2dgpu_capture_end: no capture is open");
  *out = g_capture;
  g_capture = nullptr;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_capture_info(const b2dgpu_capture* c, uint32_t* batches, uint64_t* commands) {
  if (!c) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_capture_info: null");
  uint64_t n = 0;
  for (const CaptureEntry& e : c->entries) n += e.batch->command_count;
  if (batches) *batches = uint32_t(c->entries.size());
  if (commands) *commands = n;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_capture_replay(b2dgpu_capture* c, uint32_t times, float* ms_out) {
  if (!c) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_capture_replay: null");
  if (ms_out) *ms_out = 0.f;
  if (c->entries.empty() || !times) return B2DGPU_SUCCESS;
  {
    std::lock_guard<std::mutex> g(g_registry_mutex);
    for (const CaptureEntry& e : c->entries)
      if (!registry_has(g_runtimes, e.rt) || !registry_has(g_targets, e.target))
        return fail(B2DGPU_ERROR_INVALID_STATE, "b2dgpu_capture_replay: the context that was captured has been destroyed");
  }
  b2dgpu_runtime* rt0 = c->entries[0].rt;
  cudaSetDevice(rt0->device);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  CU_TRY(cudaEventCreate(&e0));
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return fail(B2DGPU_ERROR_UNKNOWN, "cudaEventCreate"); }
  cudaEventRecord(e0, rt0->stream);
  b2dgpu_result r = B2DGPU_SUCCESS;
  for (uint32_t t = 0; t < times && !r; t++)
    for (const CaptureEntry& e : c->entries) {
      r = b2dgpu_batch_render(e.rt, e.target, e.batch);
      if (r) break;
    }
  cudaEventRecord(e1, rt0->stream);
  for (const CaptureEntry& e : c->entries) cudaStreamSynchronize(e.rt->stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ms_out) *ms_out = ms;
  return r;
}

extern "C" b2dgpu_result b2dgpu_capture_destroy(b2dgpu_capture* c) {
  if (!c) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_capture_destroy: null");
  {
    std::lock_guard<std::mutex> g(g_registry_mutex);
    if (g_capture == c) g_capture = nullptr;
  }
  for (const CaptureEntry& e : c->entries) {
    bool live;
    { std::lock_guard<std::mutex> g(g_registry_mutex); live = registry_has(g_runtimes, e.rt); }
    if (live) b2dgpu_batch_destroy(e.batch);
    else { e.batch->block.release(); e.batch->edges.release(); delete e.batch; }
  }
  delete c;
  return B2DGPU_SUCCESS;
}

extern "C" b2dgpu_result b2dgpu_debug_build_edges(b2dgpu_runtime* rt, const b2dgpu_batch_view* view, b2dgpu_edge* edges_out,
                                                  uint32_t capacity, uint32_t* count_out, uint32_t* per_command_begin_out) {
  if (!rt || rt->magic != kRuntimeMagic || !view || !count_out) return fail(B2DGPU_ERROR_INVALID_VALUE, "b2dgpu_debug_build_edges: invalid argument");
  b2dgpu_batch* b = nullptr;
  b2dgpu_result r = b2dgpu_batch_upload(rt, view, &b);
  if (r) return r;
  std::lock_guard<std::mutex> lock(rt->mutex);
  cudaSetDevice(rt->device);
  uint8_t* blk = static_cast<uint8_t*>(b->block.ptr);
  const BlockLayout& L = b->lay;
  cudaStream_t s = rt->stream;

  BuildParams B;
  memset(&B, 0, sizeof(B));
  B.vertices = reinterpret_cast<const double*>(blk + L.vertices);
  B.segments = reinterpret_cast<const b2dgpu_segment*>(blk + L.segments);
  B.segment_count = b->segment_count;
  B.commands = reinterpret_cast<const b2dgpu_command*>(blk + L.commands);
  B.states = reinterpret_cast<const b2dgpu_geometry_state*>(blk + L.states);
  B.seg_counts = reinterpret_cast<uint32_t*>(blk + L.seg_counts);
  B.seg_offsets = reinterpret_cast<uint32_t*>(blk + L.seg_offsets);
  B.edge_base = 0;
  B.cmd_bbox_fixed = reinterpret_cast<int4*>(blk + L.bbox_fixed);
  B.error_flag = rt->d_scalars + 1;

  cudaError_t e = cudaSuccess;
  uint32_t built = 0;
  std::vector<uint32_t> offs(size_t(b->segment_count) + 1, 0);
  if (b->instance_count) {
    GlyphParams G;
    G.cache = static_cast<const uint32_t*>(rt->glyph_cache.ptr);
    G.instances = reinterpret_cast<const b2dgpu_glyph_instance*>(blk + L.instances);
    G.instance_count = b->instance_count;
    G.vertices = reinterpret_cast<double*>(blk + L.vertices);
    G.segments = reinterpret_cast<b2dgpu_segment*>(blk + L.segments);
    G.error_flag = rt->d_scalars + 1;
    cudaStreamWaitEvent(s, rt->glyph_cache_ready, 0);
    launch_glyph_instances(G, s);
  }
  if (b->segment_count) {
    launch_init_bbox(B.cmd_bbox_fixed, b->command_count, s);
    launch_count_edges(B, s);
    launch_exclusive_scan(B.seg_counts, reinterpret_cast<uint32_t*>(blk + L.seg_offsets), b->segment_count,
                          reinterpret_cast<uint32_t*>(blk + L.scan_scratch), rt->d_scalars, s);
    e = cudaMemcpyAsync(offs.data(), B.seg_offsets, sizeof(uint32_t) * offs.size(), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    built = offs[b->segment_count];
    if (e == cudaSuccess && built) {
      e = b->edges.ensure(sizeof(b2dgpu_edge) * size_t(built));
      if (e == cudaSuccess) {
        B.edges = static_cast<b2dgpu_edge*>(b->edges.ptr);
        B.edge_capacity = uint32_t(b->edges.cap / sizeof(b2dgpu_edge));
        launch_write_edges(B, s);
        if (edges_out) {
          uint32_t n = built < capacity ? built : capacity;
          e = cudaMemcpyAsync(edges_out, B.edges, sizeof(b2dgpu_edge) * n, cudaMemcpyDeviceToHost, s);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      }
    }
  }
  *count_out = built;
  if (per_command_begin_out) {
    // Commands own contiguous segment ranges, so their edge ranges follow from the segment offsets.
    uint32_t last = 0;
    for (uint32_t i = 0; i < view->command_count; i++) {
      const b2dgpu_command& c = view->commands[i];
      if (c.type == B2DGPU_CMD_FILL_GEOMETRY && c.data_count) {
        per_command_begin_out[i] = offs[c.data_offset];
        last = offs[c.data_offset + c.data_count];
      }
      else per_command_begin_out[i] = last;
    }
    per_command_begin_out[view->command_count] = built;
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  b2dgpu_result res = (e == cudaSuccess) ? B2DGPU_SUCCESS : cuda_fail(e, "b2dgpu_debug_build_edges");
  b->block.release(); b->edges.release(); delete b;
  return res;
}
