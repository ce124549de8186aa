// b2dgpu_shim_impl.h - second half of the Blend2D <-> libb2dgpu binding (see b2dgpu_shim_fwd.h).
//
// Included by the overlay INSIDE `namespace bl::RasterEngine` of the build-tree copy of rastercontext.cpp, after the
// asynchronous enqueue helpers (rastercontext.cpp:2410-2520) and before the fill frontends that call into it.
//
// What it does, all of it on the user thread (a GPU context has no worker threads):
//   * create_runtime()   BL_CONTEXT_CREATE_FLAG 0x10000000 -> a PipeRuntime (piperuntime_p.h:39-62) whose get()/test()
//                        forward to b2dgpu_runtime_get/test; DispatchData carries tokens, never host code.
//   * record_path/poly   small paths and polygons, which the reference flattens on the user thread even when it is
//                        asynchronous (rastercontext.cpp:2784-2866), are recorded as path segments instead: the GPU
//                        edge builder (K1) flattens and clips them.
//   * consume_batch()    replaces WorkerProc::process_work_data() (workerproc.cpp:322-352): walks the command queues
//                        and the job queue of a RenderBatch, turns them into one b2dgpu_batch_view and submits it.
//                        Jobs: fill-geometry -> path segments; stroke-geometry -> the reference's own stroker
//                        (core/pathstroke.cpp) runs on the host and its a/b/c output paths become segments
//                        (b reversed, as EdgeSourceReversePathFromStrokeSink does); text -> glyph outlines decoded
//                        by the reference (bl_font_get_glyph_run_outlines) become segments.
//   * sync_to_host()     flush(BL_CONTEXT_FLUSH_SYNC) / end(): device canvas -> BLImage pixels.
//
// Environment (debug aids): B2DGPU_SHIM_CPU_EDGES=1 builds every edge with the reference's EdgeBuilder on the host
// and ships edges instead of segments; B2DGPU_SHIM_PIN=0 disables page-locking of the target image.
#ifndef B2DGPU_SHIM_IMPL_H_INCLUDED
#define B2DGPU_SHIM_IMPL_H_INCLUDED

namespace GpuShim {

// ---------------------------------------------------------------------------------------------------------------
// Per-context state, owned by the PipeRuntime object (destroyed by detach(): the runtime is flagged kIsolated).
// ---------------------------------------------------------------------------------------------------------------
struct DirectGeometry { uint32_t seg_begin, seg_count, state_index; };

struct State {
  b2dgpu_runtime* rt = nullptr;
  b2dgpu_target* target = nullptr;
  void* registered_pixels = nullptr;
  bool device_dirty = false;        // the device canvas is newer than the host image
  bool cpu_edges = false;
  bool pin_images = true;

  // Geometry: filled at call time by record_*() and at flush time by the job pass.
  std::vector<double> vtx;
  std::vector<b2dgpu_segment> segs;
  std::vector<b2dgpu_geometry_state> states;
  std::vector<DirectGeometry> direct;

  // Flush-time arrays.
  std::vector<b2dgpu_command> cmds;
  std::vector<b2dgpu_fetch_data> fetch;
  std::vector<b2dgpu_edge> edges;
  std::vector<const void*> fetch_keys;

  BLPath tmp_path[5];               // stroker scratch: a, b, c, input copy, glyph outlines

  // Glyph cache: TrueType outlines in the format of dev_glyph.cuh, append-only, mirrored on the device by
  // the runtime; a text fill is described by one b2dgpu_glyph_instance per glyph instead of its decoded outline.
  // kind: 0 cached outline, 1 no outline, 2 decode on the host, 3 compound: components [comp_begin, comp_begin + comp_count)
  struct GlyphEntry { uint32_t blob_offset, vertices, segments; uint8_t kind; uint32_t comp_begin, comp_count; };
  struct GlyphComponent { uint32_t glyph_id; BLMatrix2D local; };                   // opentype/otglyf.cpp:505-596
  std::vector<GlyphComponent> glyph_components;
  std::unordered_map<uint64_t, GlyphEntry> glyph_map;
  std::vector<uint32_t> glyph_cache;
  std::vector<b2dgpu_glyph_instance> instances;     // vertex_base / segment_base relative to the generated ranges until submit
  uint32_t gen_vertices = 0, gen_segments = 0;
  bool use_glyph_cache = true;
  bool adaptive_batches = false;
  bool device_luts = true;          // gradient tables interpolated on the device (B2DGPU_SHIM_DEVICE_LUTS=0: on the host)
  std::vector<b2dgpu_lut_request> lut_requests;
  std::vector<b2dgpu_gradient_stop> lut_stops;
  uint64_t glyphs_instanced = 0, glyph_runs_on_host = 0;    // B2DGPU_SHIM_STATS=1 prints them when the context goes away

  void clear_geometry() noexcept { vtx.clear(); segs.clear(); states.clear(); direct.clear(); instances.clear(); gen_vertices = 0; gen_segments = 0; }
};

struct Runtime : public Pipeline::PipeRuntime {
  State* state;
};

static BL_INLINE State* state_of(const BLRasterContextImpl* ctx_impl) noexcept {
  return static_cast<Runtime*>(ctx_impl->pipe_provider.runtime())->state;
}

// ---------------------------------------------------------------------------------------------------------------
// Seam B: PipeRuntime
// ---------------------------------------------------------------------------------------------------------------
static BLResult BL_CDECL runtime_lookup(Pipeline::PipeRuntime* self, uint32_t signature, Pipeline::DispatchData* out, Pipeline::PipeLookupCache* cache, bool is_test) noexcept {
  State* st = static_cast<Runtime*>(self)->state;
  b2dgpu_dispatch_data dd;
  b2dgpu_result r = is_test ? b2dgpu_runtime_test(st->rt, signature, &dd, nullptr) : b2dgpu_runtime_get(st->rt, signature, &dd, nullptr);
  if (r != B2DGPU_SUCCESS)
    return bl_make_error(BLResult(r));            // BL_ERROR_NOT_IMPLEMENTED / BL_ERROR_NO_ENTRY, like fixedpiperuntime.cpp:314
  out->init(Pipeline::FillFunc(dd.fill_func), Pipeline::FetchFunc(dd.fetch_func));
  if (cache)
    cache->store(signature, out);                 // fixedpiperuntime.cpp:319-320
  return BL_SUCCESS;
}

static BLResult BL_CDECL runtime_test(Pipeline::PipeRuntime* self, uint32_t signature, Pipeline::DispatchData* out, Pipeline::PipeLookupCache* cache) noexcept {
  return runtime_lookup(self, signature, out, cache, true);
}

static BLResult BL_CDECL runtime_get(Pipeline::PipeRuntime* self, uint32_t signature, Pipeline::DispatchData* out, Pipeline::PipeLookupCache* cache) noexcept {
  return runtime_lookup(self, signature, out, cache, false);
}

static void BL_CDECL runtime_destroy(Pipeline::PipeRuntime* self) noexcept {
  Runtime* rt = static_cast<Runtime*>(self);
  State* st = rt->state;
  if (st) {
    if (const char* e = getenv("B2DGPU_SHIM_STATS"))
      if (e[0] == '1')
        fprintf(stderr, "b2dgpu shim: %llu glyphs instanced from a cache of %zu glyphs (%zu KiB), %llu glyph runs decoded on the host\n",
                (unsigned long long)st->glyphs_instanced, st->glyph_map.size(), st->glyph_cache.size() / 256, (unsigned long long)st->glyph_runs_on_host);
    if (st->rt) b2dgpu_sync(st->rt);
    if (st->registered_pixels) b2dgpu_host_unregister(st->rt, st->registered_pixels);
    if (st->target) b2dgpu_target_destroy(st->target);
    if (st->rt) b2dgpu_runtime_destroy(st->rt);
    delete st;
  }
  delete rt;
}

static constexpr uint32_t kCreateFlagAdaptiveBatches = 0x20000000u;    // private: set by adjust_create_info() on its own copy
static constexpr uint32_t kFirstBatchDefault = 512, kLargestBatch = 8192;
// B2DGPU_SHIM_FIRST_BATCH: experiment knob (commands in the first batch of a frame).
static uint32_t first_batch() noexcept {
  static const uint32_t n = [] { const char* e = getenv("B2DGPU_SHIM_FIRST_BATCH"); const int v = e ? atoi(e) : 0; return uint32_t(v >= 16 && v <= int(kLargestBatch) ? v : int(kFirstBatchDefault)); }();
  return n;
}

static const BLContextCreateInfo* adjust_create_info(const BLContextCreateInfo* options, BLContextCreateInfo* storage) noexcept {
  if (!(options->flags & kCreateFlagGpuRuntime))
    return options;
  *storage = *options;
  storage->thread_count = 1;                                        // workermanager.cpp:36-42: the user thread is the worker
  storage->flags &= ~uint32_t(BL_CONTEXT_CREATE_FLAG_FALLBACK_TO_SYNC);
  // b2dgpu_submit() is asynchronous and double buffered: shorter batches let the device start while the frontend is
  // still recording (the reference's default is 10240 commands, rastercontext_p.h:62).
  if (!storage->command_queue_limit) {
    // Adaptive (B2DGPU_SHIM_ADAPTIVE=0: fixed 2048): the first batch of a frame is small so that the device starts early,
    // every further batch is four times larger - the frontend records faster than the device composites, so bigger batches only
    // mean fewer per-batch kernels; consume_batch() and sync_to_host() move the limit.
    const char* e = getenv("B2DGPU_SHIM_ADAPTIVE");
    if (e && e[0] == '0') storage->command_queue_limit = 2048;
    else {
      storage->command_queue_limit = first_batch();
      storage->flags |= kCreateFlagAdaptiveBatches;
    }
  }
  return storage;
}

static BLResult create_runtime(const BLContextCreateInfo* options, Pipeline::PipeRuntime** runtime) noexcept {
  if (*runtime && bl_test_flag((*runtime)->runtime_flags(), Pipeline::PipeRuntimeFlags::kIsolated))
    (*runtime)->destroy();
  *runtime = nullptr;

  Runtime* rt = new (std::nothrow) Runtime();
  State* st = new (std::nothrow) State();
  if (!rt || !st) {
    delete rt; delete st;
    return bl_make_error(BL_ERROR_OUT_OF_MEMORY);
  }

  b2dgpu_create_info ci;
  ci.struct_size = sizeof(ci);
  ci.device = int32_t(options->reserved[0]);
  ci.stream = nullptr;
  ci.flags = 0;
  b2dgpu_result r = b2dgpu_runtime_create(&ci, &st->rt);           // no CPU fallback: fails without a CUDA device
  if (r != B2DGPU_SUCCESS) {
    delete rt; delete st;
    return bl_make_error(BLResult(r));
  }

  const char* e = getenv("B2DGPU_SHIM_CPU_EDGES");
  st->cpu_edges = e && e[0] == '1';
  e = getenv("B2DGPU_SHIM_PIN");
  st->pin_images = !(e && e[0] == '0');
  e = getenv("B2DGPU_SHIM_DEVICE_LUTS");
  st->device_luts = !(e && e[0] == '0');
  st->adaptive_batches = (options->flags & kCreateFlagAdaptiveBatches) != 0;
  e = getenv("B2DGPU_SHIM_GLYPH_CACHE");
  st->use_glyph_cache = !(e && e[0] == '0');

  rt->_runtime_type = Pipeline::PipeRuntimeType(kPipeRuntimeTypeGpu);
  rt->_runtime_flags = Pipeline::PipeRuntimeFlags::kIsolated;
  rt->_runtime_size = uint16_t(sizeof(Runtime));
  rt->_destroy = runtime_destroy;
  rt->_funcs.test = runtime_test;
  rt->_funcs.get = runtime_get;
  rt->state = st;
  *runtime = rt;
  return BL_SUCCESS;
}

static bool defer_gradient_table(BLRasterContextImpl* ctx_impl, RenderFetchData* fetch_data) noexcept {
  State* st = state_of(ctx_impl);
  if (!st->device_luts || !fetch_data->signature.is_gradient()) return false;
  if (BLGradientQuality(fetch_data->extra.custom[0]) >= BL_GRADIENT_QUALITY_DITHER) return false;   // 64-bit tables: host
  // what compute_pending_fetch_data() (raster/renderfetchdata.cpp:14-35) does, minus ensure_lut32()
  fetch_data->signature.clear_pending_bit();
  fetch_data->pipeline_data.gradient.lut.data = nullptr;
  return true;
}

// The device canvas is created on first use from ctx_impl->dst_data (known only at the end of attach()) and
// initialised with the image's current pixels.
static BLResult ensure_target(BLRasterContextImpl* ctx_impl, State* st) noexcept {
  if (st->target)
    return BL_SUCCESS;
  const BLImageData& d = ctx_impl->dst_data;
  b2dgpu_result r = b2dgpu_target_create(st->rt, d.size.w, d.size.h, d.format, &st->target);
  if (r != B2DGPU_SUCCESS)
    return bl_make_error(BLResult(r));
  const size_t bytes = size_t(d.stride) * size_t(d.size.h);
  if (st->pin_images && d.stride > 0 && bytes >= (size_t(1) << 20)) {
    if (b2dgpu_host_register(st->rt, d.pixel_data, bytes) == B2DGPU_SUCCESS)
      st->registered_pixels = d.pixel_data;
  }
  b2dgpu_image_data img = { d.pixel_data, d.stride, d.size.w, d.size.h, d.format, 0 };
  r = b2dgpu_target_upload(st->target, &img);
  return r == B2DGPU_SUCCESS ? BL_SUCCESS : bl_make_error(BLResult(r));
}

static BLResult sync_to_host(BLRasterContextImpl* ctx_impl) noexcept {
  State* st = state_of(ctx_impl);
  if (!st->target || !st->device_dirty)
    return BL_SUCCESS;
  const BLImageData& d = ctx_impl->dst_data;
  b2dgpu_image_data img = { d.pixel_data, d.stride, d.size.w, d.size.h, d.format, 0 };
  b2dgpu_result r = b2dgpu_target_download(st->target, &img);      // stream ordered, returns when the pixels are there
  if (r != B2DGPU_SUCCESS)
    return ctx_impl->accumulate_error(bl_make_error(BLResult(r)));
  st->device_dirty = false;
  if (st->adaptive_batches) ctx_impl->worker_mgr()._command_queue_limit = first_batch();     // a new frame starts small again
  return BL_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------
// Geometry recording: BLPathView / polygons -> b2dgpu_segment[] + vertices (the K1 input format).
// ---------------------------------------------------------------------------------------------------------------
static BL_INLINE void push_segment(State* st, uint32_t p0, uint32_t p1, uint32_t kind) noexcept {
  b2dgpu_segment s;
  s.p0 = p0;
  s.p1_kind = (p1 << 2) | kind;
  s.command = 0;                                  // patched in consume_batch()
  st->segs.push_back(s);
}

static uint32_t add_state(State* st, const BLMatrix2D& m, BLTransformType transform_type, const BLBox& clip_fixed, double tolerance_fixed) noexcept {
  b2dgpu_geometry_state gs;
  memset(&gs, 0, sizeof(gs));
  gs.m[0] = m.m00; gs.m[1] = m.m01; gs.m[2] = m.m10; gs.m[3] = m.m11; gs.m[4] = m.m20; gs.m[5] = m.m21;
  gs.clip[0] = clip_fixed.x0; gs.clip[1] = clip_fixed.y0; gs.clip[2] = clip_fixed.x1; gs.clip[3] = clip_fixed.y1;
  gs.tolerance_sq = Math::square(tolerance_fixed);
  gs.transform_type = uint32_t(transform_type);
  if (!st->states.empty() && memcmp(&st->states.back(), &gs, sizeof(gs)) == 0)
    return uint32_t(st->states.size() - 1);
  st->states.push_back(gs);
  return uint32_t(st->states.size() - 1);
}

// EdgeSourcePath + EdgeBuilder::add_from_source (edgebuilder_p.h:165-281, 1032-1070): a figure starts at a MOVE, curve
// commands need all of their vertices, and a figure is closed when `closed` or when a CLOSE command follows.
static void append_path(State* st, const BLPathView& view, bool closed) noexcept {
  const uint8_t* cmd = view.command_data;
  const size_t n = view.size;
  if (!n) return;
  const uint32_t base = uint32_t(st->vtx.size() / 2);
  const double* src = reinterpret_cast<const double*>(view.vertex_data);
  st->vtx.insert(st->vtx.end(), src, src + n * 2);

  size_t i = 0;
  while (i < n) {
    if (cmd[i] != BL_PATH_CMD_MOVE) { i++; continue; }
    const size_t start = i;
    size_t cur = i;
    i++;
    for (;;) {
      if (i < n && cmd[i] == BL_PATH_CMD_ON) { push_segment(st, base + uint32_t(cur), base + uint32_t(i), B2DGPU_SEG_LINE); cur = i; i++; }
      else if (i + 2 <= n && cmd[i] == BL_PATH_CMD_QUAD) { push_segment(st, base + uint32_t(cur), base + uint32_t(i), B2DGPU_SEG_QUAD); cur = i + 1; i += 2; }
      else if (i + 2 < n && cmd[i] == BL_PATH_CMD_CUBIC) { push_segment(st, base + uint32_t(cur), base + uint32_t(i), B2DGPU_SEG_CUBIC); cur = i + 2; i += 3; }
      else if (i + 2 < n && cmd[i] == BL_PATH_CMD_CONIC) { push_segment(st, base + uint32_t(cur), base + uint32_t(i), B2DGPU_SEG_CONIC); cur = i + 2; i += 2; }
      else {
        if (closed || (i < n && cmd[i] == BL_PATH_CMD_CLOSE))
          push_segment(st, base + uint32_t(cur), base + uint32_t(start), B2DGPU_SEG_LINE);
        break;
      }
    }
  }
}

// EdgeSourceReversePathFromStrokeSink (edgebuilder_p.h:286-406): the stroker's `b` path, walked from its end.  The
// vertices are stored reversed, so a reversed curve is an ordinary ascending segment of the reversed array.
static void append_reverse_path_from_stroke_sink(State* st, const BLPathView& view) noexcept {
  const uint8_t* cmd = view.command_data;
  size_t n = view.size;
  const bool must_close = n > 0 && cmd[n - 1] == BL_PATH_CMD_CLOSE;
  n -= size_t(must_close);
  if (!n || cmd[n - 1] != BL_PATH_CMD_ON)
    return;

  const uint32_t base = uint32_t(st->vtx.size() / 2);
  for (size_t k = 0; k < n; k++) {                // reversed[k] = vertex[n - 1 - k]
    st->vtx.push_back(view.vertex_data[n - 1 - k].x);
    st->vtx.push_back(view.vertex_data[n - 1 - k].y);
  }
  // Position p in the original array corresponds to n - 1 - p in the reversed one; `pos` is the original index of
  // the current point (the source's _cmd_ptr).
  size_t pos = n - 1;
  const uint32_t start = base;
  uint32_t cur = base;
  while (pos != 0) {
    const uint8_t c = cmd[pos - 1];
    const uint32_t r = base + uint32_t(n - pos);  // reversed index of original vertex pos - 1
    if (c <= BL_PATH_CMD_ON) { push_segment(st, cur, r, B2DGPU_SEG_LINE); cur = r; pos -= 1; }
    else if (c == BL_PATH_CMD_QUAD && pos >= 2) { push_segment(st, cur, r, B2DGPU_SEG_QUAD); cur = r + 1; pos -= 2; }
    else if (c == BL_PATH_CMD_CUBIC && pos >= 3) { push_segment(st, cur, r, B2DGPU_SEG_CUBIC); cur = r + 2; pos -= 3; }
    else break;                                   // conics never come out of the stroker; the reference's source stops too
  }
  if (must_close)
    push_segment(st, cur, start, B2DGPU_SEG_LINE);
}

template<typename PointType>
static void append_poly(State* st, const PointType* pts, size_t n) noexcept {
  const uint32_t base = uint32_t(st->vtx.size() / 2);
  for (size_t i = 0; i < n; i++) {
    st->vtx.push_back(double(pts[i].x));
    st->vtx.push_back(double(pts[i].y));
  }
  for (size_t i = 1; i < n; i++)
    push_segment(st, base + uint32_t(i - 1), base + uint32_t(i), B2DGPU_SEG_LINE);
  push_segment(st, base + uint32_t(n - 1), base, B2DGPU_SEG_LINE);
}

// Tagged pointer stored in RenderCommand::_payload.analytic.edges for commands whose geometry was recorded at call
// time: (index into State::direct << 1) | 1.  Real EdgeVector pointers are 8-byte aligned.
static BL_INLINE const EdgeVector<int>* direct_tag(size_t index) noexcept { return reinterpret_cast<const EdgeVector<int>*>((uintptr_t(index) << 1) | 1u); }
static BL_INLINE bool is_direct_tag(const EdgeVector<int>* p) noexcept { return (uintptr_t(p) & 1u) != 0; }
static BL_INLINE size_t direct_index(const EdgeVector<int>* p) noexcept { return size_t(uintptr_t(p) >> 1); }

// Enqueues a FillAnalytic command whose geometry is segments [seg_begin, end) recorded just now.
static BLResult enqueue_direct(BLRasterContextImpl* ctx_impl, DispatchInfo di, DispatchStyle ds, BLFillRule fill_rule,
                               size_t seg_begin, size_t vtx_begin, const BLMatrix2D& transform, BLTransformType transform_type) noexcept {
  State* st = state_of(ctx_impl);
  if (st->segs.size() == seg_begin) {
    st->vtx.resize(vtx_begin);
    return BL_SUCCESS;
  }

  RenderCommand* command = ctx_impl->worker_mgr->current_command();
  di.add_fill_type(Pipeline::FillType::kAnalytic);
  command->init_command(di.alpha);
  command->init_fill_analytic(const_cast<EdgeVector<int>*>(direct_tag(st->direct.size())), 0, fill_rule);

  BLResult result = ensure_fetch_and_dispatch_data(ctx_impl, di.signature, ds.fetch_data, command->pipe_dispatch_data());
  if (BL_UNLIKELY(result != BL_SUCCESS)) {
    st->segs.resize(seg_begin);
    st->vtx.resize(vtx_begin);
    return result;
  }

  DirectGeometry dg;
  dg.seg_begin = uint32_t(seg_begin);
  dg.seg_count = uint32_t(st->segs.size() - seg_begin);
  dg.state_index = add_state(st, transform, transform_type, ctx_impl->final_clip_box_fixed_d(), ctx_impl->internal_state.toleranceFixedD);
  st->direct.push_back(dg);

  return enqueue_command(ctx_impl, command, kInvalidQuantizedCoordinate, ds.fetch_data, [&](RenderCommand* command) noexcept {
    command->_payload.analytic.state_slot_index = ctx_impl->worker_mgr().next_state_slot_index();
  });
}

//! True when paths / polygons are recorded as segments for the GPU edge builder instead of being flattened here.
static BL_INLINE bool records_geometry(const BLRasterContextImpl* ctx_impl) noexcept {
  return is_gpu(ctx_impl) && !state_of(ctx_impl)->cpu_edges;
}

//! fill_unclipped_path<kAsync>() on a GPU context (rastercontext.cpp:2787-2797).
static BLResult record_path(BLRasterContextImpl* ctx_impl, DispatchInfo di, DispatchStyle ds, const BLPathView& view,
                            BLFillRule fill_rule, const BLMatrix2D& transform, BLTransformType transform_type) noexcept {
  State* st = state_of(ctx_impl);
  const size_t seg_begin = st->segs.size(), vtx_begin = st->vtx.size();
  append_path(st, view, true);
  return enqueue_direct(ctx_impl, di, ds, fill_rule, seg_begin, vtx_begin, transform, transform_type);
}

//! fill_unclipped_polygon_t<kAsync>() on a GPU context (rastercontext.cpp:2848-2858).
template<typename PointType>
static BLResult record_poly(BLRasterContextImpl* ctx_impl, DispatchInfo di, DispatchStyle ds, const PointType* pts, size_t size,
                            BLFillRule fill_rule, const BLMatrix2D& transform, BLTransformType transform_type) noexcept {
  State* st = state_of(ctx_impl);
  const size_t seg_begin = st->segs.size(), vtx_begin = st->vtx.size();
  if (size)
    append_poly(st, pts, size);
  return enqueue_direct(ctx_impl, di, ds, fill_rule, seg_begin, vtx_begin, transform, transform_type);
}

// ---------------------------------------------------------------------------------------------------------------
// Job pass (renderjobproc_p.h:114-270 with the EdgeBuilder replaced by segment recording)
// ---------------------------------------------------------------------------------------------------------------
struct SegmentSink {
  State* st;
  // Stroke only:
  BLPath* paths;
  const BLStrokeOptions* stroke_options;
  const BLApproximationOptions* approximation_options;
};

static BLResult BL_CDECL fill_glyph_run_segment_sink(BLPathCore* path, const void* info, void* user_data) noexcept {
  bl_unused(info);
  SegmentSink* sink = static_cast<SegmentSink*>(user_data);
  append_path(sink->st, path->dcast().view(), true);                // fill_glyph_run_sink, rastercontextops.cpp:51-59
  return path->dcast().clear();
}

static BLResult BL_CDECL stroke_geometry_segment_sink(BLPathCore* a, BLPathCore* b, BLPathCore* c, size_t figure_start, size_t figure_end, void* user_data) noexcept {
  bl_unused(figure_start, figure_end);
  SegmentSink* sink = static_cast<SegmentSink*>(user_data);
  append_path(sink->st, a->dcast().view(), false);                  // stroke_geometry_sink, rastercontextops.cpp:61-74
  append_reverse_path_from_stroke_sink(sink->st, b->dcast().view());
  if (!c->dcast().is_empty())
    append_path(sink->st, c->dcast().view(), false);
  return a->dcast().clear();
}

static BLResult BL_CDECL stroke_glyph_run_segment_sink(BLPathCore* path, const void* info, void* user_data) noexcept {
  bl_unused(info);
  SegmentSink* sink = static_cast<SegmentSink*>(user_data);
  BLPath& a = sink->paths[0];
  BLPath& b = sink->paths[1];
  BLPath& c = sink->paths[2];
  a.clear();
  BLResult result = PathInternal::stroke_path(path->dcast().view(), *sink->stroke_options, *sink->approximation_options,
                                              a, b, c, stroke_geometry_segment_sink, sink);
  bl_path_clear(path);                                              // stroke_glyph_run_sink, rastercontextops.cpp:76-96
  return result;
}

// ---------------------------------------------------------------------------------------------------------------
// Glyph cache
// ---------------------------------------------------------------------------------------------------------------
struct GlyphPathPut {
  const BLPoint* want;          // the reference decoder's vertices
  size_t size;
  bool ok;
  void operator()(uint32_t index, double x, double y) noexcept {
    if (index >= size || memcmp(&want[index].x, &x, 8) != 0 || memcmp(&want[index].y, &y, 8) != 0) ok = false;
  }
};

static State::GlyphEntry build_glyph_entry(State* st, const BLFontFacePrivateImpl* face_impl, BLGlyphId glyph_id) noexcept;

static const State::GlyphEntry& glyph_entry(State* st, const BLFontFacePrivateImpl* face_impl, BLGlyphId glyph_id) noexcept {
  const uint64_t key = (uint64_t(face_impl->unique_id) << 32) | glyph_id;
  auto it = st->glyph_map.find(key);
  if (it == st->glyph_map.end()) {
    const State::GlyphEntry e = build_glyph_entry(st, face_impl, glyph_id);       // may insert other glyphs (components)
    it = st->glyph_map.emplace(key, e).first;
  }
  return it->second;
}

// Every vertex a cached glyph (simple or compound) produces under `m`, in the reference's order, through `sink(x, y)`;
// CLOSE slots are reported as NaN.  Returns false when some component is not cached.
template<typename Sink>
static bool replay_glyph(State* st, const BLFontFacePrivateImpl* face_impl, BLGlyphId glyph_id, const BLMatrix2D& m, Sink& sink, int level) noexcept {
  const State::GlyphEntry e = glyph_entry(st, face_impl, glyph_id);
  if (e.kind == 1) return true;
  if (e.kind == 0) {
    struct Put {
      std::vector<BLPoint>* v; size_t base;
      void operator()(uint32_t index, double x, double y) noexcept { (*v)[base + index] = BLPoint(x, y); }
    };
    std::vector<BLPoint> tmp(e.vertices, BLPoint(Math::nan<double>(), Math::nan<double>()));
    Put put = { &tmp, 0 };
    const double mm[6] = { m.m00, m.m01, m.m10, m.m11, m.m20, m.m21 };
    if (b2d::glyph_emit(b2d::glyph_blob_view(st->glyph_cache.data() + e.blob_offset), mm, put) != e.vertices) return false;
    for (const BLPoint& pt : tmp) sink(pt.x, pt.y);
    return true;
  }
  if (e.kind != 3 || level >= 15) return false;
  for (uint32_t k = 0; k < e.comp_count; k++) {
    const State::GlyphComponent comp = st->glyph_components[e.comp_begin + k];
    BLMatrix2D cm = comp.local;
    TransformInternal::multiply(cm, cm, m);                                      // otglyf.cpp:601
    if (!replay_glyph(st, face_impl, comp.glyph_id, cm, sink, level + 1)) return false;
  }
  return true;
}

// A compound glyph: the list of its components with their local matrices (opentype/otglyf.cpp:505-601, restated: flags,
// glyph id, byte / word arguments, F2Dot14 scale / scale-xy / 2x2, FreeType-style scaled offsets).  An instance of it is
// expanded on the host into instances of the components with the composed matrices the reference would use; like a
// simple entry it is only kept if that reproduces the reference's decoder bit for bit under two matrices.
static State::GlyphEntry build_compound_entry(State* st, const BLFontFacePrivateImpl* face_impl, BLGlyphId glyph_id, const uint8_t* p, const uint8_t* end) noexcept {
  State::GlyphEntry host = { 0, 0, 0, 2, 0, 0 };
  const OpenType::OTFaceImpl* ot = static_cast<const OpenType::OTFaceImpl*>(face_impl);
  std::vector<State::GlyphComponent> comps;
  for (;;) {
    if (end - p < 6) return host;
    const uint32_t flags = MemOps::readU16uBE(p);
    const uint32_t comp_glyph = MemOps::readU16uBE(p + 2);
    if (comp_glyph >= ot->face_info.glyph_count || comp_glyph == glyph_id) return host;
    int arg1 = int(int8_t(p[4])), arg2 = int(int8_t(p[5]));
    p += 6;
    if (flags & 0x0001u) {                                                          // kArgsAreWords
      if (end - p < 2) return host;
      arg1 = int(uint32_t(arg1) << 8) | (arg2 & 0xFF);
      arg2 = int(int16_t(MemOps::readU16uBE(p)));
      p += 2;
    }
    if (!(flags & 0x0002u)) { arg1 &= 0xFFFF; arg2 &= 0xFFFF; }                     // kArgsAreXYValues
    const double kScaleF2x14 = 1.0 / 16384.0;
    BLMatrix2D cm(1.0, 0.0, 0.0, 1.0, double(arg1), double(arg2));
    if (flags & (0x0008u | 0x0040u | 0x0080u)) {                                    // kAnyCompoundScale
      if (flags & 0x0008u) {
        if (end - p < 2) return host;
        const double scale = double(int16_t(MemOps::readU16uBE(p))) * kScaleF2x14;
        cm.m00 = scale; cm.m11 = scale; p += 2;
      }
      else if (flags & 0x0040u) {
        if (end - p < 4) return host;
        cm.m00 = double(int16_t(MemOps::readU16uBE(p))) * kScaleF2x14;
        cm.m11 = double(int16_t(MemOps::readU16uBE(p + 2))) * kScaleF2x14; p += 4;
      }
      else {
        if (end - p < 8) return host;
        cm.m00 = double(int16_t(MemOps::readU16uBE(p))) * kScaleF2x14;
        cm.m01 = double(int16_t(MemOps::readU16uBE(p + 2))) * kScaleF2x14;
        cm.m10 = double(int16_t(MemOps::readU16uBE(p + 4))) * kScaleF2x14;
        cm.m11 = double(int16_t(MemOps::readU16uBE(p + 6))) * kScaleF2x14; p += 8;
      }
      if ((flags & (0x0002u | 0x0800u | 0x1000u)) == (0x0002u | 0x0800u)) {          // scaled component offset
        cm.m20 *= Geometry::magnitude(BLPoint(cm.m00, cm.m01));
        cm.m21 *= Geometry::magnitude(BLPoint(cm.m10, cm.m11));
      }
    }
    comps.push_back(State::GlyphComponent{ comp_glyph, cm });
    if (comps.size() > 64) return host;
    if (!(flags & 0x0020u)) break;                                                  // kMoreComponents
  }

  // register (tentatively) so that replay_glyph() can walk it, verify against the reference, roll back on a mismatch
  State::GlyphEntry e = { 0, 0, 0, 3, uint32_t(st->glyph_components.size()), uint32_t(comps.size()) };
  st->glyph_components.insert(st->glyph_components.end(), comps.begin(), comps.end());
  const uint64_t key = (uint64_t(face_impl->unique_id) << 32) | glyph_id;
  st->glyph_map[key] = e;
  bool ok = true;
  const BLMatrix2D tests[2] = { BLMatrix2D(1.0, 0.0, 0.0, 1.0, 0.0, 0.0), BLMatrix2D(0.37109375, -0.113, 0.2291, 0.90625, 5.53, -3.2517) };
  for (const BLMatrix2D& m : tests) {
    BLPath path;
    size_t contour_count = 0;
    ScopedBufferTmp<BL_FONT_GET_GLYPH_OUTLINE_BUFFER_SIZE> tmp_buffer;
    if (face_impl->funcs.get_glyph_outlines(face_impl, glyph_id, &m, &path, &contour_count, &tmp_buffer) != BL_SUCCESS) { ok = false; break; }
    struct Check {
      const BLPoint* want; size_t size, at; bool ok;
      void operator()(double x, double y) noexcept {
        if (at >= size) { ok = false; return; }
        const bool close_slot = x != x;
        if (close_slot ? !(want[at].x != want[at].x) : (memcmp(&want[at].x, &x, 8) != 0 || memcmp(&want[at].y, &y, 8) != 0)) ok = false;
        at++;
      }
    } check = { path.vertex_data(), path.size(), 0, true };
    if (!replay_glyph(st, face_impl, glyph_id, m, check, 0) || !check.ok || check.at != path.size()) { ok = false; break; }
  }
  if (!ok) { st->glyph_map[key] = host; return host; }
  return e;
}

// Builds the cache entry of one glyph: parses its `glyf` record (TrueType "Simple Glyph Description": flags with
// repeats, byte / word coordinate deltas), takes the path STRUCTURE from the reference's own decoder and keeps the entry
// only if replaying the deltas (b2d::glyph_emit) reproduces the reference's vertices bit for bit under the identity AND
// under a general matrix.  Anything else - compound glyphs, empty contours, malformed data - is decoded on the host.
static State::GlyphEntry build_glyph_entry(State* st, const BLFontFacePrivateImpl* face_impl, BLGlyphId glyph_id) noexcept {
  State::GlyphEntry host = { 0, 0, 0, 2, 0, 0 };
  const OpenType::OTFaceImpl* ot = static_cast<const OpenType::OTFaceImpl*>(face_impl);
  if (glyph_id >= ot->face_info.glyph_count) return host;
  const OpenType::RawTable glyf = ot->glyf.glyf_table, loca = ot->glyf.loca_table;
  size_t offset, end_off;
  if (ot->loca_offset_size() == 2) {
    const size_t index = size_t(glyph_id) * 2u;
    if (index + 4u > loca.size) return host;
    offset = size_t(MemOps::readU16uBE(loca.data + index)) * 2u;
    end_off = size_t(MemOps::readU16uBE(loca.data + index + 2)) * 2u;
  }
  else {
    const size_t index = size_t(glyph_id) * 4u;
    if (index + 8u > loca.size) return host;
    offset = MemOps::readU32uBE(loca.data + index);
    end_off = MemOps::readU32uBE(loca.data + index + 4);
  }
  if (offset == end_off && end_off <= glyf.size) return State::GlyphEntry{ 0, 0, 0, 1, 0, 0 };      // no outline (space)
  if (offset >= end_off || end_off > glyf.size || end_off - offset < 12u) return host;

  const uint8_t* p = glyf.data + offset;
  const uint8_t* end = glyf.data + end_off;
  const int contours = int(int16_t(MemOps::readU16uBE(p)));
  if (contours == -1) return build_compound_entry(st, face_impl, glyph_id, p + 10, end);
  if (contours <= 0 || contours > 4096) return host;
  p += 10;
  if (size_t(end - p) < size_t(contours) * 2u + 2u) return host;
  std::vector<uint32_t> ends(size_t(contours), 0u);
  for (int c = 0; c < contours; c++) ends[size_t(c)] = MemOps::readU16uBE(p + c * 2);
  p += size_t(contours) * 2u;
  const size_t instructions = MemOps::readU16uBE(p);
  p += 2;
  if (size_t(end - p) < instructions) return host;
  p += instructions;
  const size_t n = size_t(ends.back()) + 1u;
  if (n < 2 || n > 0xFFFFu) return host;

  std::vector<uint8_t> flags(n);
  for (size_t i = 0; i < n; ) {
    if (p == end) return host;
    const uint8_t f = *p++;
    flags[i++] = f;
    if (f & 0x08u) {                                                                            // kRepeatFlag
      if (p == end) return host;
      size_t r = *p++;
      if (r > n - i) return host;
      while (r--) flags[i++] = f;
    }
  }
  std::vector<int> dx(n, 0), dy(n, 0);
  for (size_t i = 0; i < n; i++) {
    const uint8_t f = flags[i];
    if (f & 0x02u) { if (p == end) return host; int v = *p++; dx[i] = (f & 0x10u) ? v : -v; }
    else if (!(f & 0x10u)) { if (end - p < 2) return host; dx[i] = int(int16_t(MemOps::readU16uBE(p))); p += 2; }
  }
  for (size_t i = 0; i < n; i++) {
    const uint8_t f = flags[i];
    if (f & 0x04u) { if (p == end) return host; int v = *p++; dy[i] = (f & 0x20u) ? v : -v; }
    else if (!(f & 0x20u)) { if (end - p < 2) return host; dy[i] = int(int16_t(MemOps::readU16uBE(p))); p += 2; }
  }

  // The path structure: what the reference's decoder (the variant this process dispatches to) emits.
  BLPath path;
  size_t contour_count = 0;
  ScopedBufferTmp<BL_FONT_GET_GLYPH_OUTLINE_BUFFER_SIZE> tmp_buffer;
  const BLMatrix2D identity(1.0, 0.0, 0.0, 1.0, 0.0, 0.0);
  if (face_impl->funcs.get_glyph_outlines(face_impl, glyph_id, &identity, &path, &contour_count, &tmp_buffer) != BL_SUCCESS) return host;
  const size_t nv = path.size();
  if (!nv || nv > 0xFFFFu) return host;

  // segments relative to the glyph, exactly as append_path() would record them for the decoded path
  State scratch;
  append_path(&scratch, path.view(), true);
  const size_t ns = scratch.segs.size();

  std::vector<uint32_t> blob;
  blob.push_back(uint32_t(n) | (uint32_t(contours) << 16));
  blob.push_back(uint32_t(nv));
  blob.push_back(uint32_t(ns));
  for (int c = 0; c < contours; c += 2)
    blob.push_back(ends[size_t(c)] | (c + 1 < contours ? ends[size_t(c) + 1] << 16 : 0u));
  for (size_t i = 0; i < n; i++) blob.push_back(uint32_t(uint16_t(int16_t(dx[i]))) | (uint32_t(uint16_t(int16_t(dy[i]))) << 16));
  for (size_t i = 0; i < n; i += 32) {
    uint32_t w = 0;
    for (size_t k = 0; k < 32 && i + k < n; k++) w |= uint32_t(flags[i + k] & 1u) << k;
    blob.push_back(w);
  }
  for (const b2dgpu_segment& sg : scratch.segs) { blob.push_back(sg.p0); blob.push_back(sg.p1_kind); }

  // Trust, but verify: the replay must reproduce the reference bit for bit.
  const b2d::GlyphBlobView view = b2d::glyph_blob_view(blob.data());
  if (view.total_words() != blob.size()) return host;
  {
    const double m[6] = { 1.0, 0.0, 0.0, 1.0, 0.0, 0.0 };
    GlyphPathPut put = { path.vertex_data(), nv, true };
    if (b2d::glyph_emit(view, m, put) != nv || !put.ok) return host;
  }
  {
    const BLMatrix2D general(0.37109375, -0.113, 0.2291, 0.90625, 5.53, -3.2517);
    BLPath path2;
    if (face_impl->funcs.get_glyph_outlines(face_impl, glyph_id, &general, &path2, &contour_count, &tmp_buffer) != BL_SUCCESS || path2.size() != nv) return host;
    const double m[6] = { general.m00, general.m01, general.m10, general.m11, general.m20, general.m21 };
    GlyphPathPut put = { path2.vertex_data(), nv, true };
    if (b2d::glyph_emit(view, m, put) != nv || !put.ok) return host;
  }

  State::GlyphEntry e = { uint32_t(st->glyph_cache.size()), uint32_t(nv), uint32_t(ns), 0, 0, 0 };
  st->glyph_cache.insert(st->glyph_cache.end(), blob.begin(), blob.end());
  return e;
}

// A filled glyph run as glyph instances: the per-glyph matrices of bl_font_get_glyph_run_outlines (core/font.cpp:659-744)
// - font matrix x user transform, translated by the glyph's placement / advance - without decoding an outline.  Returns
// false (and records nothing) when some glyph of the run has to be decoded on the host; the caller then takes the
// reference's path for the whole run, so that a command's segments stay contiguous.
static bool instance_glyph_run(State* st, const BLFontCore* font, const BLGlyphRun* glyph_run, const BLMatrix2D& user_transform) noexcept {
  BLFontPrivateImpl* font_impl = FontInternal::get_impl(font);
  BLFontFacePrivateImpl* face_impl = FontFaceInternal::get_impl(&font_impl->face);
  if (face_impl->face_info.outline_type != BL_FONT_OUTLINE_TYPE_TRUETYPE) return false;
  if (!glyph_run->size) return true;

  const size_t first_instance = st->instances.size();
  const uint32_t gen_v0 = st->gen_vertices, gen_s0 = st->gen_segments;
  bool ok = true;

  struct Emit {
    State* st; const BLFontFacePrivateImpl* face_impl; bool* ok;
    void operator()(BLGlyphId glyph_id, const BLMatrix2D& m, int level = 0) const noexcept {
    const State::GlyphEntry e = glyph_entry(st, face_impl, glyph_id);
    if (e.kind == 1) return;
    if (e.kind == 3 && level < 15) {
      // a compound glyph: one instance per component with the matrix the reference composes (otglyf.cpp:601)
      for (uint32_t k = 0; k < e.comp_count && *ok; k++) {
        const State::GlyphComponent comp = st->glyph_components[e.comp_begin + k];
        BLMatrix2D cm = comp.local;
        TransformInternal::multiply(cm, cm, m);
        (*this)(comp.glyph_id, cm, level + 1);
      }
      return;
    }
    if (e.kind != 0) { *ok = false; return; }
    b2dgpu_glyph_instance gi;
    gi.m[0] = m.m00; gi.m[1] = m.m01; gi.m[2] = m.m10; gi.m[3] = m.m11; gi.m[4] = m.m20; gi.m[5] = m.m21;
    gi.blob_offset = e.blob_offset;
    gi.vertex_base = st->gen_vertices;
    gi.segment_base = st->gen_segments;
    gi.command = 0;                                // patched in consume_batch()
    st->instances.push_back(gi);
    st->gen_vertices += e.vertices;
    st->gen_segments += e.segments;
    }
  } emit = { st, face_impl, &ok };

  BLMatrix2D final_transform;
  const BLFontMatrix& fMat = font_impl->matrix;
  bl_font_matrix_multiply(&final_transform, &fMat, &user_transform);

  const uint32_t placement_type = glyph_run->placement_type;
  BLGlyphRunIterator it(*glyph_run);
  if (it.has_placement() && placement_type != BL_GLYPH_PLACEMENT_TYPE_NONE) {
    BLMatrix2D offset_transform(1.0, 0.0, 0.0, 1.0, final_transform.m20, final_transform.m21);
    switch (placement_type) {
      case BL_GLYPH_PLACEMENT_TYPE_ADVANCE_OFFSET:
      case BL_GLYPH_PLACEMENT_TYPE_DESIGN_UNITS:
        offset_transform.m00 = final_transform.m00; offset_transform.m01 = final_transform.m01;
        offset_transform.m10 = final_transform.m10; offset_transform.m11 = final_transform.m11;
        break;
      case BL_GLYPH_PLACEMENT_TYPE_USER_UNITS:
        offset_transform.m00 = user_transform.m00; offset_transform.m01 = user_transform.m01;
        offset_transform.m10 = user_transform.m10; offset_transform.m11 = user_transform.m11;
        break;
    }
    if (placement_type == BL_GLYPH_PLACEMENT_TYPE_ADVANCE_OFFSET) {
      double ox = final_transform.m20, oy = final_transform.m21;
      while (!it.at_end() && ok) {
        const BLGlyphPlacement& pos = it.placement<BLGlyphPlacement>();
        double px = pos.placement.x, py = pos.placement.y;
        final_transform.m20 = px * offset_transform.m00 + py * offset_transform.m10 + ox;
        final_transform.m21 = px * offset_transform.m01 + py * offset_transform.m11 + oy;
        emit(it.glyph_id(), final_transform);
        it.advance();
        px = pos.advance.x; py = pos.advance.y;
        ox += px * offset_transform.m00 + py * offset_transform.m10;
        oy += px * offset_transform.m01 + py * offset_transform.m11;
      }
    }
    else {
      while (!it.at_end() && ok) {
        const BLPoint& placement = it.placement<BLPoint>();
        final_transform.m20 = placement.x * offset_transform.m00 + placement.y * offset_transform.m10 + offset_transform.m20;
        final_transform.m21 = placement.x * offset_transform.m01 + placement.y * offset_transform.m11 + offset_transform.m21;
        emit(it.glyph_id(), final_transform);
        it.advance();
      }
    }
  }
  else {
    while (!it.at_end() && ok) {
      emit(it.glyph_id(), final_transform);
      it.advance();
    }
  }

  if (!ok) {
    st->instances.resize(first_instance);
    st->gen_vertices = gen_v0; st->gen_segments = gen_s0;
    st->glyph_runs_on_host++;
  }
  else st->glyphs_instanced += st->instances.size() - first_instance;
  return ok;
}

struct JobGeometry {
  bool valid;
  bool generated;               // the segments come from glyph instances: seg_begin is relative to the generated range
  uint32_t seg_begin, seg_count, state_index;
  uint32_t inst_begin, inst_count;
};

static BL_INLINE const BLGlyphRun* job_glyph_run(WorkData* work_data, RenderJob_TextOp* job, BLResult& result) noexcept {
  const uint32_t data_type = job->text_data_type();
  if (data_type == RenderJob::kTextDataGlyphRun)
    return &job->_glyph_run;
  BLGlyphBuffer* glyph_buffer;
  if (data_type != RenderJob::kTextDataGlyphBuffer) {
    glyph_buffer = &work_data->glyph_buffer;
    glyph_buffer->set_text(job->text_data(), job->text_size(), (BLTextEncoding)data_type);
  }
  else {
    glyph_buffer = &job->_glyph_buffer.dcast();
  }
  result = job->_font.dcast().shape(*glyph_buffer);
  return &glyph_buffer->glyph_run();
}

static JobGeometry process_job(State* st, WorkData* work_data, RenderJob* base_job) noexcept {
  JobGeometry out = { false, false, uint32_t(st->segs.size()), 0, 0, 0, 0 };
  const size_t vtx_begin = st->vtx.size();
  BLResult result = BL_SUCCESS;
  BLMatrix2D state_transform;
  BLTransformType state_transform_type = BL_TRANSFORM_TYPE_IDENTITY;
  const SharedFillState* fill_state = static_cast<RenderJob_BaseOp*>(base_job)->fill_state();

  switch (base_job->job_type()) {
    case RenderJobType::kFillGeometry: {
      RenderJob_GeometryOp* job = static_cast<RenderJob_GeometryOp*>(base_job);
      JobProc::JobStateAccessor accessor(job);
      BLPath* path = JobProc::get_geometry_as_path(work_data, job);
      if (path) {
        append_path(st, path->view(), true);
        state_transform = accessor.final_transform_fixed(job->origin_fixed());
        state_transform_type = accessor.final_transform_fixed_type();
      }
      else result = BL_ERROR_INVALID_GEOMETRY;
      JobProc::finalize_geometry_data(work_data, job);
      break;
    }

    case RenderJobType::kStrokeGeometry: {
      // add_stroked_path_edges (rastercontextops_p.h:100-146) with a segment sink.
      RenderJob_GeometryOp* job = static_cast<RenderJob_GeometryOp*>(base_job);
      JobProc::JobStateAccessor accessor(job);
      const BLPath* path = JobProc::get_geometry_as_path(work_data, job);
      if (path) {
        SegmentSink sink = { st, st->tmp_path, nullptr, nullptr };
        BLPath* a = &st->tmp_path[0];
        BLPath* b = &st->tmp_path[1];
        BLPath* c = &st->tmp_path[2];
        state_transform = accessor.final_transform_fixed(job->origin_fixed());
        state_transform_type = accessor.final_transform_fixed_type();
        if (accessor.stroke_options().transform_order != BL_STROKE_TRANSFORM_ORDER_AFTER) {
          BLPath* in = &st->tmp_path[3];
          in->clear();
          result = in->add_path(*path, accessor.user_transform());
          path = in;
          state_transform = accessor.meta_transform_fixed(job->origin_fixed());
          state_transform_type = accessor.meta_transform_fixed_type();
        }
        if (result == BL_SUCCESS) {
          a->clear();
          result = PathInternal::stroke_path(path->view(), accessor.stroke_options(), accessor.approximation_options(),
                                             *a, *b, *c, stroke_geometry_segment_sink, &sink);
        }
      }
      else result = BL_ERROR_INVALID_GEOMETRY;
      JobProc::finalize_geometry_data(work_data, job);
      break;
    }

    case RenderJobType::kFillText: {
      // add_filled_glyph_run_edges (rastercontextops_p.h:76-98): outlines arrive already transformed to 24.8 space.
      RenderJob_TextOp* job = static_cast<RenderJob_TextOp*>(base_job);
      JobProc::JobStateAccessor accessor(job);
      const BLGlyphRun* glyph_run = job_glyph_run(work_data, job, result);
      if (result == BL_SUCCESS) {
        BLMatrix2D transform(accessor.final_transform_fixed(job->origin_fixed()));
        state_transform.reset();
        state_transform_type = BL_TRANSFORM_TYPE_IDENTITY;
        const size_t inst_begin = st->instances.size();
        const uint32_t gen_begin = st->gen_segments;
        if (st->use_glyph_cache && instance_glyph_run(st, &job->_font, glyph_run, transform)) {
          // every glyph of the run is in the device cache: no outline is decoded here
          job->destroy();
          out.generated = true;
          out.seg_begin = gen_begin;
          out.seg_count = st->gen_segments - gen_begin;
          out.inst_begin = uint32_t(inst_begin);
          out.inst_count = uint32_t(st->instances.size() - inst_begin);
          if (!out.seg_count) return out;
          out.state_index = add_state(st, state_transform, state_transform_type, fill_state->final_clip_box_fixed_d, fill_state->toleranceFixedD);
          out.valid = true;
          return out;
        }
        BLPath* path = &st->tmp_path[4];
        path->clear();
        SegmentSink sink = { st, st->tmp_path, nullptr, nullptr };
        result = bl_font_get_glyph_run_outlines(&job->_font, glyph_run, &transform, path, fill_glyph_run_segment_sink, &sink);
      }
      job->destroy();
      break;
    }

    case RenderJobType::kStrokeText: {
      // add_stroked_glyph_run_edges (rastercontextops_p.h:148-195).
      RenderJob_TextOp* job = static_cast<RenderJob_TextOp*>(base_job);
      JobProc::JobStateAccessor accessor(job);
      const BLGlyphRun* glyph_run = job_glyph_run(work_data, job, result);
      if (result == BL_SUCCESS) {
        SegmentSink sink = { st, st->tmp_path, &accessor.stroke_options(), &accessor.approximation_options() };
        BLMatrix2D glyph_run_transform;
        if (accessor.stroke_options().transform_order == BL_STROKE_TRANSFORM_ORDER_AFTER) {
          glyph_run_transform.reset();
          state_transform = accessor.final_transform_fixed(job->origin_fixed());
          state_transform_type = accessor.final_transform_fixed_type();
        }
        else {
          glyph_run_transform = accessor.user_transform();
          state_transform = accessor.meta_transform_fixed(job->origin_fixed());
          state_transform_type = accessor.meta_transform_fixed_type();
        }
        BLPath* path = &st->tmp_path[4];
        path->clear();
        result = bl_font_get_glyph_run_outlines(&job->_font, glyph_run, &glyph_run_transform, path, stroke_glyph_run_segment_sink, &sink);
      }
      job->destroy();
      break;
    }

    default:
      result = BL_ERROR_INVALID_STATE;
      break;
  }

  if (result != BL_SUCCESS) {
    work_data->accumulate_error(result);
    st->segs.resize(out.seg_begin);
    st->vtx.resize(vtx_begin);
    return out;
  }

  out.seg_count = uint32_t(st->segs.size()) - out.seg_begin;
  if (!out.seg_count) {
    st->vtx.resize(vtx_begin);
    return out;
  }
  out.state_index = add_state(st, state_transform, state_transform_type, fill_state->final_clip_box_fixed_d, fill_state->toleranceFixedD);
  out.valid = true;
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// Seam A: batch consumer
// ---------------------------------------------------------------------------------------------------------------
static BL_INLINE uint32_t fill_rule_mask(uint32_t fill_rule) noexcept {
  return fill_rule == BL_FILL_RULE_NON_ZERO ? B2DGPU_FILL_RULE_MASK_NON_ZERO : B2DGPU_FILL_RULE_MASK_EVEN_ODD;
}

// CPU-built edges (EdgeVector lists, edgestorage_p.h:38-60) -> one record per line in its original direction.
static void append_edge_vectors(State* st, const EdgeVector<int>* ev) noexcept {
  for (; ev; ev = ev->next) {
    const size_t count = ev->count();
    const bool flipped = ev->sign_bit() != 0;
    for (size_t i = 1; i < count; i++) {
      b2dgpu_edge e;
      if (!flipped) { e.x0 = ev->pts[i - 1].x; e.y0 = ev->pts[i - 1].y; e.x1 = ev->pts[i].x; e.y1 = ev->pts[i].y; }
      else { e.x0 = ev->pts[i].x; e.y0 = ev->pts[i].y; e.x1 = ev->pts[i - 1].x; e.y1 = ev->pts[i - 1].y; }
      st->edges.push_back(e);
    }
  }
}

static uint32_t add_fetch_data(State* st, const RenderFetchData* rfd) noexcept {
  // Consecutive commands usually share their style: look at the most recent entries first.
  const void* key = rfd;
  const size_t n = st->fetch_keys.size();
  for (size_t k = 0; k < n && k < 4; k++)
    if (st->fetch_keys[n - 1 - k] == key)
      return uint32_t(n - 1 - k);
  b2dgpu_fetch_data fd;
  memcpy(&fd, &rfd->pipeline_data, sizeof(fd));
  st->fetch.push_back(fd);
  st->fetch_keys.push_back(key);
  if (rfd->signature.is_gradient() && !fd.gradient.lut.data) {
    // deferred by defer_gradient_table(): the device interpolates the table from the gradient's stops, which the
    // FetchData keeps alive through its style reference (renderfetchdata_p.h:43, 181-184)
    const BLGradientPrivateImpl* gi = GradientInternal::get_impl(&rfd->style_as<BLGradientCore>());
    b2dgpu_lut_request q;
    q.fetch_index = uint32_t(n); q.stop_offset = uint32_t(st->lut_stops.size());
    q.stop_count = uint32_t(gi->size); q.lut_size = fd.gradient.lut.size;
    for (size_t k = 0; k < gi->size; k++) st->lut_stops.push_back(b2dgpu_gradient_stop{ gi->stops[k].offset, gi->stops[k].rgba.value });
    st->lut_requests.push_back(q);
  }
  return uint32_t(n);
}

static void consume_batch(BLRasterContextImpl* ctx_impl, WorkData* work_data, RenderBatch* batch) noexcept {
  State* st = state_of(ctx_impl);
  BL_STATIC_ASSERT(sizeof(b2dgpu_fetch_data) == sizeof(Pipeline::FetchData));

  BLResult result = ensure_target(ctx_impl, st);
  if (result != BL_SUCCESS)
    work_data->accumulate_error(result);

  const uint32_t command_count = batch->command_count();
  st->cmds.clear();
  st->cmds.resize(command_count);
  st->fetch.clear();
  st->fetch_keys.clear();
  st->lut_requests.clear();
  st->lut_stops.clear();
  st->edges.clear();

  // Queues are chained; a job addresses its command as (queue, index).
  struct QueueBase { const RenderCommandQueue* queue; uint32_t base; };
  std::vector<QueueBase> queue_bases;

  // ---- pass 1: jobs ----
  struct PendingJob { RenderJob* job; uint32_t command; };
  std::vector<JobGeometry> job_geometry(command_count, JobGeometry{ false, false, 0, 0, 0, 0, 0 });
  {
    uint32_t base = 0;
    for (const RenderCommandQueue* q = batch->command_list().first(); q; q = q->next()) {
      queue_bases.push_back(QueueBase{ q, base });
      base += uint32_t(q->size());
    }
  }
  auto global_index = [&](const RenderCommandQueue* q, size_t index) noexcept -> uint32_t {
    for (const QueueBase& qb : queue_bases)
      if (qb.queue == q) return qb.base + uint32_t(index);
    return 0xFFFFFFFFu;
  };

  if (batch->job_count()) {
    size_t remaining = batch->job_count();
    for (const RenderJobQueue* jq = batch->job_list().first(); jq && remaining; jq = jq->next()) {
      for (size_t i = 0; i < jq->size() && remaining; i++, remaining--) {
        RenderJob* job = jq->at(i);
        if (st->cpu_edges) {
          JobProc::process_job(work_data, job);                     // reference EdgeBuilder; the command gets its edges
          continue;
        }
        const uint32_t ci = global_index(job->command_queue(), job->command_index());
        JobGeometry g = process_job(st, work_data, job);
        if (ci < command_count)
          job_geometry[ci] = g;
      }
    }
  }

  // ---- pass 2: commands ----
  bool ok = result == BL_SUCCESS;
  uint32_t out_count = 0;
  uint32_t ci = 0;
  for (const RenderCommandQueue* q = batch->command_list().first(); q && ok; q = q->next()) {
    for (size_t i = 0; i < q->size(); i++, ci++) {
      const RenderCommand& rc = q->at(i);
      const Pipeline::DispatchData* dd = rc.pipe_dispatch_data();
      if (!B2DGPU_DISPATCH_IS_GPU(dd)) { work_data->accumulate_error(BL_ERROR_INVALID_STATE); continue; }

      b2dgpu_command c;
      memset(&c, 0, sizeof(c));
      c.signature = B2DGPU_DISPATCH_SIGNATURE(dd);
      c.alpha = rc.alpha();

      switch (rc.type()) {
        case RenderCommandType::kFillBoxA:
        case RenderCommandType::kFillBoxU: {
          c.type = rc.type() == RenderCommandType::kFillBoxA ? B2DGPU_CMD_FILL_BOX_A : B2DGPU_CMD_FILL_BOX_U;
          const BLBoxI& b = rc.box_i();
          c.box[0] = b.x0; c.box[1] = b.y0; c.box[2] = b.x1; c.box[3] = b.y1;
          if (!(b.x0 < b.x1 && b.y0 < b.y1)) continue;
          break;
        }

        case RenderCommandType::kFillAnalytic: {
          c.fill_rule_mask = fill_rule_mask(rc.analytic_fill_rule());
          const EdgeVector<int>* ev = rc.analytic_edges();
          if (is_direct_tag(ev)) {
            const DirectGeometry& dg = st->direct[direct_index(ev)];
            c.type = B2DGPU_CMD_FILL_GEOMETRY;
            c.data_offset = dg.seg_begin; c.data_count = dg.seg_count; c.state_index = dg.state_index;
          }
          else if (ev) {
            c.type = B2DGPU_CMD_FILL_ANALYTIC;
            c.data_offset = uint32_t(st->edges.size());
            append_edge_vectors(st, ev);
            c.data_count = uint32_t(st->edges.size()) - c.data_offset;
            if (!c.data_count) continue;
          }
          else if (job_geometry[ci].valid) {
            const JobGeometry& g = job_geometry[ci];
            c.type = B2DGPU_CMD_FILL_GEOMETRY;
            c.data_offset = g.seg_begin; c.data_count = g.seg_count; c.state_index = g.state_index;
            if (g.generated) {
              // glyph instances: their segments live behind the uploaded ones (all jobs have run: the count is final)
              c.data_offset += uint32_t(st->segs.size());
              for (uint32_t k = 0; k < g.inst_count; k++) st->instances[g.inst_begin + k].command = out_count;
              break;
            }
          }
          else continue;                                            // everything clipped out / failed job
          if (c.type == B2DGPU_CMD_FILL_GEOMETRY)
            for (uint32_t s = 0; s < c.data_count; s++)
              st->segs[c.data_offset + s].command = out_count;
          break;
        }

        case RenderCommandType::kFillBoxMaskA: {
          const RenderCommand::FillBoxMaskA& p = rc._payload.box_mask_a;
          const BLImageImpl* mask = p.mask_image_i.ptr;
          if (mask->depth != 8) { work_data->accumulate_error(BL_ERROR_NOT_IMPLEMENTED); continue; }
          c.type = B2DGPU_CMD_FILL_BOX_MASK_A;
          c.box[0] = p.box_i.x0; c.box[1] = p.box_i.y0; c.box[2] = p.box_i.x1; c.box[3] = p.box_i.y1;
          if (!(p.box_i.x0 < p.box_i.x1 && p.box_i.y0 < p.box_i.y1)) continue;
          b2dgpu_fetch_data fd;
          memset(&fd, 0, sizeof(fd));
          fd.pattern.src.pixel_data = static_cast<const uint8_t*>(mask->pixel_data) + intptr_t(p.mask_offset_i.y) * mask->stride + p.mask_offset_i.x;
          fd.pattern.src.stride = mask->stride;
          fd.pattern.src.w = p.box_i.x1 - p.box_i.x0;
          fd.pattern.src.h = p.box_i.y1 - p.box_i.y0;
          c.reserved[0] = uint32_t(st->fetch.size());
          st->fetch.push_back(fd);
          st->fetch_keys.push_back(nullptr);
          break;
        }

        default:
          work_data->accumulate_error(BL_ERROR_NOT_IMPLEMENTED);
          continue;
      }

      if (rc.has_style_fetch_data())
        c.fetch_index = add_fetch_data(st, static_cast<const RenderFetchData*>(rc._source.fetch_data));
      else
        c.solid_prgb32 = rc._source.solid.prgb32;
      st->cmds[out_count++] = c;
    }
  }

  if (ok && out_count) {
    b2dgpu_batch_view v;
    memset(&v, 0, sizeof(v));
    v.struct_size = sizeof(v);
    v.command_count = out_count;
    v.commands = st->cmds.data();
    v.fetch_data = st->fetch.data();        v.fetch_count = uint32_t(st->fetch.size());
    v.edges = st->edges.data();             v.edge_count = uint32_t(st->edges.size());
    v.vertices = st->vtx.data();            v.vertex_count = uint32_t(st->vtx.size() / 2);
    v.segments = st->segs.data();           v.segment_count = uint32_t(st->segs.size());
    v.geometry_states = st->states.data();  v.geometry_state_count = uint32_t(st->states.size());
    v.lut_requests = st->lut_requests.data();  v.lut_request_count = uint32_t(st->lut_requests.size());
    v.lut_stops = st->lut_stops.data();        v.lut_stop_count = uint32_t(st->lut_stops.size());
    if (!st->instances.empty()) {
      // the generated ranges start where the uploaded arrays end
      const uint32_t v0 = v.vertex_count, s0 = v.segment_count;
      for (b2dgpu_glyph_instance& gi : st->instances) { gi.vertex_base += v0; gi.segment_base += s0; }
      v.glyph_cache = st->glyph_cache.data(); v.glyph_cache_words = uint32_t(st->glyph_cache.size());
      v.glyph_cache_id = uint64_t(uintptr_t(st));
      v.glyph_instances = st->instances.data(); v.glyph_instance_count = uint32_t(st->instances.size());
      v.generated_vertex_count = st->gen_vertices; v.generated_segment_count = st->gen_segments;
    }
    v.pixel_origin_x = work_data->ctx_data.pixel_origin.x;
    v.pixel_origin_y = work_data->ctx_data.pixel_origin.y;
    // Copies everything it needs before returning (rastercontext.cpp:1060-1063 frees the batch right after).
    b2dgpu_result r = b2dgpu_submit(st->rt, st->target, &v);
    if (r != B2DGPU_SUCCESS)
      work_data->accumulate_error(bl_make_error(BLResult(r)));
    else
      st->device_dirty = true;
  }

  st->clear_geometry();

  if (st->adaptive_batches) {
    WorkerManager& mgr = ctx_impl->worker_mgr();
    // Batches grow by a factor of 4 (512, 2048, 8192): the first one starts the device early, the later ones keep the number
    // of passes over the canvas small - measured on config 1 (recording is ~2.7x faster than compositing): e2e 16.82 ms with
    // a factor of 2, 16.48 with 3, 16.16 with 4, 16.41 with 6, 16.8 with 8 (the device idles).  B2DGPU_SHIM_BATCH_GROWTH overrides.
    static const uint32_t growth = [] { const char* e = getenv("B2DGPU_SHIM_BATCH_GROWTH"); const int g = e ? atoi(e) : 4; return uint32_t(g >= 2 && g <= 8 ? g : 4); }();
    if (mgr._command_queue_limit < kLargestBatch) mgr._command_queue_limit = bl_min<uint32_t>(mgr._command_queue_limit * growth, kLargestBatch);
  }
}

} // {GpuShim}

#endif // B2DGPU_SHIM_IMPL_H_INCLUDED
