#!/usr/bin/env python3
"""Produces the GPU-enabled copy of the reference's raster/rastercontext.cpp in a build tree.

The reference tree is read-only and none of its source is kept in this repository: this script reads
<ref>/blend2d/raster/rastercontext.cpp where it lies, inserts the handful of hook lines listed in HOOKS (each anchored
on one exact line of the original, which must occur exactly once) and writes the result to <out>.  Everything the
hooks call lives in shim/b2dgpu_shim_fwd.h and shim/b2dgpu_shim_impl.h.

usage: apply_overlay.py <ref>/blend2d/raster/rastercontext.cpp <out>.cpp
"""
import sys

# (anchor line, where, inserted text).  `where`: "before" / "after" / "replace".
HOOKS = [
    # 0. declarations (file scope, before the engine's namespace opens)
    ("namespace bl::RasterEngine {", "before",
     '#include "b2dgpu_shim_fwd.h"\n\n'),

    # 1. flush_render_batch(), rastercontext.cpp:1043-1053: the user thread is the only worker of a GPU context; it
    #    hands the batch to the device instead of running the CPU command processors.
    ("      WorkerProc::process_work_data(work_data, batch);", "replace",
     "      if (GpuShim::is_gpu(ctx_impl))\n"
     "        GpuShim::consume_batch(ctx_impl, work_data, batch);\n"
     "      else\n"
     "        WorkerProc::process_work_data(work_data, batch);\n"),

    # 2. flush_impl(), rastercontext.cpp:1487-1489: BL_CONTEXT_FLUSH_SYNC makes the image's host pixels coherent.
    ("    BL_PROPAGATE(flush_render_batch(ctx_impl));", "after",
     "    if (GpuShim::is_gpu(ctx_impl))\n"
     "      BL_PROPAGATE(GpuShim::sync_to_host(ctx_impl));\n"),

    # 2b. ensure_fetch_and_dispatch_data_slow(), rastercontext.cpp:1130-1131: a gradient whose table does not exist yet is
    #     not interpolated on the host for a GPU context - the device builds it from the stops.
    ("    BL_PROPAGATE(compute_pending_fetch_data(static_cast<RenderFetchData*>(fetch_data)));", "replace",
     "    if (!(GpuShim::is_gpu(ctx_impl) && GpuShim::defer_gradient_table(ctx_impl, static_cast<RenderFetchData*>(fetch_data))))\n"
     "      BL_PROPAGATE(compute_pending_fetch_data(static_cast<RenderFetchData*>(fetch_data)));\n"),

    # 3. implementation, placed after the asynchronous enqueue helpers it uses (rastercontext.cpp:2410-2630).
    ("// bl::RasterEngine - ContextImpl - Internals - Fill Clipped Box", "before",
     '#include "b2dgpu_shim_impl.h"\n\n'),

    # 4. fill_unclipped_path<kRM>(), rastercontext.cpp:2787-2797: small paths are recorded, not flattened on the CPU.
    ("  BL_PROPAGATE(add_filled_path_edges(&ctx_impl->sync_work_data, path.view(), transform, transform_type));", "before",
     "  if constexpr (kRM == kAsync) {\n"
     "    if (GpuShim::records_geometry(ctx_impl))\n"
     "      return GpuShim::record_path(ctx_impl, di, ds, path.view(), fill_rule, transform, transform_type);\n"
     "  }\n"),

    # 5. fill_unclipped_polygon_t<kRM>(), rastercontext.cpp:2848-2858: same for polygons.
    ("  BL_PROPAGATE(add_filled_polygon_edges(&ctx_impl->sync_work_data, pts, size, transform, transform_type));", "before",
     "  if constexpr (kRM == kAsync) {\n"
     "    if (GpuShim::records_geometry(ctx_impl))\n"
     "      return GpuShim::record_poly(ctx_impl, di, ds, pts, size, fill_rule, transform, transform_type);\n"
     "  }\n"),

    # 6. attach(), rastercontext.cpp:4211-4216: a GPU context is asynchronous with the user thread as its worker.
    ("  BL_ASSERT(options != nullptr);", "after",
     "  BLContextCreateInfo gpu_create_info_storage;\n"
     "  options = GpuShim::adjust_create_info(options, &gpu_create_info_storage);\n"),

    # 7. attach() step 3, rastercontext.cpp:4256-4293: BL_CONTEXT_CREATE_FLAG 0x10000000 selects the GPU PipeRuntime.
    ("    // Step 4: Allocate zeroed memory for the user thread and all worker threads.", "before",
     "    if (options->flags & GpuShim::kCreateFlagGpuRuntime) {\n"
     "      result = GpuShim::create_runtime(options, &pipe_runtime);\n"
     "      if (result != BL_SUCCESS)\n"
     "        break;\n"
     "    }\n\n"),
    # detach() needs no hook: the runtime is flagged kIsolated, so rastercontext.cpp:4447-4450 destroys it (and with
    # it the device canvas) after the final flush_impl(BL_CONTEXT_FLUSH_SYNC).
]


def main():
    src, out = sys.argv[1], sys.argv[2]
    lines = open(src).read().split("\n")
    for anchor, where, text in HOOKS:
        hits = [i for i, l in enumerate(lines) if l == anchor]
        if len(hits) != 1:
            sys.exit(f"apply_overlay: anchor {anchor!r} matches {len(hits)} lines (expected exactly 1) - "
                     "the reference changed; update shim/apply_overlay.py")
        i = hits[0]
        ins = text.rstrip("\n").split("\n") if text.strip() else []
        if text.endswith("\n\n"):
            ins.append("")
        if where == "before":
            lines[i:i] = ins
        elif where == "after":
            lines[i + 1:i + 1] = ins
        else:
            lines[i:i + 1] = ins
    open(out, "w").write("\n".join(lines))


if __name__ == "__main__":
    main()
