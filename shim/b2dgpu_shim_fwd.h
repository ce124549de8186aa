// b2dgpu_shim_fwd.h - first half of the Blend2D <-> libb2dgpu binding.
//
// shim/apply_overlay.py inserts `#include "b2dgpu_shim_fwd.h"` into a BUILD-TREE COPY of the reference's
// blend2d/raster/rastercontext.cpp, right before its `namespace bl::RasterEngine {`.  This header only declares
// what the hooks placed earlier in that file need (attach(), flush_render_batch(), flush_impl()); the
// implementation (b2dgpu_shim_impl.h) is included further down, after the frontend's static helpers
// (enqueue_command(), ensure_fetch_and_dispatch_data(), ...) that it calls.
//
// Nothing here is compiled into libb2dgpu.so; it is the reference-side half of the boundary, written against the
// reference's private headers (-I/root/reference) and its C-ABI counterpart include/b2dgpu.h.
#ifndef B2DGPU_SHIM_FWD_H_INCLUDED
#define B2DGPU_SHIM_FWD_H_INCLUDED

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include <unordered_map>

#include <b2dgpu.h>
#include <dev_glyph.cuh>                         // glyph cache format + the decoder replay shared with the CUDA kernel
#include <blend2d/raster/renderjobproc_p.h>      // JobProc::process_job (CPU edge building, B2DGPU_SHIM_CPU_EDGES=1)
#include <blend2d/core/font_p.h>
#include <blend2d/core/fontface_p.h>
#include <blend2d/opentype/otface_p.h>
#include <blend2d/opentype/otglyf_p.h>
#include <blend2d/support/scopedbuffer_p.h>

namespace bl::RasterEngine {
namespace GpuShim {

//! BLContextCreateFlags bit that selects the GPU pipeline runtime (free bit, core/context.h:113-154).
//! BLContextCreateInfo::reserved[0] = CUDA device ordinal.
static constexpr uint32_t kCreateFlagGpuRuntime = 0x10000000u;

//! PipeRuntimeType of the GPU runtime (piperuntime_p.h:23-28 has 0 = static, 1 = JIT).
static constexpr uint8_t kPipeRuntimeTypeGpu = 2;

static BL_INLINE bool is_gpu(const BLRasterContextImpl* ctx_impl) noexcept {
  return uint8_t(ctx_impl->pipe_provider.runtime()->runtime_type()) == kPipeRuntimeTypeGpu;
}

//! attach(): a GPU context is always asynchronous (work must arrive as batches, SURVEY 8b "Sync mode"), the user
//! thread is its only worker.  Returns `options` or `storage` filled with the adjusted copy.
static const BLContextCreateInfo* adjust_create_info(const BLContextCreateInfo* options, BLContextCreateInfo* storage) noexcept;

//! attach() step 3: creates the GPU pipeline runtime (replaces whatever was selected before).
static BLResult create_runtime(const BLContextCreateInfo* options, Pipeline::PipeRuntime** runtime) noexcept;

//! flush_render_batch(): consumes the batch instead of WorkerProc::process_work_data().
static void consume_batch(BLRasterContextImpl* ctx_impl, WorkData* work_data, RenderBatch* batch) noexcept;

//! ensure_fetch_and_dispatch_data_slow(): leaves a pending nearest-neighbour gradient table to the device (the FetchData
//! keeps lut.data == nullptr; consume_batch() ships the stops).  Returns false when the host has to build it.
static bool defer_gradient_table(BLRasterContextImpl* ctx_impl, RenderFetchData* fetch_data) noexcept;

//! flush_impl(BL_CONTEXT_FLUSH_SYNC): makes the host pixels of the target image coherent.
static BLResult sync_to_host(BLRasterContextImpl* ctx_impl) noexcept;

} // {GpuShim}
} // {bl::RasterEngine}

#endif // B2DGPU_SHIM_FWD_H_INCLUDED
