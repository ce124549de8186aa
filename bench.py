#!/usr/bin/env python
"""bench.py - measures the rendering hot path on the workload BASELINE.json quotes the metric on.

Workload (config 1): 10 000 fills per step on a 3840x2160 PRGB32 canvas - bl_bench polygons (10/20/40 points in an
8..256 px box) and the reference tester's random quad / cubic paths (points uniform in the canvas +-30 px), NonZero and
EvenOdd alternating, linear / radial / conic gradients with pad / repeat / reflect extend, SrcOver.  The scene is
generated with numpy from a fixed seed and replayed natively (C) through either front end, so Python never runs inside
a timed region.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation (rank 0 only)

`value`  : Mpix/s (pixels composited per second) with every input already resident in HBM: K x b2dgpu_batch_render().
`e2e`    : the same metric through the public host API with HOST buffers: per step b2d_scene_replay() (front-end work)
           + flush(SYNC) = serialise + H2D + kernels + D2H of the canvas into the host image.
`roofline`: the tile compositor (k_tile_render), algorithmic bytes = 8 B per composited pixel (4 read + 4 written,
           SURVEY 8d) / its CUDA-event duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160


# ---------------------------------------------------------------------------------------------------------------------
# Scene description (include/b2d_scene.h)
# ---------------------------------------------------------------------------------------------------------------------
class SceneStop(C.Structure):
    _fields_ = [("offset", C.c_double), ("rgba64", C.c_uint64)]


class SceneFill(C.Structure):
    _fields_ = [("geom", C.c_uint32), ("vtx_offset", C.c_uint32), ("vtx_count", C.c_uint32), ("fill_rule", C.c_uint32),
                ("comp_op", C.c_uint32), ("style", C.c_uint32), ("extend", C.c_uint32), ("stop_offset", C.c_uint32),
                ("stop_count", C.c_uint32), ("rgba32", C.c_uint32), ("quality", C.c_uint32), ("has_transform", C.c_uint32),
                ("rect", C.c_double * 4), ("values", C.c_double * 6), ("angle", C.c_double), ("cx", C.c_double), ("cy", C.c_double)]


class Scene(C.Structure):
    _fields_ = [("fills", C.POINTER(SceneFill)), ("fill_count", C.c_uint32), ("_pad0", C.c_uint32),
                ("vertices", C.POINTER(C.c_double)), ("vertex_count", C.c_uint32), ("_pad1", C.c_uint32),
                ("path_cmds", C.POINTER(C.c_uint8)),
                ("stops", C.POINTER(SceneStop)), ("stop_count", C.c_uint32), ("_pad2", C.c_uint32),
                ("texture", C.POINTER(C.c_uint32)), ("texture_w", C.c_int32), ("texture_h", C.c_int32)]


def _rgba64(c):
    a, r, g, b = (c >> 24) & 0xFF, (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF
    return ((a * 0x101) << 48) | ((r * 0x101) << 32) | ((g * 0x101) << 16) | (b * 0x101)


def make_config1_scene(n_fills, W, H, seed):
    """Config 1 of BASELINE.json.  Returns (Scene, keepalive)."""
    rng = np.random.default_rng(seed)
    fills = (SceneFill * n_fills)()
    vtx, cmds, stops = [], [], []
    sizes = [8, 16, 32, 64, 128, 256]
    for i in range(n_fills):
        f = fills[i]
        kind = i % 3
        if os.environ.get("B2D_BENCH_KIND"):             # experiment knob: 0 polygons only, 1 quads only, 2 cubics only
            kind = int(os.environ["B2D_BENCH_KIND"])
        f.fill_rule = (i // 3) % 2
        f.comp_op = 0
        f.style = 1 + (i % 3 + i // 7) % 3
        if os.environ.get("B2D_BENCH_STYLE"):            # experiment knob: force one gradient type (1 linear, 2 radial, 3 conic)
            f.style = int(os.environ["B2D_BENCH_STYLE"])
        f.extend = (i // 5) % 3
        f.quality = 0
        f.vtx_offset = len(vtx)
        if kind == 0:
            s = sizes[int(rng.integers(0, len(sizes)))]
            npts = (10, 20, 40)[int(rng.integers(0, 3))]
            bx, by = rng.uniform(0, W - s), rng.uniform(0, H - s)
            xs, ys = rng.uniform(bx, bx + s, npts), rng.uniform(by, by + s, npts)
            f.geom = 2
            for x, y in zip(xs, ys):
                vtx.append((x, y)); cmds.append(1)
        else:
            m = 30.0
            k = 3 if kind == 1 else 4
            xs, ys = rng.uniform(-m, W + m, k), rng.uniform(-m, H + m, k)
            f.geom = 3
            vtx.append((xs[0], ys[0])); cmds.append(0)
            if kind == 1:
                vtx += [(xs[1], ys[1]), (xs[2], ys[2])]; cmds += [2, 1]
            else:
                vtx += [(xs[1], ys[1]), (xs[2], ys[2]), (xs[3], ys[3])]; cmds += [4, 4, 1]
        f.vtx_count = len(vtx) - f.vtx_offset
        bx0, by0 = float(xs.min()), float(ys.min())
        bw, bh = float(xs.max()) - bx0, float(ys.max()) - by0
        c = [int(v) for v in rng.integers(0, 2 ** 32, 4)]
        f.stop_offset = len(stops)
        if f.style == 1:
            vals = [bx0 + bw * 0.2, by0 + bh * 0.2, bx0 + bw * 0.8, by0 + bh * 0.8, 0, 0]
            stops += [(0.0, c[0]), (0.5, c[1]), (1.0, c[2])]
        elif f.style == 2:
            cx, cy, cr = bx0 + bw / 2, by0 + bh / 2, (bw + bh) / 4
            vals = [cx, cy, cx - cr / 2, cy - cr / 2, cr, 0.0]
            stops += [(0.0, c[0]), (0.5, c[1]), (1.0, c[2])]
        else:
            vals = [bx0 + bw / 2, by0 + bh / 2, 0.0, 1.0, 0, 0]
            stops += [(0.0, c[0]), (0.33, c[1]), (0.66, c[2]), (1.0, c[3])]
        f.stop_count = len(stops) - f.stop_offset
        for j, v in enumerate(vals):
            f.values[j] = v
    vtx_arr = np.ascontiguousarray(np.asarray(vtx, dtype=np.float64))
    cmd_arr = np.ascontiguousarray(np.asarray(cmds, dtype=np.uint8))
    stop_arr = (SceneStop * len(stops))()
    for j, (o, c) in enumerate(stops):
        stop_arr[j].offset = o
        stop_arr[j].rgba64 = _rgba64(c)
    sc = Scene()
    sc.fills = fills; sc.fill_count = n_fills
    sc.vertices = vtx_arr.ctypes.data_as(C.POINTER(C.c_double)); sc.vertex_count = len(vtx)
    sc.path_cmds = cmd_arr.ctypes.data_as(C.POINTER(C.c_uint8))
    sc.stops = stop_arr; sc.stop_count = len(stops)
    sc.texture = None; sc.texture_w = 0; sc.texture_h = 0
    return sc, (fills, vtx_arr, cmd_arr, stop_arr)


# ---------------------------------------------------------------------------------------------------------------------
# The reference arm (CPU): oracle/_ref/libref_scene_driver.so, built from /root/reference by oracle/Makefile.ref
# ---------------------------------------------------------------------------------------------------------------------
def load_ref_driver():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_scene_driver.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_scene_run.restype = C.c_uint32
    lib.ref_scene_run.argtypes = [C.POINTER(Scene), C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.POINTER(C.c_double), C.c_void_p, C.c_ssize_t]
    lib.ref_scene_count_pixels.restype = C.c_uint32
    lib.ref_scene_count_pixels.argtypes = [C.POINTER(Scene), C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    return lib


def host_threads():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(32, n))            # BL_RUNTIME_MAX_THREAD_COUNT = 32 (blend2d/core/runtime.h:25)


def run_reference(scene, sample, W, H, steps, warmup, threads, return_pixels=False):
    """Times the reference's CPU renderer (async MT) on the first `sample` fills.  Returns a dict or None."""
    lib = load_ref_driver()
    if lib is None:
        return None
    px = C.c_uint64(0)
    rc = lib.ref_scene_count_pixels(C.byref(scene), 0, sample, W, H, C.byref(px))
    if rc != 0:
        raise RuntimeError(f"ref_scene_count_pixels failed: 0x{rc:08X}")
    secs = (C.c_double * (steps + warmup))()
    pixels = np.zeros((H, W), dtype=np.uint32) if return_pixels else None
    rc = lib.ref_scene_run(C.byref(scene), 0, sample, W, H, 1, threads, steps + warmup, secs,
                           pixels.ctypes.data_as(C.c_void_p) if return_pixels else None, W * 4 if return_pixels else 0)
    if rc != 0:
        raise RuntimeError(f"ref_scene_run failed: 0x{rc:08X}")
    timed = list(secs)[warmup:]
    total = sum(timed)
    out = {"mpix_s": px.value * steps / total / 1e6, "fills_s": sample * steps / total, "ms_per_step": total / steps * 1e3,
           "pixels_per_step": int(px.value), "threads": threads}
    if return_pixels:
        out["pixels"] = pixels
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Clock sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.samples, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import blend2d_b200 as G
    from blend2d_b200 import _native as N

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - blend2d_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H, n_fills = args.width, args.height, args.fills
    # Frame sharding: every rank renders its own independent frame (per-rank seed); no data-path collective.
    scene, keep = make_config1_scene(n_fills, W, H, seed=1234 + rank)

    # A non-default torch stream: its handle is handed to the runtime, so every kernel of ours runs on the stream the
    # torch.cuda.Event timers below are recorded on.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rt = G.Runtime(device=local_rank, stream=stream.cuda_stream)
    img = G.Image(W, H, G.FORMAT_PRGB32)
    # The e2e context flushes in batches (adaptive by default: 512 commands, then doubling): batch k renders while the
    # host builds batch k + 1.
    ctx = G.Context(img, device=local_rank, runtime=rt, command_queue_limit=args.queue_limit)
    lib = N.lib
    G_check = N.check

    # ---- record the scene once and upload it as a device-resident batch (inputs in HBM) ----
    rec = G.Context(G.Image(W, H, G.FORMAT_PRGB32), record_only=True)
    G_check(lib.b2d_scene_replay(rec._h, C.byref(scene), 0, n_fills), "b2d_scene_replay(record)")
    view = rec.peek_batch()
    batch = G.ResidentBatch(rt._h, view)
    target = ctx.target_handle()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(timed):
        G_check(lib.b2dgpu_target_clear(target), "target_clear")
        flush_buf.fill_(rank + 1)                                            # evict the canvas / batch from L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        batch.render(target)
        e1.record(stream)
        return (e0, e1) if timed else None

    for _ in range(args.warmup):
        resident_step(False)
    barrier()
    rt.stats(reset=True)
    G_check(lib.b2dgpu_set_profiling(rt._h, 1), "set_profiling")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    events = [resident_step(True) for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    st = rt.stats(reset=True)
    G_check(lib.b2dgpu_set_profiling(rt._h, 0), "set_profiling")
    step_ms = [a.elapsed_time(b) for a, b in events]
    total_ms = float(sum(step_ms))
    px_per_step = st["pixels_composited"] / args.steps
    launches = int(st["kernel_launches"])
    tile_ms_avg = st["tile_kernel_ms"] / max(1, st["tile_kernel_launches"])
    build_ms_avg = st["build_kernels_ms"] / max(1, st["tile_kernel_launches"])

    # ---- end to end through the public host API with host buffers ----
    e2e_acc = {"px": 0, "h2d": 0, "d2h": 0}

    def e2e_step(timed=True):
        G_check(lib.b2d_context_clear_all(ctx._h), "clear_all")
        ctx.flush(sync=True)
        flush_buf.fill_(rank + 2)
        torch.cuda.synchronize()
        rt.stats(reset=True)
        t0 = time.perf_counter()
        G_check(lib.b2d_scene_replay(ctx._h, C.byref(scene), 0, n_fills), "b2d_scene_replay")
        ctx.flush(sync=True)                                                  # submit + D2H of the canvas into `img`
        dt = time.perf_counter() - t0
        s_ = rt.stats(reset=True)
        if timed:
            e2e_acc["px"] += s_["pixels_composited"]; e2e_acc["h2d"] += s_["h2d_bytes"]; e2e_acc["d2h"] += s_["d2h_bytes"]
        return dt

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step(False)
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    e2e_s = sum(e2e_step() for _ in range(e2e_steps))
    barrier()
    e2e_px = e2e_acc["px"] / e2e_steps
    st2 = {"h2d_bytes": e2e_acc["h2d"], "d2h_bytes": e2e_acc["d2h"]}
    checksum = int(np.bitwise_xor.reduce(img.pixels().ravel()))

    # ---- the bandwidth-bound case: full-canvas SrcOver fills, one command per launch (8 B per pixel) ----
    full = None
    if rank == 0 and not args.no_full_canvas:
        full = measure_full_canvas(G, N, rt, lib, torch, stream, flush_buf)

    # ---- reduce over ranks: time = max, work = sum ----
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w_ = torch.tensor([px_per_step, e2e_px, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(w_, op=dist.ReduceOp.SUM)
        total_ms, e2e_s = float(t[0]), float(t[1])
        px_all, e2e_px_all, launches = float(w_[0]), float(w_[1]), int(w_[2])
    else:
        px_all, e2e_px_all = px_per_step, e2e_px

    band = None
    if world > 1 and not args.no_band:
        band = measure_band_sharded(G, N, lib, torch, dist, stream, rank, world, local_rank, args)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg_bytes = px_per_step * 8.0
        achieved = alg_bytes / (tile_ms_avg * 1e-3) / 1e9 if tile_ms_avg > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_tile_render_dram_bytes_per_launch")
        except (OSError, ValueError):
            pass

        out = {
            "metric": "Mpix/s", "unit": "Mpix/s",
            "value": px_all * args.steps / (total_ms * 1e-3) / 1e6,
            "fills_per_s": n_fills * world * args.steps / (total_ms * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 (premultiplied 8-bit channels, u32 coverage cells, f64 flattening)", "data": "synthetic",
            "config": {"workload": f"config1: {n_fills} fills/step (bl_bench polygons + tester quad/cubic paths, NonZero+EvenOdd, "
                                   f"linear/radial/conic gradients pad/repeat/reflect, SrcOver) on {W}x{H} PRGB32",
                       "fills_per_step": n_fills, "canvas": [W, H], "frames": world,
                       "sharding": "frame-sharded, one frame per GPU, no collective" if world > 1 else "single frame",
                       "l2": "256 MiB buffer written between timed steps (canvas + batch evicted from L2)"},
            "e2e": {"value": e2e_px_all * e2e_steps / e2e_s / 1e6, "unit": "Mpix/s",
                    "h2d_bytes_per_step": int(st2["h2d_bytes"] / e2e_steps), "d2h_bytes_per_step": int(st2["d2h_bytes"] / e2e_steps),
                    "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                    "path": "b2d_scene_replay() -> b2d_context_* -> b2dgpu_submit(host batch) -> kernels -> b2dgpu_target_download(host image)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_tile_render<4>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": tile_ms_avg, "other_kernels_ms": build_ms_avg,
                         "peak_source": peak_src,
                         "note": "tile-resident compositing: the destination tile is read and written once per step, so "
                                 "algorithmic bytes (8 B per composited pixel) exceed DRAM traffic by the overdraw factor"},
            "clocks": clocks,
            "pixels_per_step": px_per_step, "canvas_checksum": checksum,
            "rasterizer": {"edges_per_step": st["edges"] / args.steps, "segments_per_step": int(view.segment_count),
                           "edge_builder_and_binning_ms": build_ms_avg,
                           "edges_per_s": (st["edges"] / args.steps) / (build_ms_avg * 1e-3) if build_ms_avg > 0 else None,
                           "note": "K1 = k_count_edges + scan + k_write_edges + bbox + k_band_extents (CUDA events around them); "
                                   "issue-slot utilisation of K2/K3 (k_tile_render) is in profiles/*.summary.txt"},
            "band_sharded": band,
            "roofline_full_canvas": None if full is None else {
                k: {"bound": "hbm", "kernel": "k_stream_solid<SrcOver>", "achieved": v["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": v["gbs"] / peak, "kernel_ms": v["ms"], "algorithmic_bytes_per_launch": v["bytes"],
                    "workload": v["what"]} for k, v in full.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = min(n_fills, args.cpu_sample)
            ref = run_reference(scene, sample, W, H, 1, 0, host_threads())
            if ref is not None:
                out["cpu_baseline"] = {"value": ref["mpix_s"], "unit": "Mpix/s", "cores": ref["threads"], "kind": "reference",
                                       "sample": f"first {sample} fills of the same scene, 1 step, reference built from /root/reference "
                                                 f"(portable non-JIT pipeline: asmjit is not vendored), BLContextCreateInfo.thread_count={ref['threads']}",
                                       "fills_per_s": ref["fills_s"], "ms": ref["ms_per_step"]}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "Mpix/s", "cores": 0, "kind": "reference",
                                       "sample": "oracle/_ref/libref_scene_driver.so missing"}
        print(json.dumps(out))

    batch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def measure_full_canvas(G, N, rt, lib, torch, stream, flush_buf):
    """One translucent full-canvas SrcOver fill per launch: reads 4 B and writes 4 B per pixel."""
    out = {}
    for name, (W, H) in (("4k", (3840, 2160)), ("16k", (16384, 16384))):
        img = G.Image(W, H, G.FORMAT_PRGB32)
        rec = G.Context(img, record_only=True)
        rec.set_fill_style(0x80336699)
        rec.fill_all()
        batch = G.ResidentBatch(rt._h, rec.peek_batch())
        tgt = C.c_void_p()
        N.check(lib.b2dgpu_target_create(rt._h, W, H, G.FORMAT_PRGB32, C.byref(tgt)), "target_create")
        for _ in range(3):
            batch.render(tgt)
        torch.cuda.synchronize()
        rt.stats(reset=True)
        N.check(lib.b2dgpu_set_profiling(rt._h, 1), "set_profiling")
        reps = 10
        for _ in range(reps):
            flush_buf.fill_(7)
            batch.render(tgt)
        torch.cuda.synchronize()
        st = rt.stats(reset=True)
        N.check(lib.b2dgpu_set_profiling(rt._h, 0), "set_profiling")
        ms = st["tile_kernel_ms"] / reps
        nbytes = W * H * 8.0
        out[name] = {"ms": ms, "bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9,
                     "what": f"full-canvas SrcOver solid fill (alpha 0.5) of a {W}x{H} PRGB32 canvas, L2 flushed between launches"}
        batch.close()
        N.check(lib.b2dgpu_target_destroy(tgt), "target_destroy")
    return out


def measure_band_sharded(G, N, lib, torch, dist, stream, rank, world, local_rank, args):
    """Config 5(i): ONE large canvas cut into tile-aligned slabs of rows, one per GPU (band sharding).
    Every rank replays the whole command list clipped to its slab (b2dgpu_target_create_slab); the only exchange is the
    final gather of the slabs to rank 0 over NCCL.  Strong scaling: the frame is fixed, the rows per GPU shrink."""
    from blend2d_b200 import sharding as SH
    side, n_fills = args.band_canvas, args.band_fills
    scene, keep = make_config1_scene(n_fills, side, side, seed=4321)           # the same frame on every rank
    rt = G.Runtime(device=local_rank, stream=stream.cuda_stream)
    rec = G.Context(G.Image(side, side, G.FORMAT_PRGB32), record_only=True)    # host image: clip box only, never touched
    N.check(lib.b2d_scene_replay(rec._h, C.byref(scene), 0, n_fills), "b2d_scene_replay(record)")
    batch = G.ResidentBatch(rt._h, rec.peek_batch())
    # Interleaved ownership: `k` stripes per rank spread over the canvas (coverage is not uniform over the rows).
    k = args.band_stripes
    stripes = SH.stripes_of(rank, world, k, side)
    tgts = []
    for (y0, y1) in stripes:
        t_ = C.c_void_p()
        N.check(lib.b2dgpu_target_create_slab(rt._h, side, side, y0, y1, G.FORMAT_PRGB32, C.byref(t_)), "target_create_slab")
        tgts.append(t_)

    tgt_array = (C.c_void_p * len(tgts))(*[t_.value for t_ in tgts])

    def render_all():                                                           # one geometry pass, one compositing pass per stripe
        N.check(lib.b2dgpu_batch_render_multi(rt._h, tgt_array, len(tgts), batch._h), "batch_render_multi")

    for _ in range(2):
        for t_ in tgts:
            N.check(lib.b2dgpu_target_clear(t_), "clear")
        render_all()
    torch.cuda.synchronize(); dist.barrier()
    rt.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for t_ in tgts:
        N.check(lib.b2dgpu_target_clear(t_), "clear")
    e0.record(stream); render_all(); e1.record(stream)
    torch.cuda.synchronize()
    st = rt.stats(reset=True)

    # gather of the stripes (device to device over NVLink), timed separately
    class _Mem:
        pass

    def view_of(t_, rows):
        ptr, stride, pw, ph = C.c_void_p(), C.c_ssize_t(), C.c_int32(), C.c_int32()
        N.check(lib.b2dgpu_target_device_view(t_, C.byref(ptr), C.byref(stride), C.byref(pw), C.byref(ph)), "device_view")
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (ph.value, stride.value), "typestr": "|u1", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(m, device=torch.device("cuda", local_rank))[:rows, : side * 4]
    local = [view_of(t_, y1 - y0) for t_, (y0, y1) in zip(tgts, stripes)]
    full = SH.gather_stripes(local, side, k, dst=0)                            # warm-up: NCCL channel setup, allocator
    del full
    dist.barrier(); torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    full = SH.gather_stripes(local, side, k, dst=0)
    g1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    px = torch.tensor([float(st["pixels_composited"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(px, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        out = {"workload": f"config5(i): {n_fills} fills on one {side}x{side} PRGB32 canvas, band-sharded into {world} x {k} interleaved stripes of rows",
               "render_ms_max_over_ranks": float(t[0]), "gather_ms": float(t[1]), "gathered_bytes": int(full.numel()),
               "value": float(px[0]) / (float(t[0]) * 1e-3) / 1e6, "unit": "Mpix/s", "scaling": "strong",
               "collective": "one torch.distributed.gather of the row slabs (NCCL), outside the render"}
    del full
    batch.close()
    for t_ in tgts:
        N.check(lib.b2dgpu_target_destroy(t_), "target_destroy")
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, H = args.width, args.height
    sample = min(args.fills, args.cpu_sample)
    scene, keep = make_config1_scene(args.fills, W, H, seed=1234)
    threads = host_threads()
    ref = run_reference(scene, sample, W, H, args.steps, args.warmup, threads)
    if ref is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_scene_driver.so not built (needs /root/reference)"}))
        return
    out = {
        "impl": "reference", "metric": "Mpix/s", "unit": "Mpix/s", "value": ref["mpix_s"], "fills_per_s": ref["fills_s"],
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"config1: bounded sample = first {sample} of {args.fills} fills/step (same scene, seed 1234) on {W}x{H} PRGB32",
                   "fills_per_step": sample, "canvas": [W, H]},
        "cpu_baseline": {"value": ref["mpix_s"], "unit": "Mpix/s", "cores": threads, "kind": "reference",
                         "sample": f"first {sample} fills, portable (non-JIT) pipeline, async rendering with thread_count={threads}"},
        "e2e": {"value": ref["mpix_s"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-band", action="store_true", help="skip the band-sharded 16384^2 measurement that runs when N > 1")
    ap.add_argument("--band-canvas", type=int, default=16384)
    ap.add_argument("--band-fills", type=int, default=600)
    ap.add_argument("--band-stripes", type=int, default=8, help="interleaved stripes per GPU in the band-sharded measurement")
    ap.add_argument("--queue-limit", type=int, default=0, help="commands per submitted batch on the e2e path (BLContextCreateInfo.command_queue_limit); 0 = adaptive (512, doubling)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fills", type=int, default=10000)
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--cpu-sample", type=int, default=400, help="fills in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-canvas", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
