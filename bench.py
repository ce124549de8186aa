#!/usr/bin/env python
"""bench.py - measures the rendering hot path on the workload BASELINE.json quotes the metric on.

Workload (config 1): 10 000 fills per step on a 3840x2160 PRGB32 canvas - bl_bench polygons (10/20/40 points in an
8..256 px box) and the reference tester's random quad / cubic paths (points uniform in the canvas +-30 px), NonZero and
EvenOdd alternating, linear / radial / conic gradients with pad / repeat / reflect extend, SrcOver.  Scenes are generated
with numpy from fixed seeds (bench_scenes.py) and replayed natively by ONE Blend2D application (shim/bl_scene_driver.cpp,
public bl_* C API only) that is linked twice: against the GPU-enabled Blend2D build (shim/_build/libblend2d_gpu.so:
BLContextCreateInfo.flags |= 0x10000000 selects the B200 pipeline runtime of libb2dgpu.so) for our arm and against the
unmodified reference (oracle/_ref) for the CPU arm.  Python never runs inside a timed region.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation (rank 0 only)

`value`   : Mpix/s (pixels composited per second) with every input already resident in HBM: the batches the application
            submitted are captured on the device (b2dgpu_capture_*) and replayed K times, CUDA events around each replay.
`e2e`     : the same metric through bl_context_* with HOST buffers: per step the application's render calls +
            flush(BL_CONTEXT_FLUSH_SYNC) = record + serialise + H2D + kernels + D2H of the canvas into the BLImage.
`roofline`: the tile compositor (k_tile_render), algorithmic bytes = 8 B per composited pixel (4 read + 4 written,
            SURVEY 8d) / its CUDA-event duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`configs` : the other BASELINE.json configurations (0: bl_bench rects 512x600, 2: patterns + mixed operators, 3: 100 000
            glyphs, 4-ii: independent 1080p frames), each parity-gated against the reference at full size before timing.
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

import bench_scenes as BS                                      # noqa: E402
from bench_scenes import Scene, make_config1_scene             # noqa: E402,F401  (re-exported for the tests)

W4K, H4K = 3840, 2160
FLAG_DISABLE_JIT, FLAG_GPU = 0x1, 0x10000000
GPU_DRIVER = os.path.join(ROOT, "shim", "_build", "libgpu_scene_driver.so")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "libref_scene_driver.so")


# ---------------------------------------------------------------------------------------------------------------------
# The Blend2D application (shim/bl_scene_driver.cpp) bound to one of its two builds
# ---------------------------------------------------------------------------------------------------------------------
class Driver:
    def __init__(self, path):
        self.path = path
        self.lib = lib = C.CDLL(path)
        P, u32, vp = C.POINTER, C.c_uint32, C.c_void_p
        sig = {
            "bl_scene_open": [P(Scene), C.c_int, C.c_int, u32, u32, u32, u32, u32, P(vp)],
            "bl_scene_close": [vp], "bl_scene_clear": [vp], "bl_scene_draw": [vp, u32, u32], "bl_scene_flush": [vp, C.c_int],
            "bl_scene_step": [vp, u32, u32, P(C.c_double)], "bl_scene_error_flags": [vp],
            "bl_scene_read_pixels": [vp, vp, C.c_ssize_t], "bl_scene_checksum": [vp, P(u32)],
            "bl_scene_run_frames": [vp, u32, u32, u32, P(C.c_double), P(u32)],
            "bl_scene_run": [P(Scene), u32, u32, C.c_int, C.c_int, u32, u32, u32, u32, P(C.c_double), vp, C.c_ssize_t],
            "bl_scene_count_pixels": [P(Scene), u32, u32, C.c_int, C.c_int, P(C.c_uint64)],
        }
        for name, args in sig.items():
            fn = getattr(lib, name)
            fn.restype = u32
            fn.argtypes = args

    @staticmethod
    def load(path):
        return Driver(path) if os.path.exists(path) else None

    def open(self, scene, W, H, fmt=1, flags=FLAG_DISABLE_JIT, threads=0, queue_limit=0, device=0):
        return Session(self, scene, W, H, fmt, flags, threads, queue_limit, device)


def _ck(code, where):
    if code != 0:
        raise RuntimeError(f"{where} failed: BLResult 0x{code:08X}")


class Session:
    def __init__(self, drv, scene, W, H, fmt, flags, threads, queue_limit, device):
        self.drv, self.lib, self.scene, self.W, self.H, self.fmt = drv, drv.lib, scene, W, H, fmt
        self.h = C.c_void_p()
        _ck(self.lib.bl_scene_open(C.byref(scene), W, H, fmt, flags, threads, queue_limit, device, C.byref(self.h)), "bl_scene_open")

    def clear(self): _ck(self.lib.bl_scene_clear(self.h), "bl_scene_clear")
    def draw(self, first, count): _ck(self.lib.bl_scene_draw(self.h, first, count), "bl_scene_draw")
    def flush(self, sync=True): _ck(self.lib.bl_scene_flush(self.h, 1 if sync else 0), "bl_scene_flush")
    def error_flags(self): return int(self.lib.bl_scene_error_flags(self.h))

    def step(self, first, count):
        dt = C.c_double(0)
        _ck(self.lib.bl_scene_step(self.h, first, count, C.byref(dt)), "bl_scene_step")
        return dt.value

    def pixels(self):
        out = np.zeros((self.H, self.W), dtype=np.uint8 if self.fmt == 3 else np.uint32)
        _ck(self.lib.bl_scene_read_pixels(self.h, out.ctypes.data_as(C.c_void_p), out.strides[0]), "bl_scene_read_pixels")
        return out

    def run_frames(self, first_frame, frame_count, fills_per_frame):
        dt, ck = C.c_double(0), C.c_uint32(0)
        _ck(self.lib.bl_scene_run_frames(self.h, first_frame, frame_count, fills_per_frame, C.byref(dt), C.byref(ck)), "bl_scene_run_frames")
        return dt.value, ck.value

    def close(self):
        if self.h:
            self.lib.bl_scene_close(self.h)
            self.h = C.c_void_p()


def channel_diff(a, b):
    """(pixels that differ, maximum per-channel difference) - ImageUtils::diff_info semantics (imagediff.h:22-)."""
    if a.dtype == np.uint32:
        d = 0
        for s in (0, 8, 16, 24):
            d = max(d, int(np.abs(((a >> s) & 0xFF).astype(np.int16) - ((b >> s) & 0xFF).astype(np.int16)).max()))
    else:
        d = int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max())
    return int((a != b).sum()), d


def host_threads():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(32, n))            # BL_RUNTIME_MAX_THREAD_COUNT = 32 (blend2d/core/runtime.h:25)


def run_reference(scene, n_fills, W, H, steps, warmup, threads, fmt=1, return_pixels=False, count_pixels=False):
    """Times the reference's CPU renderer (portable pipeline, async MT) on fills [0, n_fills).  Returns a dict or None."""
    drv = Driver.load(REF_DRIVER)
    if drv is None:
        return None
    secs = (C.c_double * (steps + warmup))()
    pixels = np.zeros((H, W), dtype=np.uint8 if fmt == 3 else np.uint32) if return_pixels else None
    _ck(drv.lib.bl_scene_run(C.byref(scene), 0, n_fills, W, H, fmt, FLAG_DISABLE_JIT, threads, steps + warmup, secs,
                             pixels.ctypes.data_as(C.c_void_p) if return_pixels else None, pixels.strides[0] if return_pixels else 0), "bl_scene_run(reference)")
    timed = list(secs)[warmup:]
    total = sum(timed)
    out = {"fills_s": n_fills * steps / total, "ms_per_step": total / steps * 1e3, "threads": threads, "pixels": pixels}
    if count_pixels:
        px = C.c_uint64(0)
        _ck(drv.lib.bl_scene_count_pixels(C.byref(scene), 0, n_fills, W, H, C.byref(px)), "bl_scene_count_pixels")
        out["pixels_per_step"] = int(px.value)
        out["mpix_s"] = px.value * steps / total / 1e6
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
class GpuBench:
    """Shared state of the GPU arm: the application bound to the GPU-enabled Blend2D, libb2dgpu's process-wide
    statistics / capture entry points, the L2 flush buffer."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - blend2d_b200 has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        from blend2d_b200 import _native as N
        self.N, self.lib = N, N.lib
        self.drv = Driver.load(GPU_DRIVER)
        if self.drv is None:
            raise SystemExit(f"bench.py: {GPU_DRIVER} is missing - run `make -C shim` where /root/reference exists (there is no fallback)")
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
        self._flush_val = 0

    def flush_l2(self):
        self._flush_val = (self._flush_val + 1) & 0xFF
        self.flush_buf.fill_(self._flush_val)
        self.torch.cuda.synchronize()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def stats(self, reset=True):
        st = self.N.Stats()
        self.N.check(self.lib.b2dgpu_global_stats(C.byref(st), 1 if reset else 0), "b2dgpu_global_stats")
        return {k: getattr(st, k) for k, _ in st._fields_}

    def open(self, scene, W, H, queue_limit=0, fmt=1):
        return self.drv.open(scene, W, H, fmt, FLAG_GPU, 0, queue_limit, self.local_rank)

    # -----------------------------------------------------------------------------------------------------------
    def leg(self, name, scene, n_fills, W, H, steps, warmup, tol, parity_scene=None, parity_note=None, do_e2e=True, ref_threads=None):
        """One configuration: parity against the reference at full size, resident replay (value), end to end."""
        lib, N = self.lib, self.N
        out = {"workload": name, "canvas": [W, H], "fills_per_step": n_fills}

        # ---- render once with capture on: the application's batches stay resident in HBM ----
        cap_sess = self.open(scene, W, H, queue_limit=n_fills + 512)
        cap_sess.clear()
        self.stats(reset=True)
        N.check(lib.b2dgpu_capture_begin(), "capture_begin")
        cap_sess.draw(0, n_fills)
        cap_sess.flush(True)
        cap = C.c_void_p()
        N.check(lib.b2dgpu_capture_end(C.byref(cap)), "capture_end")
        st0 = self.stats(reset=True)
        px_per_step = int(st0["pixels_composited"])
        nb, nc = C.c_uint32(0), C.c_uint64(0)
        lib.b2dgpu_capture_info(cap, C.byref(nb), C.byref(nc))
        out["pixels_per_step"] = px_per_step
        out["captured"] = {"batches": nb.value, "commands": int(nc.value), "edges": int(st0["edges"])}
        if cap_sess.error_flags():
            raise SystemExit(f"bench.py: {name}: the context reported errors {cap_sess.error_flags():#x}")
        gpu_px = cap_sess.pixels()
        out["canvas_checksum"] = int(np.bitwise_xor.reduce(gpu_px.ravel()))

        # ---- parity gate (rank 0): the same calls through the unmodified reference, full size ----
        if self.rank == 0 and not self.args.no_parity:
            th = ref_threads or host_threads()
            if parity_scene is not None:
                s2 = self.open(parity_scene, W, H, queue_limit=0)
                s2.clear(); s2.draw(0, n_fills); s2.flush(True)
                cmp_gpu = s2.pixels()
                s2.close()
                ref = run_reference(parity_scene, n_fills, W, H, 1, 0, th, return_pixels=True)
            else:
                cmp_gpu = gpu_px
                ref = run_reference(scene, n_fills, W, H, 1, 0, th, return_pixels=True)
            if ref is None:
                out["parity"] = {"checked": False, "why": "oracle/_ref/libref_scene_driver.so missing"}
            else:
                n_diff, d = channel_diff(ref["pixels"], cmp_gpu)
                out["parity"] = {"checked": True, "pixels_differing": n_diff, "max_channel_diff": d, "tolerance": tol,
                                 "against": "unmodified reference (portable pipeline), same bl_* calls, full size" + (f"; {parity_note}" if parity_note else "")}
                out["reference_cpu"] = {"ms_per_step": ref["ms_per_step"], "fills_per_s": ref["fills_s"], "threads": th, "steps": 1,
                                        "mpix_s": None if parity_scene is not None else px_per_step / (ref["ms_per_step"] * 1e-3) / 1e6}
                if d > tol:
                    raise SystemExit(f"bench.py: {name}: PARITY FAILED - {n_diff} pixels differ from the reference, max channel diff {d} > {tol}; nothing is timed")

        # ---- value: resident replay ----
        ms = C.c_float(0)
        for _ in range(warmup):
            cap_sess.clear()
            N.check(lib.b2dgpu_capture_replay(cap, 1, C.byref(ms)), "capture_replay")
        self.barrier()
        self.stats(reset=True)
        N.check(lib.b2dgpu_global_set_profiling(1), "set_profiling")
        # pixels_per_step was measured by the capture pass; the statistic is not maintained inside the timed steps
        N.check(lib.b2dgpu_global_set_pixel_counting(0), "set_pixel_counting")
        step_ms = []
        for _ in range(steps):
            cap_sess.clear()
            self.stats(reset=False)
            self.flush_l2()
            N.check(lib.b2dgpu_capture_replay(cap, 1, C.byref(ms)), "capture_replay")
            step_ms.append(float(ms.value))
        self.barrier()
        st = self.stats(reset=True)
        N.check(lib.b2dgpu_global_set_profiling(0), "set_profiling")
        N.check(lib.b2dgpu_global_set_pixel_counting(1), "set_pixel_counting")
        # the clears between the steps are single solid fills (k_stream_solid): they add launches and pixels that are
        # not part of the step, so both are taken from the capture pass, which held exactly one step
        total_ms = float(sum(step_ms))
        out["resident"] = {"total_ms": total_ms, "steps": steps, "ms_per_step": total_ms / steps,
                           "launches_per_step": int(st0["kernel_launches"]),
                           "tile_kernel_ms": None, "build_kernels_ms": None}
        # per-kernel event times: the clear's launch is a "tile kernel" record too (stream path); subtract by taking the
        # records of the replays only: every step contributes 1 clear + nb batch records
        out["resident"]["tile_kernel_ms_sum"] = st["tile_kernel_ms"]
        out["resident"]["build_kernels_ms_sum"] = st["build_kernels_ms"]
        out["resident"]["profile_records"] = int(st["tile_kernel_launches"])
        N.check(lib.b2dgpu_capture_destroy(cap), "capture_destroy")
        cap_sess.close()

        # ---- e2e: bl_context_* with host buffers, flush(SYNC) inside the timed region ----
        if do_e2e:
            sess = self.open(scene, W, H, queue_limit=self.args.queue_limit)
            for _ in range(max(1, min(warmup, 2))):
                sess.clear(); sess.step(0, n_fills)
            self.barrier()
            e2e_steps = max(1, min(steps, 5))
            e2e_s, h2d, d2h, e2e_px = 0.0, 0, 0, 0
            for _ in range(e2e_steps):
                sess.clear()
                self.flush_l2()
                self.stats(reset=True)
                e2e_s += sess.step(0, n_fills)
                s_ = self.stats(reset=True)
                h2d += s_["h2d_bytes"]; d2h += s_["d2h_bytes"]; e2e_px += s_["pixels_composited"]
            self.barrier()
            out["e2e"] = {"seconds": e2e_s, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                          "pixels": e2e_px, "h2d_bytes_per_step": int(h2d / e2e_steps), "d2h_bytes_per_step": int(d2h / e2e_steps),
                          "host_checksum": int(np.bitwise_xor.reduce(sess.pixels().ravel()))}
            sess.close()
        return out


def tile_kernel_ms_per_step(gb, scene, n_fills, W, H, steps):
    """CUDA-event time of k_tile_render alone (and of the kernels before it) for a resident batch: a second capture
    replayed without the clears in between, so every profile record belongs to the step."""
    lib, N = gb.lib, gb.N
    sess = gb.open(scene, W, H, queue_limit=n_fills + 512)
    sess.clear()
    N.check(lib.b2dgpu_capture_begin(), "capture_begin")
    sess.draw(0, n_fills); sess.flush(True)
    cap = C.c_void_p()
    N.check(lib.b2dgpu_capture_end(C.byref(cap)), "capture_end")
    ms = C.c_float(0)
    N.check(lib.b2dgpu_capture_replay(cap, 2, C.byref(ms)), "capture_replay")
    gb.stats(reset=True)
    N.check(lib.b2dgpu_global_set_profiling(1), "set_profiling")
    N.check(lib.b2dgpu_global_set_pixel_counting(0), "set_pixel_counting")
    for _ in range(steps):
        gb.flush_l2()
        N.check(lib.b2dgpu_capture_replay(cap, 1, C.byref(ms)), "capture_replay")
    st = gb.stats(reset=True)
    N.check(lib.b2dgpu_global_set_profiling(0), "set_profiling")
    N.check(lib.b2dgpu_global_set_pixel_counting(1), "set_pixel_counting")
    N.check(lib.b2dgpu_capture_destroy(cap), "capture_destroy")
    sess.close()
    n = max(1, int(st["tile_kernel_launches"]))
    per_step = n / steps
    return st["tile_kernel_ms"] / n * per_step, st["build_kernels_ms"] / n * per_step


def run_gpu(args):
    gb = GpuBench(args)
    torch, dist, rank, world = gb.torch, gb.dist, gb.rank, gb.world
    W, H, n_fills = args.width, args.height, args.fills

    # ---- main leg: config 1.  Frame sharding: every rank renders its own frame (per-rank seed). ----
    scene, keep = make_config1_scene(n_fills, W, H, seed=1234 + rank)
    sampler = ClockSampler(gb.local_rank)
    if rank == 0:
        sampler.start()
    main = gb.leg(f"config1: {n_fills} fills/step (bl_bench polygons + tester quad/cubic paths, NonZero+EvenOdd, "
                  f"linear/radial/conic gradients pad/repeat/reflect, SrcOver) on {W}x{H} PRGB32",
                  scene, n_fills, W, H, args.steps, args.warmup, tol=1)
    clocks = sampler.stop() if rank == 0 else None
    tile_ms, build_ms = tile_kernel_ms_per_step(gb, scene, n_fills, W, H, max(2, min(args.steps, 5)))

    total_ms, px_per_step = main["resident"]["total_ms"], main["pixels_per_step"]
    e2e_s, e2e_steps, e2e_px = main["e2e"]["seconds"], main["e2e"]["steps"], main["e2e"]["pixels"] / main["e2e"]["steps"]
    launches = main["resident"]["launches_per_step"] * args.steps
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w_ = torch.tensor([px_per_step, e2e_px, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(w_, op=dist.ReduceOp.SUM)
        total_ms, e2e_s = float(t[0]), float(t[1])
        px_all, e2e_px_all, launches = float(w_[0]), float(w_[1]), int(w_[2])
    else:
        px_all, e2e_px_all = px_per_step, e2e_px

    # ---- the other configurations ----
    configs = {}
    if not args.no_configs:
        if world == 1:
            configs.update(other_config_legs(gb, args))
        configs["config4_frames"] = frames_leg(gb, args)

    full = None
    if rank == 0 and not args.no_full_canvas:
        full = measure_full_canvas(gb)

    band = None
    if not args.no_band:                                   # N = 1 is the anchor of the strong-scaling curve
        band = measure_band_sharded(gb, args)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg_bytes = px_per_step * 8.0
        achieved = alg_bytes / (tile_ms * 1e-3) / 1e9 if tile_ms > 0 else 0.0
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
            "config": {"workload": main["workload"], "fills_per_step": n_fills, "canvas": [W, H], "frames": world,
                       "sharding": "frame-sharded, one frame per GPU, no collective" if world > 1 else "single frame",
                       "l2": "256 MiB buffer written between timed steps (canvas + batch evicted from L2)",
                       "api": "Blend2D C API (bl_context_*) of shim/_build/libblend2d_gpu.so, BLContextCreateInfo.flags = 0x10000000"},
            "e2e": {"value": e2e_px_all * e2e_steps / e2e_s / 1e6, "unit": "Mpix/s",
                    "h2d_bytes_per_step": main["e2e"]["h2d_bytes_per_step"], "d2h_bytes_per_step": main["e2e"]["d2h_bytes_per_step"],
                    "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                    "path": "bl_context_fill_* (reference frontend) -> shim consume_batch -> b2dgpu_submit(host batch) -> kernels -> "
                            "flush(BL_CONTEXT_FLUSH_SYNC) -> b2dgpu_target_download(BLImage pixels)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_tile_render<4>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": tile_ms, "other_kernels_ms": build_ms,
                         "peak_source": peak_src,
                         "note": "tile-resident compositing: the destination tile is read and written once per step, so "
                                 "algorithmic bytes (8 B per composited pixel) exceed DRAM traffic by the overdraw factor"},
            "clocks": clocks,
            "pixels_per_step": px_per_step, "canvas_checksum": main["canvas_checksum"],
            "parity": main.get("parity"),
            "rasterizer": {"edges_per_step": main["captured"]["edges"], "batches_per_step": main["captured"]["batches"],
                           "edge_builder_and_binning_ms": build_ms,
                           "edges_per_s": main["captured"]["edges"] / (build_ms * 1e-3) if build_ms > 0 else None,
                           "note": "K1 = edge count + scan + edge write + bbox + binning (CUDA events around them); "
                                   "issue-slot utilisation of K2/K3 (k_tile_render) is in profiles/*.summary.txt"},
            "configs": configs,
            "band_sharded": band,
            "roofline_full_canvas": full if full is None else {
                k: {"bound": "hbm", "kernel": v["kernel"], "achieved": v["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": v["gbs"] / peak, "kernel_ms": v["ms"], "algorithmic_bytes_per_launch": v["bytes"],
                    "workload": v["what"]} for k, v in full.items()},
        }
        if world == 1 and not args.no_cpu_baseline and main.get("reference_cpu"):
            rc = main["reference_cpu"]
            out["cpu_baseline"] = {"value": rc["mpix_s"], "unit": "Mpix/s", "cores": rc["threads"], "kind": "reference",
                                   "sample": f"the whole step ({n_fills} fills of the same scene), 1 step, reference built from /root/reference "
                                             f"(portable non-JIT pipeline: asmjit is not vendored), BLContextCreateInfo.thread_count={rc['threads']}",
                                   "fills_per_s": rc["fills_per_s"], "ms": rc["ms_per_step"]}
        print(json.dumps(out))

    if world > 1:
        dist.destroy_process_group()


def summarize_leg(leg):
    r, px = leg["resident"], leg["pixels_per_step"]
    out = {"workload": leg["workload"], "canvas": leg["canvas"], "fills_per_step": leg["fills_per_step"],
           "value_mpix_s": px / (r["ms_per_step"] * 1e-3) / 1e6, "value_fills_per_s": leg["fills_per_step"] / (r["ms_per_step"] * 1e-3),
           "ms_per_step": r["ms_per_step"], "steps": r["steps"], "pixels_per_step": px, "launches_per_step": r["launches_per_step"],
           "captured": leg["captured"], "parity": leg.get("parity"), "reference_cpu": leg.get("reference_cpu")}
    if "e2e" in leg:
        e = leg["e2e"]
        out["e2e"] = {"mpix_s": e["pixels"] / e["seconds"] / 1e6, "fills_per_s": leg["fills_per_step"] * e["steps"] / e["seconds"],
                      "ms_per_step": e["ms_per_step"], "h2d_bytes_per_step": e["h2d_bytes_per_step"], "d2h_bytes_per_step": e["d2h_bytes_per_step"]}
    return out


def other_config_legs(gb, args):
    out = {}
    steps, warmup = max(2, min(args.steps, 5)), 3
    # config 0: bl_bench FillRectA / FillRectU on 512x600
    sc, keep = BS.make_config0_scene(args.config0_fills)
    out["config0_rects"] = summarize_leg(gb.leg(
        f"config0: {args.config0_fills} bl_bench FillRectA/FillRectU (alternating), solid random-alpha colours, SrcOver, 8..256 px on 512x600 PRGB32",
        sc, args.config0_fills, 512, 600, steps, warmup, tol=0))
    # config 2: FillRectRot / FillRoundU with patterns, mixed operators
    sc, keep = BS.make_config2_scene(args.config2_fills, W4K, H4K)
    psc, pkeep = BS.slice_fills(sc, keep, {BS.PLUS: BS.SRC_OVER, BS.MULTIPLY: BS.SRC_COPY, BS.SCREEN: BS.SRC_OVER})
    out["config2_patterns"] = summarize_leg(gb.leg(
        f"config2: {args.config2_fills} FillRectRot/FillRoundU, 64x64 PRGB32 sprite REPEAT, nearest+bilinear, SrcCopy/Plus/Multiply/Screen on {W4K}x{H4K} PRGB32",
        sc, args.config2_fills, W4K, H4K, steps, warmup, tol=0, parity_scene=psc,
        parity_note="the reference's portable pipeline has no Plus/Multiply/Screen: parity on the same geometry and patterns with those operators "
                    "remapped to SrcOver/SrcCopy; Plus is checked in tests/ against the reference's own CompOp_Plus_Op template (oracle/ref_internals.cpp), Multiply / Screen against the C restatement (unpinned)"))
    # config 3: 100 000 glyphs
    sc, keep = BS.make_config3_scene(args.config3_strings, W4K, H4K)
    out["config3_glyphs"] = summarize_leg(gb.leg(
        f"config3: {args.config3_strings} fill_utf8_text calls x 4 characters = {args.config3_strings * 4} glyphs, ABeeZee 20 px, solid colours, SrcOver on {W4K}x{H4K} PRGB32",
        sc, args.config3_strings, W4K, H4K, steps, warmup, tol=0))
    out["config3_glyphs"]["glyphs_per_s"] = out["config3_glyphs"]["value_fills_per_s"] * 4
    return out


def frames_leg(gb, args):
    """Config 4 (ii): independent 1080p frames, frame-sharded: frame f -> rank f mod N.  Every frame is cleared, drawn,
    flushed (SYNC) and consumed on the host.  No collective on the data path; time = max over ranks, work = sum."""
    torch, dist, rank, world = gb.torch, gb.dist, gb.rank, gb.world
    FW, FH, k = 1920, 1080, args.frame_fills
    total_frames = args.frames
    mine = list(range(rank, total_frames, world))
    # each rank generates only its own frames (frame f has seed base + f), in chunks to bound host memory
    chunk = 512
    secs, checksum, frames_done, px = 0.0, 0, 0, 0
    parity = None
    resident = None
    for c0 in range(0, len(mine), chunk):
        ids = mine[c0:c0 + chunk]
        sc, keep = make_frames_for(ids, k, FW, FH)
        sess = gb.open(sc, FW, FH)
        if c0 == 0:
            # parity gate on the first frames of this rank + resident replay of a captured subset
            if rank == 0 and not args.no_parity:
                nchk = min(4, len(ids))
                worst, ndiff = 0, 0
                for f in range(nchk):
                    sess.clear(); sess.draw(f * k, k); sess.flush(True)
                    g = sess.pixels()
                    one = BS.Scene.from_buffer_copy(sc)
                    ref = run_reference_range(one, f * k, k, FW, FH)
                    if ref is None:
                        break
                    n_, d_ = channel_diff(ref, g)
                    worst, ndiff = max(worst, d_), ndiff + n_
                else:
                    parity = {"checked": True, "frames": nchk, "pixels_differing": ndiff, "max_channel_diff": worst, "tolerance": 0}
                    if worst > 0:
                        raise SystemExit(f"bench.py: config4 frames: PARITY FAILED (max channel diff {worst})")
            ncap = min(args.frames_resident, len(ids))
            gb.stats(reset=True)
            gb.N.check(gb.lib.b2dgpu_capture_begin(), "capture_begin")
            for f in range(ncap):
                sess.draw(f * k, k); sess.flush(False)
            sess.flush(True)
            cap = C.c_void_p()
            gb.N.check(gb.lib.b2dgpu_capture_end(C.byref(cap)), "capture_end")
            st = gb.stats(reset=True)
            ms = C.c_float(0)
            gb.N.check(gb.lib.b2dgpu_capture_replay(cap, 1, C.byref(ms)), "capture_replay")
            gb.flush_l2()
            gb.N.check(gb.lib.b2dgpu_capture_replay(cap, 1, C.byref(ms)), "capture_replay")
            resident = {"frames": ncap, "ms": float(ms.value), "pixels": int(st["pixels_composited"]), "launches": int(st["kernel_launches"])}
            gb.N.check(gb.lib.b2dgpu_capture_destroy(cap), "capture_destroy")
            sess.run_frames(0, min(8, len(ids)), k)                # warm-up
            gb.barrier()
        gb.stats(reset=True)
        dt, ck = sess.run_frames(0, len(ids), k)
        st = gb.stats(reset=True)
        secs += dt; checksum ^= ck; frames_done += len(ids); px += int(st["pixels_composited"])
        sess.close()
    gb.barrier()
    res_v = [resident["pixels"] / (resident["ms"] * 1e-3), resident["frames"] / (resident["ms"] * 1e-3), float(resident["launches"])]
    if world > 1:
        t = torch.tensor([secs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w_ = torch.tensor([float(px), float(frames_done)] + res_v, dtype=torch.float64, device="cuda")
        dist.all_reduce(w_, op=dist.ReduceOp.SUM)
        secs, px, frames_done, res_v = float(t[0]), float(w_[0]), float(w_[1]), [float(w_[2]), float(w_[3]), float(w_[4])]
    return {"workload": f"config4(ii): {total_frames} independent {FW}x{FH} PRGB32 frames x {k} fills (polygons, quad/cubic paths, linear gradients + solid), "
                        f"frame f -> GPU f mod {world}; every frame cleared, drawn, flush(SYNC) into the BLImage's host pixels; every 32nd frame checksummed by the CPU",
            "frames": int(frames_done), "fills_per_frame": k, "n_gpus": world, "scaling": "strong (fixed batch of frames)",
            "e2e": {"seconds_max_over_ranks": secs, "frames_per_s": frames_done / secs, "fills_per_s": frames_done * k / secs, "mpix_s": px / secs / 1e6,
                    "d2h_bytes_per_frame": FW * FH * 4},
            "resident": {"frames_captured_per_gpu": resident["frames"], "mpix_s": res_v[0] / 1e6, "frames_per_s": res_v[1],
                         "fills_per_s": res_v[1] * k, "launches": int(res_v[2]),
                         "note": "device-resident replay of the captured frames of each rank (inputs in HBM), summed over ranks"},
            "parity": parity, "host_checksum": int(checksum), "collective": "none"}


def make_frames_for(frame_ids, k, W, H, base_seed=100):
    """Scene holding the given frames back to back (frame f is generated from seed base + f, whatever rank draws it)."""
    return BS.make_frames_scene_ids(frame_ids, k, W, H, base_seed)


def run_reference_range(scene, first, count, W, H):
    drv = Driver.load(REF_DRIVER)
    if drv is None:
        return None
    s = drv.open(scene, W, H, 1, FLAG_DISABLE_JIT, 0, 0, 0)
    s.clear(); s.draw(first, count); s.flush(True)
    px = s.pixels()
    s.close()
    return px


def measure_full_canvas(gb):
    """One full-canvas SrcOver fill per launch: reads 4 B and writes 4 B per pixel; a PRGB32 pattern source
    adds 4 B per pixel.  Solid -> k_stream_solid; gradient / pattern -> k_stream_one (stream.cu)."""
    import blend2d_b200 as G
    N, lib, torch = gb.N, gb.lib, gb.torch
    # The runtime runs on a torch stream and the L2 flush is enqueued on the same stream right before every launch, so
    # the kernel's events are not inflated by the launch latency of an idle GPU (a 4K fill lasts ~17 us).
    stream = torch.cuda.Stream()
    rt = G.Runtime(device=gb.local_rank, stream=stream.cuda_stream)
    out = {}
    tex = None
    for name, (W, H), style in (("4k", (3840, 2160), "solid"), ("16k", (16384, 16384), "solid"),
                                 ("16k_linear", (16384, 16384), "linear"), ("16k_radial", (16384, 16384), "radial"),
                                 ("16k_conic", (16384, 16384), "conic"), ("16k_pattern", (16384, 16384), "pattern")):
        img = G.Image(W, H, G.FORMAT_PRGB32)
        rec = G.Context(img, record_only=True)
        extra = 0.0
        if style == "solid":
            rec.set_fill_style(0x80336699)
            kernel = "k_stream_solid<SrcOver>"
        elif style == "linear":
            rec.set_fill_style(G.Gradient(G.GRADIENT_LINEAR, [0.0, 0.0, float(W), float(H)], G.EXTEND_PAD,
                                          [(0.0, 0x80FF0000), (0.5, 0xC000FF00), (1.0, 0x800000FF)]))
            kernel = "k_stream_one<linear> (stream.cu)"
        elif style == "radial":
            rec.set_fill_style(G.Gradient(G.GRADIENT_RADIAL, [W / 2.0, H / 2.0, W / 2.0 - 100.0, H / 2.0 - 50.0, W / 2.0, 0.0], G.EXTEND_PAD,
                                          [(0.0, 0x80FF0000), (0.5, 0xC000FF00), (1.0, 0x800000FF)]))
            kernel = "k_stream_one<radial> (stream.cu)"
        elif style == "conic":
            rec.set_fill_style(G.Gradient(G.GRADIENT_CONIC, [W / 2.0, H / 2.0, 0.0, 1.0], G.EXTEND_PAD,
                                          [(0.0, 0x80FF0000), (0.33, 0xC000FF00), (0.66, 0x800000FF), (1.0, 0x80FF0000)]))
            kernel = "k_stream_one<conic> (stream.cu)"
        else:
            tex = G.Image(4096, 4096, G.FORMAT_PRGB32)
            rng = np.random.default_rng(1)
            a = rng.integers(1, 255, (4096, 4096)).astype(np.uint32)
            tex.from_numpy((a << 24) | ((a // 2) << 16) | ((a // 3) << 8) | (a // 4))
            rec.set_fill_style(G.Pattern(tex, None, G.EXTEND_REPEAT, [1, 0, 0, 1, 0, 0]))
            kernel = "k_stream_one<pattern32> (stream.cu), PRGB32 pattern aligned + repeat, 64 MiB source"
            extra = 4.0
        rec.fill_all()
        batch = G.ResidentBatch(rt._h, rec.peek_batch())
        tgt = C.c_void_p()
        N.check(lib.b2dgpu_target_create(rt._h, W, H, G.FORMAT_PRGB32, C.byref(tgt)), "target_create")
        for _ in range(3):
            batch.render(tgt)
        N.check(lib.b2dgpu_sync(rt._h), "sync")
        rt.stats(reset=True)
        N.check(lib.b2dgpu_set_profiling(rt._h, 1), "set_profiling")
        reps = 10
        with torch.cuda.stream(stream):
            for _ in range(reps):
                gb.flush_buf.fill_(7)
                batch.render(tgt)
        N.check(lib.b2dgpu_sync(rt._h), "sync")
        st = rt.stats(reset=True)
        N.check(lib.b2dgpu_set_profiling(rt._h, 0), "set_profiling")
        ms = st["tile_kernel_ms"] / reps
        nbytes = W * H * (8.0 + extra)
        out[name] = {"ms": ms, "bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9, "kernel": kernel,
                     "what": f"full-canvas SrcOver {style} fill of a {W}x{H} PRGB32 canvas, L2 flushed between launches"}
        batch.close()
        N.check(lib.b2dgpu_target_destroy(tgt), "target_destroy")
    return out


def measure_band_sharded(gb, args):
    """Config 4(i): ONE large canvas cut into tile-aligned stripes of rows, interleaved over the GPUs (band
    sharding).  Every rank replays the whole command list clipped to its stripes (b2dgpu_target_create_slab); the only
    exchange is the gather of the stripes into rank 0's image.  Strong scaling: the frame is fixed."""
    import blend2d_b200 as G
    from blend2d_b200 import sharding as SH
    N, lib, torch, dist = gb.N, gb.lib, gb.torch, gb.dist
    rank, world, local_rank = gb.rank, gb.world, gb.local_rank
    side, n_fills = args.band_canvas, args.band_fills
    scene, keep = make_config1_scene(n_fills, side, side, seed=4321)           # the same frame on every rank
    stream = torch.cuda.Stream()
    rt = G.Runtime(device=local_rank, stream=stream.cuda_stream)
    rec = G.Context(G.Image(side, side, G.FORMAT_PRGB32), record_only=True)    # host image: clip box only, never touched
    N.check(lib.b2d_scene_replay(rec._h, C.byref(scene), 0, n_fills), "b2d_scene_replay(record)")
    batch = G.ResidentBatch(rt._h, rec.peek_batch())
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(2):
        for t_ in tgts:
            N.check(lib.b2dgpu_target_clear(t_), "clear")
        render_all()
    barrier()
    rt.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for t_ in tgts:
        N.check(lib.b2dgpu_target_clear(t_), "clear")
    torch.cuda.synchronize()
    e0.record(stream); render_all(); e1.record(stream)
    torch.cuda.synchronize()
    st = rt.stats(reset=True)

    class _Mem:
        pass

    def view_of(t_, rows):
        ptr, stride, pw, ph = C.c_void_p(), C.c_ssize_t(), C.c_int32(), C.c_int32()
        N.check(lib.b2dgpu_target_device_view(t_, C.byref(ptr), C.byref(stride), C.byref(pw), C.byref(ph)), "device_view")
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (ph.value, stride.value), "typestr": "|u1", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(m, device=torch.device("cuda", local_rank))[:rows, : side * 4]
    local = [view_of(t_, y1 - y0) for t_, (y0, y1) in zip(tgts, stripes)]

    # The exchange: every stripe goes straight to its rows of rank 0's final image (no staging, no concatenation), and
    # the render of stripe j + 1 overlaps the transfer of stripe j (sharding.StripeGather).
    gather = SH.StripeGather(side, k, rank, world, torch.device("cuda", local_rank))
    side_stream = torch.cuda.Stream(priority=-1)
    t_render, t_gather, t_overlap = e0.elapsed_time(e1), None, None
    if world > 1:
        gather.run(local)                                                       # warm-up: NCCL channel setup
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            g0.record(stream)
            gather.run(local)
            g1.record(stream)
        torch.cuda.synchronize()
        t_gather = g0.elapsed_time(g1)
        # overlapped: per stripe render -> send, the transfers run on the side stream while the next stripe renders
        for t_ in tgts:
            N.check(lib.b2dgpu_target_clear(t_), "clear")
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        gather.begin()
        render_all()                                                        # ONE geometry pass, the stripes composited in order
        for j, t_ in enumerate(tgts):
            # the side stream waits for stripe j only (b2dgpu_target_wait), its transfer runs while j + 1 .. are composited
            N.check(lib.b2dgpu_target_wait(t_, side_stream.cuda_stream), "target_wait")
            gather.stripe_ready(j, local[j], side_stream)
        gather.finish(side_stream)
        stream.wait_stream(side_stream)
        o1.record(stream)
        torch.cuda.synchronize()
        t_overlap = o0.elapsed_time(o1)
    vals = [t_render, t_gather or 0.0, t_overlap or 0.0]
    px = [float(st["pixels_composited"])]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(v) for v in t]
        p = torch.tensor(px, dtype=torch.float64, device="cuda")
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
        px = [float(p[0])]
    out = None
    if rank == 0:
        out = {"workload": f"config4(i): {n_fills} fills on one {side}x{side} PRGB32 canvas, band-sharded into {world} x {k} interleaved stripes of rows",
               "n_gpus": world, "render_ms_max_over_ranks": vals[0], "value": px[0] / (vals[0] * 1e-3) / 1e6, "unit": "Mpix/s", "scaling": "strong",
               "gather_ms": vals[1] if world > 1 else None, "render_plus_gather_overlapped_ms": vals[2] if world > 1 else None,
               "gathered_bytes": side * side * 4 if world > 1 else 0,
               "collective": "per-stripe ncclSend/ncclRecv straight into rank 0's image rows on a side stream that waits per stripe (b2dgpu_target_wait): one geometry pass, transfer of stripe j under the compositing of j + 1 .." if world > 1 else "none (single GPU anchor)"}
    batch.close()
    for t_ in tgts:
        N.check(lib.b2dgpu_target_destroy(t_), "target_destroy")
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, H, n = args.width, args.height, args.fills
    scene, keep = make_config1_scene(n, W, H, seed=1234)
    threads = host_threads()
    ref = run_reference(scene, n, W, H, args.steps, args.warmup, threads, count_pixels=True)
    if ref is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_scene_driver.so not built (needs /root/reference)"}))
        return
    out = {
        "impl": "reference", "metric": "Mpix/s", "unit": "Mpix/s", "value": ref["mpix_s"], "fills_per_s": ref["fills_s"],
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"config1: {n} fills/step (bl_bench polygons + tester quad/cubic paths, NonZero+EvenOdd, "
                               f"linear/radial/conic gradients pad/repeat/reflect, SrcOver) on {W}x{H} PRGB32",
                   "fills_per_step": n, "canvas": [W, H], "frames": 1, "pixels_per_step": ref["pixels_per_step"],
                   "api": "Blend2D C API (bl_context_*) of the unmodified reference, BL_CONTEXT_CREATE_FLAG_DISABLE_JIT"},
        "cpu_baseline": {"value": ref["mpix_s"], "unit": "Mpix/s", "cores": threads, "kind": "reference",
                         "sample": f"the whole step ({n} fills), portable (non-JIT) pipeline, async rendering with thread_count={threads}"},
        "e2e": {"value": ref["mpix_s"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-band", action="store_true", help="skip the band-sharded 16384^2 measurement")
    ap.add_argument("--band-single", action="store_true", help="(kept for old command lines: the band-sharded measurement runs at N = 1 too by default)")
    ap.add_argument("--band-canvas", type=int, default=16384)
    ap.add_argument("--band-fills", type=int, default=600)
    ap.add_argument("--band-stripes", type=int, default=8, help="interleaved stripes per GPU in the band-sharded measurement")
    ap.add_argument("--queue-limit", type=int, default=0, help="BLContextCreateInfo.command_queue_limit on the e2e path; 0 = the shim's default (2048)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fills", type=int, default=10000)
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size comparison with the reference before timing")
    ap.add_argument("--no-full-canvas", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs of configs 0, 2, 3 and 4(ii)")
    ap.add_argument("--config0-fills", type=int, default=20000)
    ap.add_argument("--config2-fills", type=int, default=10000)
    ap.add_argument("--config3-strings", type=int, default=25000)
    ap.add_argument("--frames", type=int, default=8192, help="config 4(ii): frames in the batch (all ranks together)")
    ap.add_argument("--frame-fills", type=int, default=96)
    ap.add_argument("--frames-resident", type=int, default=256, help="config 4(ii): frames per GPU captured for the device-resident number")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
