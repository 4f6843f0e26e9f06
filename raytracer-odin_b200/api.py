"""`Renderer`: the call sequence the Odin shim makes through `foreign import`,
expressed in Python over the same C ABI.  Mirrors the reference's hot-path interface:

    finish_scene(rc, &scene)            -> Scene.finish(native_bvh_build)        raytracer.odin:62
    render_scene(rc, &scene, trials)    -> Renderer.render_scene(...)            raytracer.odin:602
    rc.pixels[0]                        -> numpy array of Sample_Stats           main.odin:34-40

No CPU fallback: constructing a Renderer without the CUDA library or without a B200-class GPU
raises RuntimeError.
"""
import ctypes as C
import statistics
import time
from typing import Optional

import numpy as np

from . import cabi
from .scene import Scene


class OrtError(RuntimeError):
    pass


class Renderer:
    def __init__(self, device: int = 0, seed: int = 0, max_paths_in_flight: int = 0, max_path_bytes: int = 0):
        self.lib = cabi.load_library()
        self._ctx = C.c_void_p()
        cfg = cabi.OrtDeviceCfg(device=device, seed=seed, max_paths_in_flight=max_paths_in_flight,
                                max_path_bytes=max_path_bytes)
        if self.lib.ort_create(C.byref(self._ctx), C.byref(cfg)) != 0:
            raise OrtError(self.lib.ort_last_error(None).decode())
        self.device = device
        self.scene: Optional[Scene] = None

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if self._ctx:
            self.lib.ort_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise OrtError(self.lib.ort_last_error(self._ctx).decode())

    # -- scene ---------------------------------------------------------------------------------
    def upload_scene(self, scene: Scene):
        cs, keep = scene.to_c()
        self._check(self.lib.ort_upload_scene(self._ctx, C.byref(cs)))
        del keep  # the library deep-copies; nothing of ours is retained
        self.scene = scene
        return self

    def set_stream(self, cuda_stream):
        """cuda_stream: a cudaStream_t handle as int (0 = legacy default stream, e.g.
        torch.cuda.current_stream().cuda_stream), or None for the context's own stream."""
        h = C.c_void_p(-1) if cuda_stream is None else C.c_void_p(cuda_stream)
        self._check(self.lib.ort_set_stream(self._ctx, h))

    def set_profiling(self, on: bool):
        self._check(self.lib.ort_set_profiling(self._ctx, 1 if on else 0))

    # -- rendering -----------------------------------------------------------------------------
    def render(self, width: int, height: int, ray_depth: int, n_samples: int, first_sample: int = 0,
               out: Optional[np.ndarray] = None, interrupt: Optional[np.ndarray] = None) -> np.ndarray:
        """One blocking trial into `out` (Sample_Stats[h*w], accumulated like rc.pixels[0])."""
        if out is None:
            out = np.zeros(width * height, cabi.STATS_DTYPE)
        assert out.dtype == cabi.STATS_DTYPE and out.size == width * height and out.flags.c_contiguous
        iptr = interrupt.ctypes.data_as(C.c_void_p) if interrupt is not None else None
        self._check(self.lib.ort_render(self._ctx, width, height, ray_depth, first_sample, n_samples,
                                        cabi.ptr(out), iptr))
        return out

    def render_scene(self, width, height, ray_depth, n_samples, number_of_trials=1, out=None, log=print):
        """render_scene (raytracer.odin:602-665): trials accumulate into the same pixels without
        clearing and replay the same sample indices (:606-610); prints the reference's summary."""
        if out is None:
            out = np.zeros(width * height, cabi.STATS_DTYPE)
        timings = []
        for trial in range(number_of_trials):
            t0 = time.perf_counter()
            self.render(width, height, ray_depth, n_samples, 0, out)
            dt = time.perf_counter() - t0
            timings.append(dt)
            if log:
                log(f"Trial {trial} >>> Rendered in {dt * 1e3:.3f}ms")
        if number_of_trials > 1 and log:
            ts = sorted(timings)
            mean = statistics.fmean(ts)
            sd = statistics.stdev(ts)
            med = (ts[len(ts) // 2] + ts[(len(ts) + 1) // 2 if (len(ts) + 1) // 2 < len(ts) else -1]) / 2
            log(">>>>>>>>> Performance Summary <<<<<<<<<")
            log(f"Trials: {number_of_trials}")
            log(f"Time: {mean * 1e3:.02f}±{sd * 1e3:.02f}ms")
            log(f"Best: {ts[0] * 1e3:.02f}ms, Median: {med * 1e3:.02f}ms, Worst: {ts[-1] * 1e3:.02f}ms")
            log(">>>>>>>>> Performance Summary <<<<<<<<<")
        return out, timings

    def render_device(self, width, height, ray_depth, first_sample, n_samples, d_accum_ptr: int):
        """Asynchronous render into a device accumulator (8 planes x h*w floats)."""
        self._check(self.lib.ort_render_device(self._ctx, width, height, ray_depth, first_sample, n_samples,
                                               C.c_void_p(d_accum_ptr)))

    def unpack_accum(self, width, height, d_accum_ptr: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.zeros(width * height, cabi.STATS_DTYPE)
        self._check(self.lib.ort_unpack_accum(self._ctx, width, height, C.c_void_p(d_accum_ptr), cabi.ptr(out)))
        return out

    def tonemap_rgb8(self, width, height, d_accum_ptr: int) -> np.ndarray:
        out = np.zeros((height, width, 3), np.uint8)
        self._check(self.lib.ort_tonemap_rgb8(self._ctx, width, height, C.c_void_p(d_accum_ptr), cabi.ptr(out)))
        return out

    def last_render_samples(self) -> int:
        return int(self.lib.ort_last_render_samples(self._ctx))

    # -- device-resident frame (--continious / live preview): accumulators stay in HBM across calls --------
    def frame_begin(self, width: int, height: int):
        self._check(self.lib.ort_frame_begin(self._ctx, width, height))
        self._frame = (width, height)
        return self

    def frame_load(self, pixels: np.ndarray):
        pixels = np.ascontiguousarray(pixels, cabi.STATS_DTYPE)
        assert pixels.size == self._frame[0] * self._frame[1]
        self._check(self.lib.ort_frame_load(self._ctx, cabi.ptr(pixels)))

    def frame_render(self, ray_depth: int, first_sample: int, n_samples: int, interrupt: Optional[np.ndarray] = None) -> int:
        """Enqueue samples [first_sample, first_sample + n_samples); returns the number enqueued."""
        done = C.c_uint64()
        iptr = interrupt.ctypes.data_as(C.c_void_p) if interrupt is not None else None
        self._check(self.lib.ort_frame_render(self._ctx, ray_depth, first_sample, n_samples, iptr, C.byref(done)))
        return int(done.value)

    def frame_wait(self):
        self._check(self.lib.ort_frame_wait(self._ctx))

    def frame_snapshot(self):
        self._check(self.lib.ort_frame_snapshot(self._ctx))

    def frame_preview_rgb8(self) -> np.ndarray:
        w, h = self._frame
        out = np.zeros((h, w, 3), np.uint8)
        self._check(self.lib.ort_frame_preview_rgb8(self._ctx, cabi.ptr(out)))
        return out

    def frame_fetch(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        w, h = self._frame
        if out is None:
            out = np.zeros(w * h, cabi.STATS_DTYPE)
        self._check(self.lib.ort_frame_fetch(self._ctx, cabi.ptr(out)))
        return out

    def frame_end(self):
        self._check(self.lib.ort_frame_end(self._ctx))

    # -- parity probes -------------------------------------------------------------------------
    def probe_shading(self, kind: str, records: np.ndarray) -> np.ndarray:
        """n evaluations of a shading device function (cabi.PROBE); u32 inputs are passed as bit patterns."""
        code, ni, no = cabi.PROBE[kind]
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, ni)
        out = np.zeros((len(rec), no), np.float32)
        self._check(self.lib.ort_probe_shading(self._ctx, code, cabi.ptr(rec), len(rec), cabi.ptr(out)))
        return out

    def trace_rays(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, cabi.RAY_DTYPE)
        out = np.zeros(len(rays), cabi.HIT_DTYPE)
        self._check(self.lib.ort_trace_rays(self._ctx, cabi.ptr(rays), len(rays), cabi.ptr(out)))
        return out

    def light_pdf(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, cabi.RAY_DTYPE)
        out = np.zeros(len(rays), np.float32)
        self._check(self.lib.ort_light_pdf(self._ctx, cabi.ptr(rays), len(rays), cabi.ptr(out)))
        return out

    def primary_hits(self, width, height, sample=0, want_rays=False):
        out = np.zeros(width * height, cabi.HIT_DTYPE)
        rays = np.zeros(width * height, cabi.RAY_DTYPE) if want_rays else None
        self._check(self.lib.ort_primary_hits(self._ctx, width, height, sample, cabi.ptr(out),
                                              cabi.ptr(rays) if want_rays else None))
        return (out, rays) if want_rays else out

    def bench_trace(self, rays: np.ndarray, mode: int = 0, iters: int = 10) -> float:
        """Mean device time (ms) of one traversal-kernel launch over `rays` (diagnostic)."""
        rays = np.ascontiguousarray(rays, cabi.RAY_DTYPE)
        ms = C.c_double()
        self._check(self.lib.ort_bench_trace(self._ctx, cabi.ptr(rays), len(rays), mode, iters, C.byref(ms)))
        return ms.value

    def bench_read_bw(self, nbytes: int, iters: int = 20) -> float:
        """Streaming-read bandwidth (GB/s) of a working set of `nbytes` (L2 peak for 32-96 MB, HBM for GBs)."""
        g = C.c_double()
        self._check(self.lib.ort_bench_read_bw(self._ctx, nbytes, iters, C.byref(g)))
        return g.value

    # -- stats ---------------------------------------------------------------------------------
    def stats(self) -> dict:
        s = cabi.OrtStats()
        self._check(self.lib.ort_get_stats(self._ctx, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._check(self.lib.ort_reset_stats(self._ctx))


class MultiRenderer:
    """Several GPUs of one box driven from this one process (ort_multi_*): full scene replica per
    GPU, contiguous sample blocks, one peer-memory reduce per render call."""

    def __init__(self, devices, seed: int = 0):
        self.lib = cabi.load_library()
        self._m = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices)
        if self.lib.ort_multi_create(C.byref(self._m), devs, len(devices), seed) != 0:
            raise OrtError(self.lib.ort_multi_last_error(None).decode())
        self.devices = list(devices)

    def close(self):
        if self._m:
            self.lib.ort_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise OrtError(self.lib.ort_multi_last_error(self._m).decode())

    def upload_scene(self, scene: Scene):
        cs, keep = scene.to_c()
        self._check(self.lib.ort_multi_upload_scene(self._m, C.byref(cs)))
        return self

    def render(self, width, height, ray_depth, n_samples, first_sample=0, out=None, interrupt=None):
        if out is None:
            out = np.zeros(width * height, cabi.STATS_DTYPE)
        iptr = interrupt.ctypes.data_as(C.c_void_p) if interrupt is not None else None
        self._check(self.lib.ort_multi_render(self._m, width, height, ray_depth, first_sample, n_samples,
                                              cabi.ptr(out), iptr))
        return out

    def stats(self) -> dict:
        s = cabi.OrtStats()
        self._check(self.lib.ort_multi_get_stats(self._m, C.byref(s)))
        return s.as_dict()

    def last_render_samples(self) -> int:
        return int(self.lib.ort_multi_last_render_samples(self._m))

    # -- device-resident frame on all GPUs --------------------------------------------------------
    def frame_begin(self, width: int, height: int):
        self._check(self.lib.ort_multi_frame_begin(self._m, width, height))
        self._frame = (width, height)
        return self

    def frame_load(self, pixels: np.ndarray):
        pixels = np.ascontiguousarray(pixels, cabi.STATS_DTYPE)
        self._check(self.lib.ort_multi_frame_load(self._m, cabi.ptr(pixels)))

    def frame_render(self, ray_depth: int, first_sample: int, n_samples: int, interrupt=None) -> int:
        done = C.c_uint64()
        iptr = interrupt.ctypes.data_as(C.c_void_p) if interrupt is not None else None
        self._check(self.lib.ort_multi_frame_render(self._m, ray_depth, first_sample, n_samples, iptr, C.byref(done)))
        return int(done.value)

    def frame_wait(self):
        self._check(self.lib.ort_multi_frame_wait(self._m))

    def frame_snapshot(self):
        self._check(self.lib.ort_multi_frame_snapshot(self._m))

    def frame_preview_rgb8(self) -> np.ndarray:
        w, h = self._frame
        out = np.zeros((h, w, 3), np.uint8)
        self._check(self.lib.ort_multi_frame_preview_rgb8(self._m, cabi.ptr(out)))
        return out

    def frame_fetch(self, out=None) -> np.ndarray:
        w, h = self._frame
        if out is None:
            out = np.zeros(w * h, cabi.STATS_DTYPE)
        self._check(self.lib.ort_multi_frame_fetch(self._m, cabi.ptr(out)))
        return out

    def frame_end(self):
        self._check(self.lib.ort_multi_frame_end(self._m))


CKPT_MAGIC = b"ORTCKPT1"


def save_checkpoint(path: str, pixels: np.ndarray, next_sample: int, width: int = 0, height: int = 0):
    """Raw Sample_Stats accumulators + the next sample index (the reference has no
    checkpointing: --continious only writes on exit, main.odin:207,244).  File format shared with the
    C++ command line (host/main.cpp): magic | u32 width | u32 height | u64 next_sample | pixels."""
    pixels = np.ascontiguousarray(pixels, cabi.STATS_DTYPE)
    if width * height != pixels.size:
        width, height = pixels.size, 1
    with open(path, "wb") as f:
        f.write(CKPT_MAGIC)
        f.write(np.array([width, height], np.uint32).tobytes())
        f.write(np.array([next_sample], np.uint64).tobytes())
        f.write(pixels.tobytes())


def load_checkpoint(path: str, width: int, height: int):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != CKPT_MAGIC or len(raw) != 24 + width * height * cabi.STATS_DTYPE.itemsize:
        raise OrtError(f"checkpoint {path} does not match a {width}x{height} Sample_Stats image")
    w, h = np.frombuffer(raw, np.uint32, 2, 8)
    if int(w) * int(h) != width * height:
        raise OrtError(f"checkpoint {path} does not match a {width}x{height} Sample_Stats image")
    next_sample = int(np.frombuffer(raw, np.uint64, 1, 16)[0])
    pixels = np.frombuffer(raw, cabi.STATS_DTYPE, width * height, 24).copy()
    return pixels, next_sample


def mean_image(stats: np.ndarray, width: int, height: int) -> np.ndarray:
    """Linear mean radiance (total / count), image row order (row 0 = top)."""
    cnt = np.maximum(stats["count"].astype(np.float32), 1)[:, None]
    return (stats["total"] / cnt).reshape(height, width, 3)


def rel_rmse(a: np.ndarray, b: np.ndarray):
    """relRMSE = sqrt(mean((a-b)^2)) / mean(b) over RGB, and mean-luminance ratio (Rec.709)."""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-30))
    lw = np.array([0.2126, 0.7152, 0.0722])
    lum = float((a @ lw).mean() / max((b @ lw).mean(), 1e-30))
    return rmse, lum
