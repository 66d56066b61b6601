"""ctypes binding of librt_gpu.so (include/rt_gpu.h) — the CUDA backend. No fallback: a missing library or a
missing GPU raises."""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (RT_FLAG_ACCUMULATE, RT_MODE_BEAUTY, RT_MODE_PRIMARY_IDS, rt_render_params, rt_scene_desc, rt_stats)

# RT_GPU_LIB selects another build of the same C ABI (kernel A/B experiments, tools/build_variants.sh)
LIB_PATH = os.environ.get("RT_GPU_LIB") or os.path.join(_abi.PKG_DIR, "librt_gpu.so")

# every symbol include/rt_gpu.h declares
SYMBOLS = ["rt_gpu_create", "rt_gpu_destroy", "rt_gpu_upload_scene", "rt_gpu_upload_text_scene", "rt_gpu_render", "rt_gpu_readback",
           "rt_gpu_accum_device_ptr", "rt_gpu_readback_rgb8", "rt_gpu_set_profiling", "rt_gpu_debug_get_bvh", "rt_gpu_fp32_peak", "rt_gpu_last_error",
           "rt_gpu_device_count", "rt_gpu_abi_version"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = _abi.load_library(LIB_PATH, "CUDA backend (librt_gpu.so)")
        L.rt_gpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
        L.rt_gpu_destroy.argtypes = [C.c_void_p]
        L.rt_gpu_destroy.restype = None
        L.rt_gpu_upload_scene.argtypes = [C.c_void_p, C.POINTER(rt_scene_desc)]
        L.rt_gpu_upload_text_scene.argtypes = [C.c_void_p, C.c_void_p]
        L.rt_gpu_render.argtypes = [C.c_void_p, C.POINTER(rt_render_params)]
        L.rt_gpu_readback.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(rt_stats)]
        L.rt_gpu_accum_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.rt_gpu_readback_rgb8.argtypes = [C.c_void_p, C.c_void_p]
        L.rt_gpu_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.rt_gpu_debug_get_bvh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]
        L.rt_gpu_fp32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.rt_gpu_last_error.restype = C.c_char_p
        _lib = L
    return _lib


class RtGpuError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        msg = lib().rt_gpu_last_error().decode(errors="replace")
        raise RtGpuError(f"{what}: {_abi.STATUS.get(rc, rc)}: {msg}")


def device_count():
    return lib().rt_gpu_device_count()


class RtGpu:
    """One handle = n_gpus devices of this process (rt_gpu_create)."""

    def __init__(self, n_gpus=1, first_device=0):
        self._h = C.c_void_p()
        _check(lib().rt_gpu_create(C.byref(self._h), n_gpus, first_device), "rt_gpu_create")
        self.n_gpus = n_gpus
        self._scene = None
        self._last = None

    def close(self):
        if self._h:
            lib().rt_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload_scene(self, scene):
        d = scene.desc()
        _check(lib().rt_gpu_upload_scene(self._h, C.byref(d)), "rt_gpu_upload_scene")
        self._scene = scene

    def upload_text_scene(self, scene):
        """Course text scene (textscene.TextScene); render / readback / primary_ids then work on it."""
        d = scene.desc()
        _check(lib().rt_gpu_upload_text_scene(self._h, C.byref(d)), "rt_gpu_upload_text_scene")
        self._scene = scene

    def set_profiling(self, enable):
        _check(lib().rt_gpu_set_profiling(self._h, int(bool(enable))), "rt_gpu_set_profiling")

    def render(self, width, height, samples, seed=0, sample_begin=0, sample_end=0, mode=RT_MODE_BEAUTY,
               max_paths_in_flight=0, accumulate=False, pixel_begin=0, pixel_end=0):
        p = rt_render_params(width, height, samples, sample_begin, sample_end, mode, seed, max_paths_in_flight,
                             RT_FLAG_ACCUMULATE if accumulate else 0, pixel_begin, pixel_end)
        _check(lib().rt_gpu_render(self._h, C.byref(p)), "rt_gpu_render")
        self._last = (width, height, mode)

    def readback(self):
        """(rgb_mean float32 [H, W, 3], stats dict) of the last beauty render."""
        w, h, _ = self._last
        out = np.empty((h, w, 3), np.float32)
        st = rt_stats()
        _check(lib().rt_gpu_readback(self._h, out.ctypes.data_as(C.c_void_p), None, C.byref(st)), "rt_gpu_readback")
        return out, st.as_dict()

    def readback_into(self, out):
        """Readback into a caller-provided (e.g. pinned) float32 array of W*H*3 elements."""
        st = rt_stats()
        _check(lib().rt_gpu_readback(self._h, C.c_void_p(out.ctypes.data), None, C.byref(st)), "rt_gpu_readback")
        return st.as_dict()

    def stats(self):
        st = rt_stats()
        _check(lib().rt_gpu_readback(self._h, None, None, C.byref(st)), "rt_gpu_readback")
        return st.as_dict()

    def readback_rgb8(self):
        w, h, _ = self._last
        out = np.empty((h, w, 3), np.uint8)
        _check(lib().rt_gpu_readback_rgb8(self._h, out.ctypes.data_as(C.c_void_p)), "rt_gpu_readback_rgb8")
        return out

    def primary_ids(self, width, height):
        self.render(width, height, 1, mode=RT_MODE_PRIMARY_IDS)
        ids = np.empty((height, width), np.int32)
        _check(lib().rt_gpu_readback(self._h, None, ids.ctypes.data_as(C.c_void_p), None), "rt_gpu_readback")
        return ids

    def debug_bvh(self, which, dtype):
        """Diagnostic download of the device-built scene BVH (rt_gpu_debug_get_bvh); numpy array of `dtype`."""
        n = C.c_size_t(0)
        _check(lib().rt_gpu_debug_get_bvh(self._h, which, None, C.byref(n)), "rt_gpu_debug_get_bvh")
        out = np.zeros(n.value // np.dtype(dtype).itemsize, dtype)
        if n.value:
            _check(lib().rt_gpu_debug_get_bvh(self._h, which, out.ctypes.data_as(C.c_void_p), C.byref(n)), "rt_gpu_debug_get_bvh")
        return out

    def fp32_peak_tflops(self):
        v = C.c_double()
        _check(lib().rt_gpu_fp32_peak(self._h, C.byref(v)), "rt_gpu_fp32_peak")
        return v.value

    def accum_device_ptr(self):
        """(device pointer, n_floats) of device 0's float4 per-pixel sums (for an external collective)."""
        p = C.c_void_p()
        n = C.c_size_t()
        _check(lib().rt_gpu_accum_device_ptr(self._h, C.byref(p), C.byref(n)), "rt_gpu_accum_device_ptr")
        return p.value, n.value
