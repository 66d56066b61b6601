"""ctypes binding of oracle/libpt_oracle.so — TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and --impl reference
legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(ORACLE_DIR, "libpt_oracle.so")
REF_TOOL = os.path.join(ORACLE_DIR, "_ref", "ref_tool")
REF_BINARY = os.path.join(ORACLE_DIR, "_ref", "raytracer")


class orc_params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32), ("sample_begin", C.c_uint32),
                ("sample_end", C.c_uint32), ("rng_mode", C.c_uint32), ("seed", C.c_uint64),
                ("seed_offset", C.c_int32), ("n_threads", C.c_uint32)]


class orc_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "extension_rays", "light_pdf_rays", "shades", "nodes_visited",
                                          "box_tests", "tri_tests", "light_nodes_visited", "light_box_tests",
                                          "light_tri_tests")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle", "CC=gcc"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_render.restype = C.c_int
        _lib.orc_primary_ids.restype = C.c_int
        _lib.orc_u01.restype = C.c_float
        _lib.orc_u01.argtypes = [C.c_uint32]
    return _lib


RNG_MINSTD = 0  # the reference's generator and draw order: bit-comparable with the reference
RNG_PHILOX = 1  # keyed (pixel, sample, bounce): path-comparable with the CUDA backend


def render(scene, width, height, samples, rng_mode=RNG_PHILOX, seed=0, sample_begin=0, sample_end=0, seed_offset=0,
           n_threads=0):
    """float32 [H, W, 3] means + stats dict."""
    p = orc_params(width, height, samples, sample_begin, sample_end, rng_mode, seed, seed_offset,
                   n_threads or (os.cpu_count() or 1))
    out = np.zeros((height, width, 3), np.float32)
    st = orc_stats()
    d = scene.desc()
    rc = lib().orc_render(C.byref(d), C.byref(p), out.ctypes.data_as(C.c_void_p), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"orc_render failed: {rc}")
    return out, st.as_dict()


def primary_ids(scene, width, height, want_hitinfo=False):
    ids = np.zeros((height, width), np.int32)
    info = np.zeros((height, width, 18), np.float32) if want_hitinfo else None
    d = scene.desc()
    rc = lib().orc_primary_ids(C.byref(d), C.c_uint32(width), C.c_uint32(height), ids.ctypes.data_as(C.c_void_p),
                               info.ctypes.data_as(C.c_void_p) if want_hitinfo else None)
    if rc != 0:
        raise RuntimeError(f"orc_primary_ids failed: {rc}")
    return (ids, info) if want_hitinfo else ids


def tonemap_rgb8(rgb):
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.zeros(rgb.shape, np.uint8)
    lib().orc_tonemap_rgb8(rgb.ctypes.data_as(C.c_void_p), C.c_size_t(rgb.size // 3), out.ctypes.data_as(C.c_void_p))
    return out


def rng_stream(seed, n):
    out = np.zeros((n, 6), np.float32)
    lib().orc_rng_stream(C.c_int32(seed), C.c_uint32(n), out.ctypes.data_as(C.c_void_p))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox(c, k, o)
    return list(o)


def have_ref_tool():
    return os.path.exists(REF_TOOL)


def ref_tool(*args):
    """Run oracle/_ref/ref_tool (the harness around the unmodified reference headers)."""
    return subprocess.run([REF_TOOL, *[str(a) for a in args]], check=True, capture_output=True, text=True)


# ---- course text scenes: host compilation of csrc/text_core.cuh (PARITY UNPINNED, see oracle/text_oracle.cpp) ----
TEXT_LIB_PATH = os.path.join(ORACLE_DIR, "libtext_oracle.so")
_text_lib = None


def text_lib():
    global _text_lib
    if _text_lib is None:
        if not os.path.exists(TEXT_LIB_PATH):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle", "CC=gcc", "CXX=g++"], check=True, capture_output=True)
        L = C.CDLL(TEXT_LIB_PATH)
        L.torc_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                  C.c_void_p, C.c_uint32]
        L.torc_ids.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.torc_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.torc_emitter_solid_angle.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
        L.torc_emitter_solid_angle.restype = C.c_double
        _text_lib = L
    return _text_lib


def text_render(scene, width, height, samples, seed=0, sample_begin=0, sample_end=0, n_threads=0):
    d = scene.desc()
    out = np.zeros((height, width, 3), np.float32)
    n_threads = n_threads or min(os.cpu_count() or 1, 16)
    text_lib().torc_render(C.addressof(d), width, height, samples, sample_begin, sample_end or samples, seed,
                           out.ctypes.data, n_threads)
    return out


def text_ids(scene, width, height):
    d = scene.desc()
    out = np.zeros((height, width), np.int32)
    text_lib().torc_ids(C.addressof(d), width, height, out.ctypes.data)
    return out


def text_closest(scene, origin, direction, tmin=0.0):
    """(t, prim, normal[3], inside) of the closest hit of one world-space ray."""
    d = scene.desc()
    o = np.asarray(origin, np.float32)
    dr = np.asarray(direction, np.float32)
    out = np.zeros(6, np.float32)
    text_lib().torc_closest(C.addressof(d), o.ctypes.data, dr.ctypes.data, tmin, out.ctypes.data)
    return float(out[0]), int(out[1]), out[2:5].copy(), bool(out[5])


def text_emitter_solid_angle(scene, x, n, seed=1):
    d = scene.desc()
    xx = np.asarray(x, np.float32)
    return text_lib().torc_emitter_solid_angle(C.addressof(d), xx.ctypes.data, n, seed)
