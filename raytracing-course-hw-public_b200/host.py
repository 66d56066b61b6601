"""ctypes binding of librt_host.so (include/rt_host.h): BVH build, tonemap, PPM, RTSC validation."""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import NODE_DTYPE, BvhData, rt_bvh_build, rt_scene_desc

LIB_PATH = os.path.join(_abi.PKG_DIR, "librt_host.so")
SYMBOLS = ["rt_scene_save", "rt_scene_load", "rt_scene_free", "rt_scene_validate", "rt_host_build_bvh",
           "rt_host_free_bvh", "rt_host_tonemap_rgb8", "rt_host_write_ppm", "rt_text_scene_parse", "rt_text_scene_free"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = _abi.load_library(LIB_PATH, "host helpers (librt_host.so)")
        L.rt_host_build_bvh.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                        C.POINTER(rt_bvh_build)]
        L.rt_host_free_bvh.argtypes = [C.POINTER(rt_bvh_build)]
        L.rt_host_free_bvh.restype = None
        L.rt_host_tonemap_rgb8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.rt_host_tonemap_rgb8.restype = None
        L.rt_host_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.rt_scene_validate.argtypes = [C.POINTER(rt_scene_desc)]
        L.rt_scene_save.argtypes = [C.POINTER(rt_scene_desc), C.c_char_p]
        _lib = L
    return _lib


def build_bvh(tri_pos, select=None, min_node_size=4, max_depth=64):
    """BVH::build (src/bvh.h:368-393) over the selected triangles -> BvhData in the reference's node layout."""
    tri_pos = np.ascontiguousarray(tri_pos, np.float32).reshape(-1, 9)
    n = tri_pos.shape[0]
    sel = None
    if select is not None:
        sel = np.ascontiguousarray(select, np.uint8)
        assert sel.shape[0] == n
    out = rt_bvh_build()
    rc = lib().rt_host_build_bvh(tri_pos.ctypes.data if n else None, n, sel.ctypes.data if sel is not None else None,
                                 min_node_size, max_depth, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"rt_host_build_bvh failed: {_abi.STATUS.get(rc, rc)}")
    try:
        nodes = np.zeros(out.n_nodes, NODE_DTYPE)
        objs = np.zeros(out.n_objects, np.uint32)
        if out.n_nodes:
            C.memmove(nodes.ctypes.data, out.nodes, out.n_nodes * NODE_DTYPE.itemsize)
        if out.n_objects:
            C.memmove(objs.ctypes.data, out.objects, out.n_objects * 4)
        return BvhData(nodes, objs, out.root)
    finally:
        lib().rt_host_free_bvh(C.byref(out))


def tonemap_rgb8(rgb_mean):
    """Image::set_pixel (src/image.h:40-82) on float means [..., 3] -> uint8."""
    rgb = np.ascontiguousarray(rgb_mean, np.float32)
    out = np.empty(rgb.shape, np.uint8)
    lib().rt_host_tonemap_rgb8(rgb.ctypes.data, rgb.size // 3, out.ctypes.data)
    return out


def write_ppm(path, rgb8):
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    h, w, _ = rgb8.shape
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)  # main.cpp:41 create_directories
    rc = lib().rt_host_write_ppm(os.fsencode(path), rgb8.ctypes.data, w, h)
    if rc != 0:
        raise RuntimeError(f"cannot write {path}")


def validate(scene):
    d = scene.desc()
    return lib().rt_scene_validate(C.byref(d))
