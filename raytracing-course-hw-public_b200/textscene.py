"""Course text scenes (sample_data/scene-*.txt, homebrew_primitives/*.txt) — host side.

PARITY UNPINNED: the reference at HEAD cannot load these files; grammar and semantics are
documented in include/rt_gpu.h (`rt_text_scene`) and csrc/text_core.cuh.  The parser is C++ (librt_host.so,
`rt_text_scene_parse`); this module mirrors the POD structs for ctypes and owns the parsed scene as numpy arrays.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import rt_camera

RT_PRIM_PLANE, RT_PRIM_ELLIPSOID, RT_PRIM_BOX, RT_PRIM_TRIANGLE = range(4)
RT_MAT_DIFFUSE, RT_MAT_METALLIC, RT_MAT_DIELECTRIC = range(3)
RT_LIGHT_DIRECTIONAL, RT_LIGHT_POINT = range(2)
RT_SHADE_FLAT, RT_SHADE_WHITTED, RT_SHADE_PATH = range(3)

PRIM_DTYPE = np.dtype([("kind", "<u4"), ("material", "<u4"), ("param", "<f4", 9), ("position", "<f4", 3),
                       ("rotation", "<f4", 4), ("color", "<f4", 3), ("emission", "<f4", 3), ("ior", "<f4")])
LIGHT_DTYPE = np.dtype([("kind", "<u4"), ("intensity", "<f4", 3), ("vec", "<f4", 3), ("attenuation", "<f4", 3)])
assert PRIM_DTYPE.itemsize == 100 and LIGHT_DTYPE.itemsize == 40


class rt_text_scene(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("ray_depth", C.c_uint32),
                ("samples", C.c_uint32), ("shading", C.c_uint32), ("n_prims", C.c_uint32), ("n_lights", C.c_uint32),
                ("bg_color", C.c_float * 3), ("ambient", C.c_float * 3), ("camera", rt_camera), ("eps", C.c_float),
                ("prims", C.c_void_p), ("lights", C.c_void_p)]


class TextScene:
    """A parsed text scene; `desc()` gives the ctypes rt_text_scene aliasing the numpy arrays."""

    def __init__(self):
        self.width, self.height, self.ray_depth, self.samples, self.shading = 640, 480, 1, 1, RT_SHADE_FLAT
        self.bg_color = np.zeros(3, np.float32)
        self.ambient = np.zeros(3, np.float32)
        self.cam_position = np.zeros(3, np.float32)
        self.cam_right = np.array([1, 0, 0], np.float32)
        self.cam_up = np.array([0, 1, 0], np.float32)
        self.cam_forward = np.array([0, 0, -1], np.float32)
        self.fov_x = np.float32(1.5708)
        self.eps = np.float32(1e-4)
        self.prims = np.zeros(0, PRIM_DTYPE)
        self.lights = np.zeros(0, LIGHT_DTYPE)

    @staticmethod
    def new_prim(kind, param, position=(0, 0, 0), rotation=(0, 0, 0, 1), color=(0, 0, 0), emission=(0, 0, 0),
                 material=RT_MAT_DIFFUSE, ior=1.5):
        p = np.zeros(1, PRIM_DTYPE)
        p["kind"], p["material"], p["ior"] = kind, material, ior
        p["param"][0, :len(param)] = param
        p["position"], p["rotation"], p["color"], p["emission"] = position, rotation, color, emission
        return p

    def desc(self):
        d = rt_text_scene()
        d.abi_version = _abi.RT_GPU_ABI_VERSION
        d.width, d.height, d.ray_depth, d.samples, d.shading = self.width, self.height, self.ray_depth, self.samples, self.shading
        d.n_prims, d.n_lights = len(self.prims), len(self.lights)
        for k in range(3):
            d.bg_color[k] = self.bg_color[k]
            d.ambient[k] = self.ambient[k]
            d.camera.position[k] = self.cam_position[k]
            d.camera.right[k] = self.cam_right[k]
            d.camera.up[k] = self.cam_up[k]
            d.camera.forward[k] = self.cam_forward[k]
        d.camera.fov_x = self.fov_x
        d.eps = self.eps
        self.prims = np.ascontiguousarray(self.prims)
        self.lights = np.ascontiguousarray(self.lights)
        d.prims = self.prims.ctypes.data if len(self.prims) else None
        d.lights = self.lights.ctypes.data if len(self.lights) else None
        return d


_SCALARS = ("width", "height", "ray_depth", "samples", "shading")
_VECTORS = ("bg_color", "ambient", "cam_position", "cam_right", "cam_up", "cam_forward")


def save_npz(scene, path):
    """Parsed scene as a small .npz (test fixtures derived from the course files)."""
    np.savez(path, scalars=np.array([getattr(scene, k) for k in _SCALARS], np.uint32),
             vectors=np.stack([getattr(scene, k) for k in _VECTORS]).astype(np.float32),
             fov_eps=np.array([scene.fov_x, scene.eps], np.float32), prims=scene.prims, lights=scene.lights)


def load_npz(path):
    z = np.load(path)
    s = TextScene()
    for k, v in zip(_SCALARS, z["scalars"]):
        setattr(s, k, int(v))
    for k, v in zip(_VECTORS, z["vectors"]):
        setattr(s, k, v.astype(np.float32))
    s.fov_x, s.eps = np.float32(z["fov_eps"][0]), np.float32(z["fov_eps"][1])
    s.prims = z["prims"].astype(PRIM_DTYPE)
    s.lights = z["lights"].astype(LIGHT_DTYPE)
    return s


def is_text_scene(path):
    return os.path.splitext(path)[1].lower() == ".txt"


def load_text_scene(path):
    """rt_text_scene_parse (librt_host.so) -> TextScene. Raises on unknown commands / malformed lines."""
    from . import host

    L = host.lib()
    L.rt_text_scene_parse.argtypes = [C.c_char_p, C.POINTER(C.POINTER(rt_text_scene))]
    L.rt_text_scene_free.argtypes = [C.POINTER(rt_text_scene)]
    L.rt_text_scene_free.restype = None
    out = C.POINTER(rt_text_scene)()
    rc = L.rt_text_scene_parse(os.fsencode(path), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"cannot parse text scene {path}: {_abi.STATUS.get(rc, rc)}")
    try:
        d = out.contents
        s = TextScene()
        s.width, s.height, s.ray_depth, s.samples, s.shading = d.width, d.height, d.ray_depth, d.samples, d.shading
        s.bg_color = np.array(d.bg_color[:], np.float32)
        s.ambient = np.array(d.ambient[:], np.float32)
        s.cam_position = np.array(d.camera.position[:], np.float32)
        s.cam_right = np.array(d.camera.right[:], np.float32)
        s.cam_up = np.array(d.camera.up[:], np.float32)
        s.cam_forward = np.array(d.camera.forward[:], np.float32)
        s.fov_x = np.float32(d.camera.fov_x)
        s.eps = np.float32(d.eps)
        s.prims = np.zeros(d.n_prims, PRIM_DTYPE)
        s.lights = np.zeros(d.n_lights, LIGHT_DTYPE)
        if d.n_prims:
            C.memmove(s.prims.ctypes.data, d.prims, d.n_prims * PRIM_DTYPE.itemsize)
        if d.n_lights:
            C.memmove(s.lights.ctypes.data, d.lights, d.n_lights * LIGHT_DTYPE.itemsize)
        return s
    finally:
        L.rt_text_scene_free(out)
