#!/usr/bin/env python
"""Builds and runs tools/wide_study.cpp on the 260k-triangle bench scene (host only, an experiment, not a test)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "build", "libwide_study.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-x", "c++", "-I" + os.path.join(ROOT, "include"),
                       "-I" + os.path.join(ROOT, "raytracing-course-hw-public_b200", "csrc"), "-o", so, os.path.join(ROOT, "tools", "wide_study.cpp")])
from rt_b200 import gltf
import bench
scene = gltf.load_gltf(bench.scene_path("big_lights") if hasattr(bench, "scene_path") else sys.argv[1], 1.0)
d = scene.desc()
L = C.CDLL(so)
L.wide_study(C.byref(d), int(sys.argv[2]) if len(sys.argv) > 2 else 48, int(sys.argv[3]) if len(sys.argv) > 3 else 48, 2)
