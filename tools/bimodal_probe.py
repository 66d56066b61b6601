"""k_shade's two speeds (k_shade): one process = one line "shade ms / extend ms per 128 spp".
usage: python tools/bimodal_probe.py [--dummy-mb N] [--iters K] [--paths P]
--dummy-mb allocates (and keeps) a device buffer before the backend allocates its queue arena, to shift the arena's
placement in the physical address space."""
import argparse, os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scenes")]
import bench
import rt_b200
from rt_b200 import gltf as gl, gpu

ap = argparse.ArgumentParser()
ap.add_argument("--dummy-mb", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--paths", type=int, default=0)
ap.add_argument("--tag", default="")
ap.add_argument("--spp", type=int, default=128)
a = ap.parse_args()
scene = gl.load_gltf(bench.scene_path("big_lights"), 1.0)
keep = None
if a.dummy_mb:
    cudart = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = cudart.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(a.dummy_mb << 20))
    assert rc == 0, rc
    keep = p
rt = gpu.RtGpu(1, 0)
rt.upload_scene(scene)
rt.set_profiling(True)
sh, ex = [], []
for i in range(a.iters):
    rt.render(1000, 1000, a.spp, seed=1, max_paths_in_flight=a.paths)
    st = rt.stats()
    if i:
        sh.append(st["kernel_ms"][2]); ex.append(st["kernel_ms"][1])
print(f"probe {a.tag} spp={a.spp} dummy_mb={a.dummy_mb} shade {min(sh):.2f}..{max(sh):.2f} extend {min(ex):.2f}..{max(ex):.2f}", flush=True)
