import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scenes")]
import bench
import rt_b200
from rt_b200 import gltf as gl, gpu
scene = gl.load_gltf(bench.scene_path("big_lights"), 1.0)
rt = gpu.RtGpu(1, 0)
rt.upload_scene(scene)
for prof in (True,):
    rt.set_profiling(prof)
    for i in range(16):
        t0 = time.perf_counter()
        rt.render(1000, 1000, 128, seed=1)
        wall = (time.perf_counter() - t0) * 1e3
        st = rt.stats()
        print(f"profiling={prof} call {i}: wall {wall:7.2f} ms, events {st['render_ms']:7.2f} ms, gen {st['kernel_ms'][0]:6.2f} ext {st['kernel_ms'][1]:7.2f} shade {st['kernel_ms'][2]:6.2f} acc {st['kernel_ms'][3]:5.2f}", flush=True)
