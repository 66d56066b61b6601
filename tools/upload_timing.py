import os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scenes')]
os.environ['RT_TIMING']='1'
import bench, rt_b200
from rt_b200 import gltf as gl, gpu
scene = gl.load_gltf(bench.scene_path("big_lights"), 1.0)
rt = gpu.RtGpu(1, 0)
for i in range(4):
    t0=time.perf_counter(); rt.upload_scene(scene); print('upload wall %.1f ms'%((time.perf_counter()-t0)*1e3), flush=True)
