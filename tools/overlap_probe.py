"""Would two half-batches on two streams hide the kernel tails?  Two handles on the same GPU render S/2 spp each,
one after the other and concurrently (two host threads: ctypes releases the GIL), against one handle rendering S spp."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scenes")]
import bench
import rt_b200
from rt_b200 import gltf as gl, gpu

S = int(sys.argv[1]) if len(sys.argv) > 1 else 124
scene = gl.load_gltf(bench.scene_path("big_lights"), 1.0)
a, b = gpu.RtGpu(1, 0), gpu.RtGpu(1, 0)
for h in (a, b):
    h.upload_scene(scene)
def one(h, s0, s1):
    h.render(1000, 1000, S, seed=1, sample_begin=s0, sample_end=s1)
for rep in range(3):
    t0 = time.perf_counter(); one(a, 0, S); t_full = time.perf_counter() - t0
    t0 = time.perf_counter(); one(a, 0, S // 2); one(b, S // 2, S); t_seq = time.perf_counter() - t0
    t0 = time.perf_counter()
    th = [threading.Thread(target=one, args=(a, 0, S // 2)), threading.Thread(target=one, args=(b, S // 2, S))]
    [t.start() for t in th]; [t.join() for t in th]
    t_conc = time.perf_counter() - t0
    print(f"S={S}: one handle {t_full*1e3:.2f} ms | two halves sequential {t_seq*1e3:.2f} ms | two halves concurrent {t_conc*1e3:.2f} ms", flush=True)
