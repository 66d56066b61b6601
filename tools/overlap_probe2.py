"""Do k_extend (L1 / issue bound) and k_shade (latency / DRAM bound) gain from sharing the SMs?  N handles on one GPU,
each rendering S/N spp of config 4 with its persistent grids limited to a share of an SM (RT_EXT_CTAS_PER_SM /
RT_SHADE_CTAS_PER_SM), started `stagger` ms apart so that one handle's k_shade meets another's k_extend.
usage: python tools/overlap_probe2.py [S]"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scenes")]
import bench
import rt_b200
from rt_b200 import gltf as gl, gpu

S = int(sys.argv[1]) if len(sys.argv) > 1 else 240
scene = gl.load_gltf(bench.scene_path("big_lights"), 1.0)

def handles(n, ext, shade):
    os.environ["RT_EXT_CTAS_PER_SM"] = str(ext)
    os.environ["RT_SHADE_CTAS_PER_SM"] = str(shade)
    hs = [gpu.RtGpu(1, 0) for _ in range(n)]
    for h in hs:
        h.upload_scene(scene)
    return hs

def run(hs, stagger_ms):
    n = len(hs)
    per = S // n
    def one(i):
        if stagger_ms: time.sleep(i * stagger_ms * 1e-3)
        hs[i].render(1000, 1000, S, seed=1, sample_begin=i * per, sample_end=(i + 1) * per)
    best = 1e9
    for rep in range(4):
        th = [threading.Thread(target=one, args=(i,)) for i in range(n)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        best = min(best, time.perf_counter() - t0)
    return best * 1e3

for n, ext, shade, stagger in [(1, 8, 7, 0), (2, 8, 7, 0), (2, 4, 3, 0), (2, 4, 3, 4), (2, 5, 2, 4), (2, 4, 4, 4), (2, 6, 2, 4), (3, 3, 2, 3),
                               (4, 2, 2, 2), (4, 2, 1, 2), (1, 6, 7, 0), (1, 8, 4, 0)]:
    hs = handles(n, ext, shade)
    ms = run(hs, stagger)
    print(f"S={S} handles={n} ext_ctas={ext} shade_ctas={shade} stagger={stagger} ms: {ms:.2f} ms  ({S * 1e6 / ms * 1e-3:.1f} Msamples/s wall)", flush=True)
    for h in hs: h.close() if hasattr(h, "close") else None
    del hs
