"""Single-process multi-device render against a one-device render: size and location of the differences."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import rt_b200
from rt_b200 import gpu
sc = rt_b200.SceneData.load(os.path.join(ROOT, "tests", "golden", "small_lights.rtsc"))
w, h = 96, 64
n = min(gpu.device_count(), 8)
for spp in (32, 16, 2):
    with gpu.RtGpu(1, 0) as one:
        one.upload_scene(sc); one.render(w, h, spp, seed=11); ref, st1 = one.readback()
        one.render(w, h, spp, seed=11, sample_begin=0, sample_end=spp // 2)
        one.render(w, h, spp, seed=11, sample_begin=spp // 2, sample_end=spp, accumulate=True); halves, _ = one.readback()
    with gpu.RtGpu(n, 0) as many:
        many.upload_scene(sc); many.render(w, h, spp, seed=11); img, stn = many.readback()
    for name, a in (("multi", img), ("one device, two halves", halves)):
        d = np.abs(a - ref); rel = d / np.maximum(np.abs(ref), 1e-30)
        bad = d > (1e-7 + 2e-6 * np.abs(ref))
        print(f"spp={spp} {name}: max abs {d.max():.3e} max rel {rel.max():.3e} violating {int(bad.sum())} of {bad.size}; worst ref {ref[np.unravel_index(d.argmax(), d.shape)]:.6f}", flush=True)
    print("  stats", st1["extension_rays"], stn["extension_rays"], st1["light_pdf_rays"], stn["light_pdf_rays"])
