#!/usr/bin/env python3
"""Render every shipped course text scene on the GPU at its own DIMENSIONS / SAMPLES (BASELINE configs 1-2) and the
homebrew scenes at 1920x1080 x 256 spp (config 3); print device time and Msamples/s, with the CPU statement's rate on
a small sample beside it, and write small PNG previews.  Parity unpinned (no reference code for these scenes)."""
import glob
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402

import rt_b200  # noqa: E402,F401
from rt_b200 import gpu, host, textscene as T  # noqa: E402
import oracle_lib as O  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out", "text")
os.makedirs(out_dir, exist_ok=True)
rows = []
with gpu.RtGpu(1, 0) as rt:
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "text", "*.npz"))):
        name = os.path.splitext(os.path.basename(path))[0]
        s = T.load_npz(path)
        configs = [(s.width, s.height, s.samples, "own")]
        if s.shading == T.RT_SHADE_PATH:
            configs.append((1920, 1080, 256, "config3"))
        rt.upload_text_scene(s)
        for w, h, spp, tag in configs:
            rt.render(w, h, spp, seed=1)  # warm-up
            rt.render(w, h, spp, seed=1)
            img, st = rt.readback()
            cw, ch, cspp = min(w, 160), min(h, 120), min(spp, 4)
            t0 = time.perf_counter()
            O.text_render(s, cw, ch, cspp, seed=1)
            cpu = cw * ch * cspp / (time.perf_counter() - t0) / 1e6
            rows.append({"scene": name, "config": tag, "width": w, "height": h, "spp": spp, "shading": int(s.shading),
                         "gpu_ms": st["render_ms"], "gpu_msamples_s": w * h * spp / st["render_ms"] / 1e3,
                         "cpu_statement_msamples_s": cpu, "cpu_threads": min(os.cpu_count() or 1, 16), "mean": float(img.mean())})
            print(rows[-1], flush=True)
            if tag == "own":
                try:
                    from PIL import Image
                    Image.fromarray(host.tonemap_rgb8(img)).resize((max(1, w // 2), max(1, h // 2))).save(os.path.join(out_dir, name + ".png"))
                except Exception as e:  # previews are optional
                    print("no preview:", e)
with open(os.path.join(ROOT, "gpurun_out", "text_bench.json"), "w") as f:
    json.dump(rows, f, indent=1)
