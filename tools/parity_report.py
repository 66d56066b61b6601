"""Path-by-path and statistical parity figures of the loaded librt_gpu build (RT_GPU_LIB selects a variant), one line
per golden scene: fraction of pixels off by > 1e-3 against the oracle's Philox mode at 32 spp, median relative difference,
relative difference of the image means; relMSE / 8-bit MAE against the reference's high-spp render with the
reference-vs-reference control."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "scenes")]
import numpy as np
import oracle_lib as O
import rt_b200
from rt_b200 import gpu, host
from conftest import rel_mse

G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "manifest.json")))["scenes"]
rt = gpu.RtGpu(1, 0)
print("lib:", gpu.LIB_PATH)
for name in ["tiny", "tiny_env", "tiny_lt", "small_lights", "small_lights_lt", "texmaps", "texall"]:
    m = man[name]; w, h, hi = m["width"], m["height"], m["hi_spp"]
    sc = rt_b200.SceneData.load(os.path.join(G, name + ".rtsc"))
    rt.upload_scene(sc)
    rt.render(w, h, 32, seed=2024)
    img, _ = rt.readback()
    ref, _ = O.render(sc, w, h, 32, rng_mode=O.RNG_PHILOX, seed=2024)
    rel = (np.abs(img - ref) / (np.abs(ref) + 1e-3)).max(axis=2)
    a = np.fromfile(os.path.join(G, name + "_refhi_a.f32"), np.float32).reshape(h, w, 3)
    b = np.fromfile(os.path.join(G, name + "_refhi_b.f32"), np.float32).reshape(h, w, 3)
    rt.render(w, h, 8 * hi, seed=77)
    big, _ = rt.readback()
    mae = np.abs(host.tonemap_rgb8(big).astype(int) - host.tonemap_rgb8(a).astype(int)).mean()
    print(f"{name:16s} off>1e-3 {100 * (rel > 1e-3).mean():6.2f} %  median {np.median(rel):.1e}  mean diff {abs(img.mean() - ref.mean()) / ref.mean():.1e} | "
          f"relMSE {rel_mse(big, a).max():.2e} (ref-vs-ref {rel_mse(a, b).max():.2e})  MAE {mae:.3f}  mean vs ref {abs(big.mean() - a.mean()) / a.mean():.1e} (ref-vs-ref {abs(b.mean() - a.mean()) / a.mean():.1e})", flush=True)
