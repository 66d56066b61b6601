#!/usr/bin/env python3
"""Parse the course text scenes shipped with the reference (sample_data/*.txt, read where they lie under
/root/reference, which does not exist on the GPU box) with rt_text_scene_parse and store the PARSED scenes as small
.npz fixtures under tests/golden/text/.  These are inputs only: the reference has no code that renders them, so there
are no golden outputs (parity unpinned)."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rt_b200  # noqa: E402,F401
from rt_b200 import textscene  # noqa: E402

SRC = "/root/reference/sample_data"
OUT = os.path.join(ROOT, "tests", "golden", "text")
os.makedirs(OUT, exist_ok=True)
for path in sorted(glob.glob(os.path.join(SRC, "*.txt")) + glob.glob(os.path.join(SRC, "homebrew_primitives", "*.txt"))):
    name = os.path.splitext(os.path.basename(path))[0]
    sc = textscene.load_text_scene(path)
    textscene.save_npz(sc, os.path.join(OUT, name + ".npz"))
    print(name, sc.width, sc.height, "depth", sc.ray_depth, "spp", sc.samples, "shading", sc.shading, len(sc.prims), "prims",
          len(sc.lights), "lights")
