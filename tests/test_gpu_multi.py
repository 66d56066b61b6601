"""Multi-GPU parity (pytest -m gpu on a box with >= 2 B200s; skipped on one GPU).

The per-pixel Monte-Carlo loop shards by samples: device g of N renders samples
[g*spp/N, (g+1)*spp/N) of every pixel, then ONE reduce(sum) of the float accumulation buffers.  Because Philox is
keyed by the global sample index, the N-device image must equal the 1-device image up to float summation order."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

from rt_b200 import gpu

pytestmark = pytest.mark.gpu


def _n_dev():
    try:
        return gpu.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_dev() < 2, reason="needs >= 2 GPUs")


@needs2
def test_single_process_multi_device_matches_one_device(golden_scene):
    """rt_gpu_create(n_gpus = N): per-device streams + ncclReduce to device 0 inside rt_gpu_render."""
    sc = golden_scene("small_lights")
    w, h, spp = 96, 64, 32
    with gpu.RtGpu(1, 0) as one:
        one.upload_scene(sc)
        one.render(w, h, spp, seed=11)
        ref, st1 = one.readback()
    n = min(_n_dev(), 8)
    with gpu.RtGpu(n, 0) as many:
        many.upload_scene(sc)
        many.render(w, h, spp, seed=11)
        img, stn = many.readback()
        rgb8 = many.readback_rgb8()
    assert np.allclose(img, ref, rtol=2e-6, atol=1e-7)
    assert stn["samples"] == st1["samples"] and stn["extension_rays"] == st1["extension_rays"]
    assert rgb8.shape == (h, w, 3)
    # fewer samples than devices: the image is split into pixel ranges instead (tile split), still complete
    with gpu.RtGpu(n, 0) as many:
        many.upload_scene(sc)
        many.render(w, h, 1, seed=11)
        img1, _ = many.readback()
    with gpu.RtGpu(1, 0) as one:
        one.upload_scene(sc)
        one.render(w, h, 1, seed=11)
        ref1, _ = one.readback()
    assert np.allclose(img1, ref1, rtol=2e-6, atol=1e-7)


@needs2
def test_accumulate_on_a_multi_device_handle(golden_scene):
    """RT_FLAG_ACCUMULATE with n_gpus > 1: the reduce leaves the total on device 0 while the other
    devices keep their partial sums, which a second accumulating render must not add again — two half renders
    equal one full render."""
    sc = golden_scene("small_lights")
    w, h, spp = 96, 64, 32
    n = min(_n_dev(), 8)
    with gpu.RtGpu(n, 0) as many:
        many.upload_scene(sc)
        many.render(w, h, spp, seed=11)
        full, _ = many.readback()
        many.render(w, h, spp, seed=11, sample_begin=0, sample_end=spp // 2)
        many.render(w, h, spp, seed=11, sample_begin=spp // 2, sample_end=spp, accumulate=True)
        halves, _ = many.readback()
        # and a third accumulating pass of zero samples changes nothing
        many.render(w, h, spp, seed=11, sample_begin=spp, sample_end=spp, accumulate=True)
        again, _ = many.readback()
    assert np.allclose(halves, full, rtol=2e-6, atol=1e-7)
    assert np.array_equal(again, halves)


_RANK_SCRIPT = r"""
import os, sys
sys.path[:0] = [{root!r}, os.path.join({root!r}, "tests")]
import numpy as np, torch, torch.distributed as dist
import rt_b200
from rt_b200 import gpu, dist as rtdist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sc = rt_b200.SceneData.load(os.path.join({root!r}, "tests", "golden", "small_lights.rtsc"))
rt = gpu.RtGpu(1, lr)
rt.upload_scene(sc)
rtdist.render_distributed(rt, 96, 64, 32, 11, rank, world, lr)
if rank == 0:
    img, _ = rt.readback()
    np.save({out!r}, img)
dist.barrier()
rt.close()
dist.destroy_process_group()
"""


@needs2
def test_one_process_per_gpu_nccl_reduce_matches_one_device(golden_scene, tmp_path):
    """The bench's layout: torchrun, one rank per GPU, torch.distributed (NCCL) reduce of the device buffers."""
    out = str(tmp_path / "img.npy")
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], check=True, env=env, timeout=600)
    img = np.load(out)
    with gpu.RtGpu(1, 0) as one:
        one.upload_scene(golden_scene("small_lights"))
        one.render(96, 64, 32, seed=11)
        ref, _ = one.readback()
    assert np.allclose(img, ref, rtol=2e-6, atol=1e-7)
