"""The N>1 path on CPU: world_size-2 gloo processes run the sample-split + reduce plumbing of rt_b200.dist with
the oracle standing in for the renderer (tests may use the oracle; the product never does)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT

from rt_b200 import dist as rtdist


def test_sample_range_is_a_partition():
    for samples in (0, 1, 7, 1000, 4096):
        for world in (1, 2, 3, 4, 8):
            r = [rtdist.sample_range(samples, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == samples
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rtdist.sample_range(10, 2, 2)


def test_work_split_samples_then_image_tiles():
    """Samples are split while every rank gets at least one; below that the image is split into pixel ranges
    (north star: "falling back to image tiles for small spp").  Either way the work is a partition."""
    n_pix = 24 * 18
    for world in (1, 2, 3, 8):
        for samples in (0, 1, 2, 7, 8, 100):
            parts = [rtdist.work_split(samples, n_pix, k, world) for k in range(world)]
            covered = np.zeros((max(samples, 1), n_pix), np.int32)
            for sb, se, pb, pe in parts:
                assert 0 <= sb <= se <= samples and 0 <= pb <= pe <= n_pix
                covered[sb:se, pb:pe] += 1
            assert (covered[:samples] == 1).all()  # every (sample, pixel) exactly once
            if samples >= world:
                assert all(p[2:] == (0, n_pix) for p in parts)  # sample split: all pixels on every rank
            else:
                assert all(p[:2] == (0, samples) for p in parts)  # tile split: all samples on every rank


def _worker(rank, world, port, out_path):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import oracle_lib as O
    import rt_b200

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = rt_b200.SceneData.load(os.path.join(GOLDEN, "tiny.rtsc"))
    w, h, seed = 24, 18, 5
    merged = {}
    for spp in (10, 1):  # 10 samples: sample split; 1 sample on 2 ranks: image-tile split
        sb, se, pb, pe = rtdist.work_split(spp, w * h, rank, world)
        mean, _ = O.render(scene, w, h, spp, rng_mode=O.RNG_PHILOX, seed=seed, sample_begin=sb, sample_end=se, n_threads=1)
        sums = (mean * np.float32(spp)).reshape(w * h, 3)
        mask = np.zeros((w * h, 1), np.float32)
        mask[pb:pe] = 1  # the backend renders only [pb, pe); the oracle renders all pixels, so mask the rest
        sums = torch.from_numpy((sums * mask).reshape(h, w, 3))
        rtdist.reduce_sums(sums, 0)
        merged[spp] = sums.numpy() / np.float32(spp)
    if rank == 0:
        np.savez(out_path, **{f"spp{k}": v for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_split_equals_single_render(tmp_path):
    import oracle_lib as O
    import rt_b200

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    merged = np.load(out)
    scene = rt_b200.SceneData.load(os.path.join(GOLDEN, "tiny.rtsc"))
    for spp in (10, 1):
        full, _ = O.render(scene, 24, 18, spp, rng_mode=O.RNG_PHILOX, seed=5)
        assert np.allclose(merged[f"spp{spp}"], full, rtol=1e-5, atol=1e-7)


def test_merge_host_sums():
    a = np.full((2, 2, 4), 3.0, np.float32)
    b = np.full((2, 2, 4), 5.0, np.float32)
    assert np.array_equal(rtdist.merge_host_sums([a, b], 4), np.full((2, 2, 4), 2.0, np.float32))
