"""Multi-GPU plumbing, one process per GPU (torchrun): sample-split + one reduce.

The per-pixel Monte-Carlo loop shards by samples: rank r of R renders samples
[spp*r/R, spp*(r+1)/R) of EVERY pixel into its own float-sum buffer (Philox is keyed by the global sample
index, so the union over ranks is exactly the single-GPU sample set), then ONE reduce(sum) of the W*H*4
float buffer to rank 0 — the only exchange step of the path — and rank 0 divides by spp.
`torch.distributed` is the plumbing (NCCL on GPUs; gloo in the CPU tests), never the compute.
"""
import numpy as np


def sample_range(samples, rank, world):
    """Samples [begin, end) of every pixel that `rank` of `world` renders; a partition of [0, samples)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return samples * rank // world, samples * (rank + 1) // world


def work_split(samples, n_pixels, rank, world):
    """(sample_begin, sample_end, pixel_begin, pixel_end) of `rank`: samples are split while there is at least one
    per rank (every rank renders all pixels); with fewer samples than ranks the IMAGE is split into contiguous
    pixel ranges instead and every rank renders all samples of its range (north star: "falling back to image tiles
    for small spp").  Either way the ranks' buffers are disjoint contributions that one reduce(sum) merges."""
    if samples >= world:
        sb, se = sample_range(samples, rank, world)
        return sb, se, 0, n_pixels
    pb, pe = sample_range(n_pixels, rank, world)
    return 0, samples, pb, pe


class _CudaView:
    """__cuda_array_interface__ wrapper around the backend's accumulation buffer (zero copy)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def accum_as_tensor(rt, device_index):
    """torch view (no copy) of the backend's per-pixel float4 sums on `cuda:device_index`."""
    import torch

    ptr, n = rt.accum_device_ptr()
    return torch.as_tensor(_CudaView(ptr, n), device=f"cuda:{device_index}")


def reduce_sums(tensor, dst=0):
    """The path's single collective: sum of the accumulation buffers onto rank `dst`."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
        if tensor.is_cuda:
            # the backend reuses this buffer from its own stream on the next render / readback
            import torch

            torch.cuda.synchronize(tensor.device)
    return tensor


def render_distributed(rt, width, height, samples, seed, rank, world, device_index, max_paths_in_flight=0):
    """Render this rank's sample range on its GPU, reduce to rank 0. Returns the torch view of the sums
    (complete on rank 0 only)."""
    sb, se, pb, pe = work_split(samples, width * height, rank, world)
    rt.render(width, height, samples, seed=seed, sample_begin=sb, sample_end=max(se, sb),
              max_paths_in_flight=max_paths_in_flight, pixel_begin=pb, pixel_end=pe)
    t = accum_as_tensor(rt, device_index)
    return reduce_sums(t, 0)


def merge_host_sums(partial_sums, samples):
    """Reference semantics of the merge on host arrays (used by the CPU tests): mean = sum of sums / spp."""
    total = np.zeros_like(partial_sums[0], dtype=np.float32)
    for p in partial_sums:
        total += p
    return total / np.float32(samples)
