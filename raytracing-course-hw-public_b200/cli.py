"""`python -m rt_b200.cli <scene.gltf|scene.txt> <width> <height> <samples> <out.ppm>` — the reference's CLI surface
(src/main.cpp:16-49) for hosts without the reference headers: Python loader (gltf.py, bit-identical flattening),
own BVH build (librt_host), CUDA integrator (librt_gpu), reference tonemap restated in librt_host, binary PPM."""
import os
import sys


def run_raytracer(scene, width, height, samples, seed=0, n_gpus=1):
    """Mirror of run_raytracer(scene, image) (src/raytracer.h:629): float means [H, W, 3] + stats."""
    from . import gpu, textscene

    with gpu.RtGpu(n_gpus, 0) as rt:
        if isinstance(scene, textscene.TextScene):
            rt.upload_text_scene(scene)
        else:
            rt.upload_scene(scene)
        # one-shot command: 128 Mi paths per batch (17 GB of queues instead of 68 GB), see host/main_ref_host.cpp
        rt.render(width, height, samples, seed=seed, max_paths_in_flight=int(os.environ.get("RT_PATHS", 128 << 20)))
        return rt.readback()


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) < 6:
        print(f"Too few arguments: expected 6, got {len(argv) - 1}", file=sys.stderr)
        return 1
    from . import gltf, host, textscene

    try:
        width, height, samples = int(argv[2]), int(argv[3]), int(argv[4])
        if textscene.is_text_scene(argv[1]):
            # course text scenes (parity unpinned): width / height / samples of 0 take DIMENSIONS / SAMPLES of the file
            scene = textscene.load_text_scene(argv[1])
            width, height, samples = width or scene.width, height or scene.height, samples or scene.samples
        else:
            scene = gltf.load_gltf(argv[1], width / height, env_map=os.environ.get("RT_ENV_MAP") or None)
        mean, stats = run_raytracer(scene, width, height, samples, seed=int(os.environ.get("RT_SEED", "0")),
                                    n_gpus=int(os.environ.get("RT_GPUS", "1")))
        host.write_ppm(argv[5], host.tonemap_rgb8(mean))
        ms = max(stats["render_ms"], 1e-9)
        print(f"rt_gpu: {ms:.1f} ms, {stats['samples'] / ms * 1e-3:.1f} Msamples/s", file=sys.stderr)
    except RuntimeError as e:
        print(e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
