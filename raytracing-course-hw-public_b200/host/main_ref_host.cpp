// main_ref_host.cpp — the reference-hosted drop-in: `raytracer_b200 <scene.gltf> <width> <height> <samples> <out.ppm>`.
//
// Same five positional arguments, same exit codes and the same output file as the reference CLI
// (src/main.cpp:16-49).  Everything on the host is the reference's own code, used through its headers
// (-I<reference>/src, never copied): parse_gltf_scene (scene.h:183), the light-BVH build of
// RaytracerStaticContext (raytracer.h:444-447), Image::set_pixel's tonemap / gamma / quantisation
// (image.h:40-82) and Image::write (image.h:34).  The only replaced call is run_raytracer(scene, img)
// (main.cpp:37): the per-pixel Monte-Carlo loop runs on the B200(s) through the rt_gpu C ABI.
//
// Optional environment (additions; the 5-argument form needs none of them):
//   RT_GPUS   number of GPUs of this box to split the samples over (default 1)
//   RT_SEED   Philox seed (default 0)
//   RT_HOST_SCENE_BVH  1: also build the reference's scene BVH and pass it (default: the library builds its own)
//   RT_ENV_MAP  image file used as the equirectangular environment map Scene::bg (the run-time form of the
//             reference's compile-time USE_ENV_MAP / ENV_MAP_PATH, src/config.h:35-37)
#define STB_IMAGE_IMPLEMENTATION

#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#include "config.h"
#include "geometry.h"
#include "image.h"
#include "raytracer.h"
#include "scene.h"

#include "flatten_ref.hpp"
#include "rt_gpu.h"

namespace {
void check(int rc, const char *what) {
    if (rc != RT_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + rt_gpu_last_error());
}
unsigned env_uint(const char *name, unsigned fallback) {
    const char *v = std::getenv(name);
    return v && *v ? static_cast<unsigned>(std::strtoul(v, nullptr, 10)) : fallback;
}
}  // namespace

int main(int argc, char **argv) try {
    if (argc < 6) {
        std::cerr << "Too few arguments: expected 6, got " << argc - 1 << std::endl;
        return EXIT_FAILURE;
    }
    const unsigned width = std::strtol(argv[2], nullptr, 10);
    const unsigned height = std::strtol(argv[3], nullptr, 10);
    const unsigned samples = std::strtol(argv[4], nullptr, 10);

    Scene scene = parse_gltf_scene(std::filesystem::path(argv[1]), static_cast<float>(width) / height);
    scene.bg_color = {ENV_MAP_INTENSITY, ENV_MAP_INTENSITY, ENV_MAP_INTENSITY};
    if constexpr (USE_ENV_MAP) scene.bg = geometry::Texture::load_img(ENV_MAP_PATH);  // main.cpp:29-31
    if (const char *env = std::getenv("RT_ENV_MAP")) scene.bg = geometry::Texture::load_img(env);
    scene.camera.width = width;
    scene.camera.height = height;
    scene.samples = samples;
    Image img(width, height, scene.bg_color);

    if (scene.ray_depth != 0) {  // run_raytracer's early-out, raytracer.h:630
        // Host-side BVH builds stay the reference's code, but only the light BVH is needed: its object order is what
        // bvh_mix_dist::sample indexes (raytracer.h:355-361).  The scene BVH the GPU traverses is the library's own
        // (include/rt_gpu.h: scene_bvh.n_nodes == 0), so the reference's 0.85 s scene build (260k triangles) is skipped
        // unless RT_HOST_SCENE_BVH=1 asks for the reference's tree to be built and passed (RT_KEEP_HOST_BVH=1 makes the
        // library traverse it as is).
        const bool host_tree = env_uint("RT_HOST_SCENE_BVH", 0) != 0;
        std::unique_ptr<RaytracerStaticContext> ctx;
        BVH light_only;
        rt_flatten::FlatScene flat;
        if (host_tree) {
            ctx.reset(new RaytracerStaticContext(scene));
            rt_flatten::flatten(scene, *ctx, flat);
        } else {
            light_only = rt_flatten::build_light_bvh(scene);
            rt_flatten::flatten(scene, nullptr, light_only, flat);
        }

        rt_gpu_ctx *gpu = nullptr;
        check(rt_gpu_create(&gpu, static_cast<int>(env_uint("RT_GPUS", 1)), 0), "rt_gpu_create");
        check(rt_gpu_upload_scene(gpu, &flat.desc), "rt_gpu_upload_scene");
        rt_render_params params{};
        params.width = width;
        params.height = height;
        params.samples = samples;
        params.mode = RT_MODE_BEAUTY;
        params.seed = env_uint("RT_SEED", 0);
        check(rt_gpu_render(gpu, &params), "rt_gpu_render");
        std::vector<float> mean(static_cast<size_t>(width) * height * 3);
        rt_stats stats{};
        check(rt_gpu_readback(gpu, mean.data(), nullptr, &stats), "rt_gpu_readback");
        rt_gpu_destroy(gpu);
        for (size_t i = 0; i < static_cast<size_t>(width) * height; ++i)
            img.set_pixel(static_cast<int>(i), {mean[i * 3], mean[i * 3 + 1], mean[i * 3 + 2]});
        std::fprintf(stderr, "rt_gpu: %.1f ms, %.1f Msamples/s, %.1f Mrays/s (extension) + %.1f Mrays/s (light pdf)\n",
                     stats.render_ms, stats.samples / stats.render_ms * 1e-3, stats.extension_rays / stats.render_ms * 1e-3,
                     stats.light_pdf_rays / stats.render_ms * 1e-3);
    }

    std::filesystem::path out_path = argv[5];
    std::filesystem::create_directories(out_path.parent_path());
    std::ofstream out(out_path, std::ios::binary);
    img.write(out);
    return EXIT_SUCCESS;
} catch (std::runtime_error &err) {
    std::cerr << err.what() << std::endl;
    return EXIT_FAILURE;
}
