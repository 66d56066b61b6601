// ref_tool.cpp — TEST INFRASTRUCTURE. Thin harness around the UNMODIFIED reference headers
// (found by include path, -I/root/reference/src; never copied into this repository).
// It exists to pin the C restatement in oracle/pt_oracle.c and to generate the golden
// fixtures under tests/golden/.  Built only where the reference is present, output to
// oracle/_ref/ (git-ignored).  Nothing in the product path links or executes this.
//
//   ref_tool dump    <gltf> <W> <H> <out.rtsc>          flattened scene (loader + both BVH builds)
//   ref_tool ids     <gltf> <W> <H> <out.i32>           primary ids, pixel-centre rays (raytracer.h:516)
//   ref_tool hitinfo <gltf> <W> <H> <out.f32>           to_intersection_info of those hits, 18 floats/pixel
//   ref_tool render  <gltf> <W> <H> <spp> <out.f32> [seed_offset]
//                                                       float means of render_pixel (raytracer.h:618) with the
//                                                       span seeding of run_raytracer (raytracer.h:646-659)
//   ref_tool rng     <seed> <n> <out.bin>               n x {uniform01, uniform(-1,1), uniform(0,2pi), int(0..6), int(0..0)}
//   ref_tool tonemap <in.f32> <n_pixels> <out.u8>       Image::set_pixel (image.h:40) on float triples
//   ref_tool time    <gltf> <W> <H> <spp>               wall seconds of run_raytracer only (load/BVH excluded)
#define STB_IMAGE_IMPLEMENTATION

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "config.h"
#include "geometry.h"
#include "image.h"
#include "raytracer.h"
#include "scene.h"

#include "flatten_ref.hpp"
#include "rt_host.h"

static Scene load(const char *path, unsigned w, unsigned h, unsigned spp) {
    // mirrors src/main.cpp:27-34
    Scene scene = parse_gltf_scene(std::filesystem::path(path), static_cast<float>(w) / h);
    scene.bg_color = {ENV_MAP_INTENSITY, ENV_MAP_INTENSITY, ENV_MAP_INTENSITY};
    // main.cpp:29-31 loads ENV_MAP_PATH when the compile-time USE_ENV_MAP is set; the harness takes the path from the
    // environment instead so that Scene::bg_at (scene.h:83-89) can be exercised without touching config.h
    if (const char *env = std::getenv("RT_ENV_MAP")) scene.bg = geometry::Texture::load_img(env);
    // likewise the compile-time ADD_LIGHT_TRIANGLE (config.h:39-47; scene.h:479-498 appends the object inside the loader)
    if (const char *lt = std::getenv("RT_ADD_LIGHT_TRIANGLE"); lt && std::atoi(lt)) rt_flatten::append_light_triangle(scene);
    scene.camera.width = w;
    scene.camera.height = h;
    scene.samples = spp;
    return scene;
}

template <class T> static void write_file(const char *path, const std::vector<T> &v) {
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
}

int main(int argc, char **argv) try {
    if (argc < 2) {
        std::fprintf(stderr, "usage: ref_tool dump|ids|hitinfo|render|rng|tonemap|time ...\n");
        return 2;
    }
    const std::string cmd = argv[1];

    if (cmd == "rng") {
        int seed = std::atoi(argv[2]);
        int n = std::atoi(argv[3]);
        std::minstd_rand gen(seed);
        std::vector<float> out;
        for (int i = 0; i < n; ++i) {
            out.push_back(std::uniform_real_distribution<float>(0.0f, 1.0f)(gen));
            out.push_back(std::uniform_real_distribution<float>(-1.0f, 1.0f)(gen));
            out.push_back(std::uniform_real_distribution<float>(0.0f, 2 * std::numbers::pi_v<float>)(gen));
            out.push_back(static_cast<float>(std::uniform_int_distribution<>(0, 6)(gen)));
            out.push_back(static_cast<float>(std::uniform_int_distribution<>(0, 0)(gen)));
            out.push_back(static_cast<float>(std::uniform_int_distribution<>(0, 1)(gen)));
        }
        write_file(argv[4], out);
        return 0;
    }
    if (cmd == "tonemap") {
        size_t n = std::strtoull(argv[3], nullptr, 10);
        std::vector<float> in(n * 3);
        std::ifstream(argv[2], std::ios::binary).read(reinterpret_cast<char *>(in.data()), in.size() * 4);
        Image img(static_cast<int>(n), 1);
        for (size_t i = 0; i < n; ++i) img.set_pixel(static_cast<int>(i), {in[i * 3], in[i * 3 + 1], in[i * 3 + 2]});
        std::ofstream out(argv[4], std::ios::binary);
        out.write(reinterpret_cast<const char *>(img.data.data()), n * 3);
        return 0;
    }

    if (argc < 5) return 2;
    const unsigned W = std::strtol(argv[3], nullptr, 10), H = std::strtol(argv[4], nullptr, 10);

    if (cmd == "dump") {
        Scene scene = load(argv[2], W, H, 1);
        RaytracerStaticContext ctx(scene);
        rt_flatten::FlatScene flat;
        rt_flatten::flatten(scene, ctx, flat);
        int rc = rt_scene_save(&flat.desc, argv[5]);
        std::printf("tris %u materials %u textures %u nodes %u lights %u light_nodes %u rc %d\n", flat.desc.n_tris,
                    flat.desc.n_materials, flat.desc.n_textures, flat.desc.scene_bvh.n_nodes,
                    flat.desc.light_bvh.n_objects, flat.desc.light_bvh.n_nodes, rc);
        return rc == 0 ? 0 : 1;
    }
    if (cmd == "ids" || cmd == "hitinfo") {
        Scene scene = load(argv[2], W, H, 1);
        RaytracerStaticContext ctx(scene);
        RaytracerThreadContext tc(ctx, 0);
        std::vector<int32_t> ids(static_cast<size_t>(W) * H, -1);
        std::vector<float> info(static_cast<size_t>(W) * H * 18, 0.0f);
        for (unsigned y = 0; y < H; ++y)
            for (unsigned x = 0; x < W; ++x) {
                auto ray = gen_ray(scene.camera, static_cast<int>(x), static_cast<int>(y));
                auto res = cast_ray(tc, ray);
                size_t p = static_cast<size_t>(y) * W + x;
                if (res.has_value()) {
                    ids[p] = static_cast<int32_t>(res->obj - scene.objects.data());
                    float *o = &info[p * 18];
                    o[0] = res->t;
                    for (int k = 0; k < 3; ++k) o[1 + k] = res->normal.val[k];
                    for (int k = 0; k < 3; ++k) o[4 + k] = res->shading_normal.val[k];
                    for (int k = 0; k < 4; ++k) o[7 + k] = res->color.val[k];
                    for (int k = 0; k < 3; ++k) o[11 + k] = res->emission.val[k];
                    o[14] = res->metallic;
                    o[15] = res->roughness;
                    o[16] = res->is_inside ? 1.0f : 0.0f;
                    o[17] = res->ior;
                }
            }
        if (cmd == "ids")
            write_file(argv[5], ids);
        else
            write_file(argv[5], info);
        return 0;
    }
    if (cmd == "render") {
        const unsigned spp = std::strtol(argv[5], nullptr, 10);
        const int seed_offset = argc > 7 ? std::atoi(argv[7]) : 0;
        Scene scene = load(argv[2], W, H, spp);
        RaytracerStaticContext ctx(scene);
        const size_t n_pix = static_cast<size_t>(W) * H;
        std::vector<float> out(n_pix * 3);
        const int span_count = static_cast<int>((n_pix + SPAN_SIZE - 1) / SPAN_SIZE);
        std::atomic_int next_span(0);
        std::vector<std::thread> workers;
        const unsigned nthreads = std::max(std::thread::hardware_concurrency(), 1u);
        for (unsigned i = 0; i < nthreads; ++i)
            workers.emplace_back([&]() {
                int span;
                while ((span = next_span.fetch_add(1)) < span_count) {
                    RaytracerThreadContext context(ctx, span + seed_offset);  // raytracer.h:648
                    size_t begin = SPAN_SIZE * span, end = std::min(begin + SPAN_SIZE, n_pix);
                    for (size_t p = begin; p < end; ++p) {
                        auto c = render_pixel(context, static_cast<int>(p % W), static_cast<int>(p / W));
                        out[p * 3] = c.r();
                        out[p * 3 + 1] = c.g();
                        out[p * 3 + 2] = c.b();
                    }
                }
            });
        for (auto &t : workers) t.join();
        write_file(argv[6], out);
        return 0;
    }
    if (cmd == "time") {
        const unsigned spp = std::strtol(argv[5], nullptr, 10);
        Scene scene = load(argv[2], W, H, spp);
        Image img(W, H, scene.bg_color);
        std::fflush(stdout);
        FILE *devnull = std::freopen("/dev/null", "w", stdout);  // run_raytracer prints progress (raytracer.h:647)
        auto t0 = std::chrono::steady_clock::now();
        run_raytracer(scene, img);
        auto t1 = std::chrono::steady_clock::now();
        (void)devnull;
        std::fprintf(stderr, "%.6f\n", std::chrono::duration<double>(t1 - t0).count());
        return 0;
    }
    return 2;
} catch (std::exception &e) {
    std::fprintf(stderr, "ref_tool: %s\n", e.what());
    return 1;
}
