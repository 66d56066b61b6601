// text_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.   *** PARITY UNPINNED ***
//
// CPU statement of the course text-scene renderer.  The reference at HEAD contains no code for these scenes
// (no parser, no plane / ellipsoid / box, no delta lights, no refraction; SURVEY.md section 0, Finding 1), so there
// is nothing to restate and nothing to pin against: this file is the host compilation of the SAME header the CUDA
// kernel includes (csrc/text_core.cuh), executed pixel by pixel.  It checks that the CUDA execution computes what
// the source says; the semantics themselves are checked by the closed-form known-answer tests in
// tests/test_text_scenes.py.  Only tests/ and bench.py's CPU leg may load this library.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "text_core.cuh"

using namespace rtt;

namespace {
struct Host {
    TextScene s;
    std::vector<uint32_t> emitters;
    rt::Camera cam;
};

void setup(const rt_text_scene *sc, uint32_t w, uint32_t h, Host &H) {
    for (uint32_t i = 0; i < sc->n_prims; ++i) {
        const rt_text_prim &p = sc->prims[i];
        if ((p.emission[0] != 0 || p.emission[1] != 0 || p.emission[2] != 0) && p.kind != RT_PRIM_PLANE) H.emitters.push_back(i);
    }
    TextScene &s = H.s;
    s.prims = sc->prims;
    s.lights = sc->lights;
    s.emitters = H.emitters.data();
    s.n_prims = sc->n_prims;
    s.n_lights = sc->n_lights;
    s.n_emitters = (uint32_t)H.emitters.size();
    s.ray_depth = sc->ray_depth;
    s.shading = sc->shading;
    for (int k = 0; k < 3; ++k) {
        s.bg[k] = sc->bg_color[k];
        s.ambient[k] = sc->ambient[k];
    }
    s.eps = sc->eps;
    rt::Camera &c = H.cam;
    c.pos = ld3(sc->camera.position);
    c.right = ld3(sc->camera.right);
    c.up = ld3(sc->camera.up);
    c.fwd = ld3(sc->camera.forward);
    c.tan_half_x = std::tan(sc->camera.fov_x / 2);
    const float fov_y = std::atan(std::tan(sc->camera.fov_x / 2) * (float)h / (float)w) * 2;  // Camera::fov_y, scene.h:69-71
    c.tan_half_y = std::tan(fov_y / 2);
    c.inv_w2 = 2.0f / (float)w;
    c.inv_h2 = 2.0f / (float)h;
}
}  // namespace

extern "C" {

// float means [h][w][3] of samples [s0, s1) divided by `samples` (what rt_gpu_readback returns)
int torc_render(const rt_text_scene *sc, uint32_t w, uint32_t h, uint32_t samples, uint32_t s0, uint32_t s1, uint64_t seed,
                float *rgb_mean, uint32_t n_threads) {
    Host H;
    setup(sc, w, h, H);
    if (n_threads == 0) n_threads = 1;
    auto work = [&](uint32_t t) {
        for (uint32_t pixel = t; pixel < w * h; pixel += n_threads) {
            const uint32_t py = pixel / w, px = pixel % w;
            f3 sum = mk3(0, 0, 0);
            if (H.s.ray_depth > 0 && s1 > s0) {
                if (H.s.shading == RT_SHADE_PATH) {
                    for (uint32_t smp = s0; smp < s1; ++smp) {
                        const rt::RngKey key{pixel, smp, (uint32_t)seed, (uint32_t)(seed >> 32)};
                        const rt::u4 j = rt::rng_jitter(key);
                        const f3 dir = rt::camera_dir(H.cam, (float)px + rt::u01(j.x), (float)py + rt::u01(j.y));
                        sum = sum + rt::sanitize(shade_path(H.s, key, H.cam.pos, dir));
                    }
                } else {
                    const f3 dir = rt::camera_dir(H.cam, (float)px + 0.5f, (float)py + 0.5f);
                    const f3 c = H.s.shading == RT_SHADE_WHITTED ? shade_whitted(H.s, H.cam.pos, dir) : shade_flat(H.s, H.cam.pos, dir, nullptr);
                    sum = rt::sanitize(c) * (float)(s1 - s0);
                }
            }
            rgb_mean[(size_t)pixel * 3 + 0] = sum.x / (float)samples;
            rgb_mean[(size_t)pixel * 3 + 1] = sum.y / (float)samples;
            rgb_mean[(size_t)pixel * 3 + 2] = sum.z / (float)samples;
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    return 0;
}

int torc_ids(const rt_text_scene *sc, uint32_t w, uint32_t h, int32_t *ids) {
    Host H;
    setup(sc, w, h, H);
    for (uint32_t pixel = 0; pixel < w * h; ++pixel) {
        int prim = -1;
        shade_flat(H.s, H.cam.pos, rt::camera_dir(H.cam, (float)(pixel % w) + 0.5f, (float)(pixel / w) + 0.5f), &prim);
        ids[pixel] = prim;
    }
    return 0;
}

// closest hit of one world-space ray: out = t, prim, n.xyz, inside   (known-answer tests)
int torc_closest(const rt_text_scene *sc, const float *o, const float *d, float tmin, float *out) {
    Host H;
    setup(sc, 1, 1, H);
    const THit h = closest(H.s, ld3(o), ld3(d), tmin);
    out[0] = h.t; out[1] = (float)h.prim; out[2] = h.n.x; out[3] = h.n.y; out[4] = h.n.z; out[5] = h.inside ? 1.0f : 0.0f;
    return 0;
}

// Monte-Carlo check of the light sampler against its pdf: returns the mean of 1/pdf over n sampled directions from x,
// which must converge to the solid angle the emitters subtend (known in closed form for a sphere).
double torc_emitter_solid_angle(const rt_text_scene *sc, const float *x, uint32_t n, uint64_t seed) {
    Host H;
    setup(sc, 1, 1, H);
    if (H.s.n_emitters == 0) return 0.0;
    double acc = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        const rt::u4 r = rt::philox4x32_10(i, 0, 0, 0, (uint32_t)seed, (uint32_t)(seed >> 32));
        const rt::u4 r1 = rt::philox4x32_10(i, 1, 0, 0, (uint32_t)seed, (uint32_t)(seed >> 32));
        uint32_t e = (uint32_t)(rt::u01(r.x) * (float)H.s.n_emitters);
        e = e < H.s.n_emitters ? e : H.s.n_emitters - 1;
        const rt_text_prim &L = sc->prims[H.emitters[e]];
        const f3 ql = prim_sample_point(L, rt::u01(r.y), rt::u01(r.z), rt::u01(r1.x));
        const f3 w = normalize(qrot(L.rotation, ql) + ld3(L.position) - ld3(x));
        const float pdf = emitters_pdf(H.s, ld3(x), w);
        if (pdf > 0.0f) acc += 1.0 / pdf;
    }
    return acc / n;
}

}  // extern "C"
