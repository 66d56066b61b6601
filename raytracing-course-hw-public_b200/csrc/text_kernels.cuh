// text_kernels.cuh — device entry for the course text scenes (semantics: text_core.cuh; PARITY UNPINNED).
//
// The scenes hold a handful of analytic primitives, so there is no BVH and no wavefront: one thread owns one
// pixel, keeps the whole scene in shared memory, and loops over its samples (RT_SHADE_PATH) or evaluates the
// deterministic pixel-centre colour once (RT_SHADE_FLAT / RT_SHADE_WHITTED).  One owner per pixel: the float
// sums need no atomics.  Philox is keyed (pixel, sample, bounce) like the glTF path, so sample ranges split
// over devices add up to the single-device image.
#ifndef RT_TEXT_KERNELS_CUH
#define RT_TEXT_KERNELS_CUH

#include <cuda_runtime.h>

#include "text_core.cuh"

namespace rtt {

struct TextParams {
    uint32_t width, height;
    uint32_t s0, s1;   // samples [s0, s1) of every pixel
    uint32_t k0, k1;   // Philox key
    uint32_t ids;      // 1: write the primitive index of the pixel-centre ray instead of radiance
    uint32_t pix0, npix;  // row-major pixel range rendered by this launch (image-tile split)
};

__global__ void __launch_bounds__(128) k_text_render(TextScene s, rt::Camera cam, TextParams tp, float4 *__restrict__ accum,
                                                     int32_t *__restrict__ ids) {
    __shared__ rt_text_prim sh_prims[RT_TEXT_MAX_PRIMS];
    __shared__ rt_text_light sh_lights[RT_TEXT_MAX_LIGHTS];
    __shared__ uint32_t sh_emit[RT_TEXT_MAX_PRIMS];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(s.prims);
        uint32_t *dst = reinterpret_cast<uint32_t *>(sh_prims);
        for (uint32_t i = threadIdx.x; i < s.n_prims * (sizeof(rt_text_prim) / 4); i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint32_t *>(s.lights);
        dst = reinterpret_cast<uint32_t *>(sh_lights);
        for (uint32_t i = threadIdx.x; i < s.n_lights * (sizeof(rt_text_light) / 4); i += blockDim.x) dst[i] = src[i];
        for (uint32_t i = threadIdx.x; i < s.n_emitters; i += blockDim.x) sh_emit[i] = s.emitters[i];
    }
    __syncthreads();
    s.prims = sh_prims;
    s.lights = sh_lights;
    s.emitters = sh_emit;

    const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= tp.npix) return;
    const uint32_t pixel = tp.pix0 + local;
    const uint32_t py = pixel / tp.width, px = pixel - py * tp.width;
    const f3 centre = rt::camera_dir(cam, static_cast<float>(px) + 0.5f, static_cast<float>(py) + 0.5f);
    if (tp.ids) {
        int prim = -1;
        shade_flat(s, cam.pos, centre, &prim);
        ids[pixel] = prim;
        return;
    }
    f3 sum = mk3(0, 0, 0);
    if (s.shading == RT_SHADE_PATH) {
        for (uint32_t smp = tp.s0; smp < tp.s1; ++smp) {
            const rt::RngKey key{pixel, smp, tp.k0, tp.k1};
            const rt::u4 j = rt::rng_jitter(key);
            const f3 dir = rt::camera_dir(cam, static_cast<float>(px) + rt::u01(j.x), static_cast<float>(py) + rt::u01(j.y));
            sum = sum + rt::sanitize(shade_path(s, key, cam.pos, dir));  // sanitize_nans, raytracer.h:607-616
        }
    } else {
        const f3 c = s.shading == RT_SHADE_WHITTED ? shade_whitted(s, cam.pos, centre) : shade_flat(s, cam.pos, centre, nullptr);
        sum = rt::sanitize(c) * static_cast<float>(tp.s1 - tp.s0);  // deterministic: every "sample" is the same value
    }
    float4 a = accum[pixel];
    a.x += sum.x;
    a.y += sum.y;
    a.z += sum.z;
    accum[pixel] = a;
}

}  // namespace rtt

#endif  // RT_TEXT_KERNELS_CUH
