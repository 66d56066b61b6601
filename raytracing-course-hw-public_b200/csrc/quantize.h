// quantize.h — one axis of a quantised wide node (QNode / QNode4 / QNode8 of rt_types.h): the grid word (origin bits |
// cell exponent) and the plane bytes of up to eight children.  Shared by the host packer (repack.h) and the device
// build (gpu_build.cuh), so that both produce the same conservative encoding.
#ifndef RT_QUANTIZE_H
#define RT_QUANTIZE_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RT_QHD __host__ __device__ inline
#else
#define RT_QHD inline
#endif

namespace rt {
namespace detail {

RT_QHD uint32_t fbits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
RT_QHD float bitsf(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
RT_QHD bool finite_f(float f) { return (fbits(f) & 0x7F800000u) != 0x7F800000u; }

// One axis of a node: grid word and the plane bytes (min planes in qlo, max planes in qhi) of its n children.
// Conservative in exact arithmetic with a margin of 1/64 cell for the device's ray-space rounding (pt_core.cuh,
// qnode_axis()).  Returns false when the extent cannot be represented (non-finite box).
RT_QHD bool quantize_axis_n(const float *lo, const float *hi, int n, uint32_t &word, uint8_t *qlo, uint8_t *qhi) {
    float mn = lo[0], mx = hi[0];
    for (int c = 1; c < n; ++c) {
        mn = lo[c] < mn ? lo[c] : mn;
        mx = hi[c] > mx ? hi[c] : mx;
    }
    if (!(finite_f(mn) && finite_f(mx)) || mx < mn) return false;
    const double margin = 1.0 / 64.0;
    const double ext = static_cast<double>(mx) - static_cast<double>(mn);
    int e_first = 1;
    if (ext > 0.0) {
        e_first = ilogb(ext / 255.0) + 127 - 1;
        if (e_first < 1) e_first = 1;
    }
    for (int e = e_first; e <= 238; ++e) {  // smallest cell that covers the extent in 255 steps (e + 16 stays a finite exponent)
        const double cell = ldexp(1.0, e - 127);
        const uint32_t eb = static_cast<uint32_t>(e);
        // origin = the float whose bits are (23 high bits chosen here | bit 8 = 0 | e); it has to be
        // <= mn - margin * cell.  Round that target down to a float, then down to the representable words.
        const double target = static_cast<double>(mn) - margin * cell;
        float tf = static_cast<float>(target);
        if (static_cast<double>(tf) > target) tf = nextafterf(tf, -HUGE_VALF);
        uint32_t w;
        if (tf > 0.0f) {
            const uint32_t tb = fbits(tf);
            w = (tb & ~0x1FFu) | eb;
            if (w > tb) w = (tb & ~0x1FFu) >= 0x200u ? w - 0x200u : (0x80000000u | eb);
        } else {  // negative (or zero): more magnitude = smaller value
            const uint32_t tb = tf == 0.0f ? 0x80000000u : fbits(tf);
            w = (tb & ~0x1FFu) | eb;
            if (w < tb) w += 0x200u;
        }
        const float orgf = bitsf(w);
        const double org = static_cast<double>(orgf);
        if (!finite_f(orgf) || !(org <= target)) continue;
        const double top = (static_cast<double>(mx) - org) / cell + margin;
        if (top > 255.0) continue;
        word = w;
        for (int c = 0; c < n; ++c) {
            double a = floor((static_cast<double>(lo[c]) - org) / cell - margin);
            double z = ceil((static_cast<double>(hi[c]) - org) / cell + margin);
            if (a < 0.0) a = 0.0;  // cannot happen: org <= mn - margin * cell
            if (z > 255.0) z = 255.0;
            qlo[c] = static_cast<uint8_t>(a);
            qhi[c] = static_cast<uint8_t>(z);
        }
        return true;
    }
    return false;
}

}  // namespace detail
}  // namespace rt

#endif  // RT_QUANTIZE_H
