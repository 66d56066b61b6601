// text_core.cuh — the course text scenes (sample_data/scene-*.txt, homebrew_primitives/*.txt): analytic
// plane / ellipsoid / box / triangle primitives with a quaternion pose, flat shading, Whitted-style lights
// with mirror metals and refracting dielectrics, and Monte-Carlo path tracing with emissive primitives.
//
// PARITY UNPINNED.  The reference at HEAD has no parser, no primitive other than the triangle, no delta
// lights and no refraction (Finding 1): there is no reference behaviour to reproduce
// and no golden output.  The semantics below are reconstructed from the vestiges that survive in the
// reference and are cited where they exist:
//   quaternion rotation  v*q                         src/geometry.h:143-147, input order x y z w :154-156
//   ellipsoid = unit-sphere quadratic on o/r, d/r    src/raytracer.h:61-77 (intersect_ray_sphere)
//   box = slab test on [-s, s]                       src/bvh.h:137-152, src/geometry.h:423-433
//   triangle (Cramer, beta/gamma acceptance)         src/bvh.h:36-65
//   camera, jitter                                   src/raytracer.h:516-538, src/scene.h:60-72
//   cosine / light mixture, solid-angle light pdf    src/raytracer.h:79-129,222-262,350-408
//   EPS = 1e-4 as t_min, next ray starts at the hit  src/config.h:15, src/raytracer.h:588-590
//   tonemap / PPM                                    src/image.h:34-82
// and, for what has no vestige (delta lights, ambient term, Fresnel dielectric), from the usual statement of
// those course stages; every such choice is marked "assumed".  The same header is compiled by nvcc for the
// kernels (text_kernels.cuh) and by the host compiler for the CPU statement used in tests
// (oracle/text_oracle.cpp) — that checks the CUDA execution, not the semantics; the semantics are checked
// by closed-form known-answer tests (tests/test_text_scenes.py).
#ifndef RT_TEXT_CORE_CUH
#define RT_TEXT_CORE_CUH

#include "pt_core.cuh"
#include "rt_gpu.h"

namespace rtt {

using rt::f3;
using rt::mk3;
using rt::dot;
using rt::cross;
using rt::normalize;
using rt::len2;

struct TextScene {  // device / host view of an rt_text_scene
    const rt_text_prim *prims;
    const rt_text_light *lights;
    const uint32_t *emitters;  // indices of the finite primitives with non-zero EMISSION
    uint32_t n_prims, n_lights, n_emitters;
    uint32_t ray_depth, shading;
    float bg[3], ambient[3];
    float eps;
};

RT_HD f3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }

// v * q (geometry.h:143-147): rotation of v by the unit quaternion q = (x, y, z, w)
RT_HD f3 qrot(const float *q, f3 v) {
    const f3 qv = mk3(q[0], q[1], q[2]);
    const f3 t = 2.0f * cross(qv, v);
    return v + q[3] * t + cross(qv, t);
}
RT_HD f3 qrot_inv(const float *q, f3 v) {  // v * q.conj()
    const f3 qv = mk3(-q[0], -q[1], -q[2]);
    const f3 t = 2.0f * cross(qv, v);
    return v + q[3] * t + cross(qv, t);
}

// All intersections of the LOCAL-space ray (o, d) with primitive p, unfiltered: t[] ascending, n[] = outward
// local normals.  Returns the number of intersections (0, 1 or 2).
RT_HD int prim_roots(const rt_text_prim &p, f3 o, f3 d, float t[2], f3 n[2]) {
    switch (p.kind) {
    case RT_PRIM_PLANE: {
        const f3 nn = normalize(ld3(p.param));
        const float dn = dot(d, nn);
        if (dn == 0.0f) return 0;
        t[0] = -dot(o, nn) / dn;
        n[0] = nn;
        return 1;
    }
    case RT_PRIM_ELLIPSOID: {  // raytracer.h:61-77
        const f3 r = ld3(p.param);
        const f3 dr = mk3(d.x / r.x, d.y / r.y, d.z / r.z), orr = mk3(o.x / r.x, o.y / r.y, o.z / r.z);
        const float a = dot(dr, dr), hb = dot(orr, dr), c = dot(orr, orr) - 1.0f;
        const float hd2 = hb * hb - a * c;
        if (hd2 < 0.0f) return 0;
        const float hd = sqrtf(hd2);
        t[0] = (-hb - hd) / a;
        t[1] = (-hb + hd) / a;
        for (int k = 0; k < 2; ++k) {
            const f3 q = o + d * t[k];
            n[k] = normalize(mk3(q.x / (r.x * r.x), q.y / (r.y * r.y), q.z / (r.z * r.z)));
        }
        return 2;
    }
    case RT_PRIM_BOX: {  // slab test, bvh.h:137-152, on the local box [-s, s]
        const f3 s = ld3(p.param);
        const float x0 = (-s.x - o.x) / d.x, x1 = (s.x - o.x) / d.x;
        const float y0 = (-s.y - o.y) / d.y, y1 = (s.y - o.y) / d.y;
        const float z0 = (-s.z - o.z) / d.z, z1 = (s.z - o.z) / d.z;
        const float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
        const float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
        if (!(tmin <= tmax)) return 0;
        t[0] = tmin;
        t[1] = tmax;
        for (int k = 0; k < 2; ++k) {  // face = the axis on which the point is relatively farthest out
            const f3 q = o + d * t[k];
            const f3 m = mk3(q.x / s.x, q.y / s.y, q.z / s.z);
            const float ax = fabsf(m.x), ay = fabsf(m.y), az = fabsf(m.z);
            if (ax >= ay && ax >= az) n[k] = mk3(m.x > 0.0f ? 1.0f : -1.0f, 0.0f, 0.0f);
            else if (ay >= az) n[k] = mk3(0.0f, m.y > 0.0f ? 1.0f : -1.0f, 0.0f);
            else n[k] = mk3(0.0f, 0.0f, m.z > 0.0f ? 1.0f : -1.0f);
        }
        return 2;
    }
    default: {  // RT_PRIM_TRIANGLE, bvh.h:36-65
        const f3 a = ld3(p.param), e1 = ld3(p.param + 3) - a, e2 = ld3(p.param + 6) - a;
        float tt, beta, gamma;
        if (!rt::tri_test(a, e1, e2, o, d, -INFINITY, tt, beta, gamma)) return 0;
        t[0] = tt;
        n[0] = normalize(cross(e1, e2));
        return 1;
    }
    }
}

struct THit {
    float t;
    int prim;      // -1 = miss
    f3 n;          // world-space geometric normal, facing the ray
    bool inside;   // the ray left the primitive through this surface (normal was flipped)
};

RT_HD THit closest(const TextScene &s, f3 o, f3 d, float tmin) {
    THit h;
    h.t = INFINITY;
    h.prim = -1;
    h.n = mk3(0, 0, 1);
    h.inside = false;
    f3 nl = mk3(0, 0, 1);
    for (uint32_t i = 0; i < s.n_prims; ++i) {
        const rt_text_prim &p = s.prims[i];
        const f3 ol = qrot_inv(p.rotation, o - ld3(p.position)), dl = qrot_inv(p.rotation, d);
        float t[2];
        f3 n[2];
        const int cnt = prim_roots(p, ol, dl, t, n);
        for (int k = 0; k < cnt; ++k)
            if (t[k] >= tmin && t[k] < h.t) {  // strictly closer wins, first found wins ties (bvh.h:132)
                h.t = t[k];
                h.prim = static_cast<int>(i);
                nl = n[k];
            }
    }
    if (h.prim >= 0) {
        h.n = qrot(s.prims[h.prim].rotation, nl);
        h.inside = dot(h.n, d) > 0.0f;
        if (h.inside) h.n = -h.n;
    }
    return h;
}

RT_HD bool occluded(const TextScene &s, f3 o, f3 d, float tmin, float tmax) {
    for (uint32_t i = 0; i < s.n_prims; ++i) {
        const rt_text_prim &p = s.prims[i];
        const f3 ol = qrot_inv(p.rotation, o - ld3(p.position)), dl = qrot_inv(p.rotation, d);
        float t[2];
        f3 n[2];
        const int cnt = prim_roots(p, ol, dl, t, n);
        for (int k = 0; k < cnt; ++k)
            if (t[k] >= tmin && t[k] < tmax) return true;
    }
    return false;
}

RT_HD f3 reflect(f3 d, f3 n) { return d - n * (2.0f * dot(d, n)); }  // geometry::reflect

// Fresnel dielectric (assumed): Schlick reflectance with R0 from the IOR, Snell refraction, total internal
// reflection -> mirror.  `n` faces the ray; `inside` = the ray travels inside the primitive.
struct Fresnel {
    f3 refl, refr;
    float r;     // reflectance in [0, 1]; 1 on total internal reflection
};
RT_HD Fresnel fresnel(f3 d, f3 n, float ior, bool inside) {
    Fresnel f;
    f.refl = reflect(d, n);
    f.refr = f.refl;
    const float eta = inside ? ior : 1.0f / ior;  // n1 / n2
    const float cos_i = fminf(1.0f, -dot(d, n));
    const float sin2_t = eta * eta * (1.0f - cos_i * cos_i);
    if (sin2_t > 1.0f) {
        f.r = 1.0f;
        return f;
    }
    const float cos_t = sqrtf(1.0f - sin2_t);
    const float r0s = (1.0f - ior) / (1.0f + ior), r0 = r0s * r0s;
    f.r = r0 + (1.0f - r0) * rt::pow5(1.0f - cos_i);
    f.refr = normalize(d * eta + n * (eta * cos_i - cos_t));
    return f;
}

// ---- RT_SHADE_FLAT: colour of the nearest primitive, else the background (assumed: the first course stage)
RT_HD f3 shade_flat(const TextScene &s, f3 o, f3 d, int *prim_out) {
    const THit h = closest(s, o, d, 0.0f);
    if (prim_out) *prim_out = h.prim;
    return h.prim < 0 ? ld3(s.bg) : ld3(s.prims[h.prim].color);
}

// ---- RT_SHADE_WHITTED (assumed): diffuse = colour * (ambient + sum of shadowed delta lights), metallic =
// colour * mirror reflection, dielectric = Fresnel blend of reflection and refraction (refraction tinted by
// the colour when entering); recursion to RAY_DEPTH, written as an explicit stack of weighted rays.
RT_HD f3 shade_whitted(const TextScene &s, f3 o0, f3 d0) {
    struct Item {
        f3 o, d, w;
        uint32_t depth;
    };
    Item stack[RT_TEXT_MAX_DEPTH + 2];  // depth-first: at most one pending sibling per level
    int sp = 0;
    f3 result = mk3(0, 0, 0);
    stack[sp++] = Item{o0, d0, mk3(1, 1, 1), s.ray_depth};
    while (sp > 0) {
        const Item it = stack[--sp];
        if (it.depth == 0) continue;  // trace_ray(depth 0) = 0, raytracer.h:596-598
        const THit h = closest(s, it.o, it.d, s.eps);
        if (h.prim < 0) {
            result = result + it.w * ld3(s.bg);
            continue;
        }
        const rt_text_prim &p = s.prims[h.prim];
        const f3 pos = it.o + it.d * h.t;
        const f3 col = ld3(p.color);
        result = result + it.w * ld3(p.emission);
        if (p.material == RT_MAT_METALLIC) {
            stack[sp++] = Item{pos, reflect(it.d, h.n), it.w * col, it.depth - 1};
        } else if (p.material == RT_MAT_DIELECTRIC) {
            const Fresnel f = fresnel(it.d, h.n, p.ior, h.inside);
            if (f.r < 1.0f) {
                const f3 tint = h.inside ? mk3(1, 1, 1) : col;
                stack[sp++] = Item{pos, f.refr, it.w * tint * (1.0f - f.r), it.depth - 1};
            }
            stack[sp++] = Item{pos, f.refl, it.w * f.r, it.depth - 1};
        } else {
            f3 light = ld3(s.ambient);
            for (uint32_t l = 0; l < s.n_lights; ++l) {
                const rt_text_light &L = s.lights[l];
                f3 dir;
                float dist = INFINITY;
                f3 inten = ld3(L.intensity);
                if (L.kind == RT_LIGHT_DIRECTIONAL) {
                    dir = normalize(ld3(L.vec));  // LIGHT_DIRECTION points towards the light
                } else {
                    const f3 to = ld3(L.vec) - pos;
                    dist = sqrtf(len2(to));
                    dir = to * (1.0f / dist);
                    inten = inten * (1.0f / (L.attenuation[0] + L.attenuation[1] * dist + L.attenuation[2] * dist * dist));
                }
                const float c = dot(h.n, dir);
                if (c <= 0.0f) continue;
                if (occluded(s, pos, dir, s.eps, dist)) continue;
                light = light + inten * c;
            }
            result = result + it.w * col * light;
        }
    }
    return result;
}

// ---- RT_SHADE_PATH: Monte-Carlo path tracing (the estimator of raytracer.h:555-605 specialised to these
// materials): emission at every hit, fixed depth without Russian roulette, diffuse bounces sampled from the
// 50/50 mixture of cosine and light-surface sampling (cosine only without emitters).
struct AreaSample {
    f3 p, n;     // local space
    float pdf;   // per unit area
};

// Uniform-direction sample on an ellipsoid (point = r * u, u uniform on the unit sphere) and its area pdf
// 1 / (4 pi |(u.x r.y r.z, r.x u.y r.z, r.x r.y u.z)|); box: a face chosen by area, uniform on it.
RT_HD float ellipsoid_area_pdf(f3 r, f3 q) {  // q on the surface
    const f3 u = mk3(q.x / r.x, q.y / r.y, q.z / r.z);
    const f3 j = mk3(u.x * r.y * r.z, r.x * u.y * r.z, r.x * r.y * u.z);
    return 1.0f / (4.0f * RT_PI * sqrtf(len2(j)));
}
RT_HD float prim_area_pdf(const rt_text_prim &p, f3 q) {
    if (p.kind == RT_PRIM_ELLIPSOID) return ellipsoid_area_pdf(ld3(p.param), q);
    if (p.kind == RT_PRIM_BOX) {
        const f3 s = ld3(p.param);
        return 1.0f / (8.0f * (s.y * s.z + s.x * s.z + s.x * s.y));
    }
    const f3 a = ld3(p.param), e1 = ld3(p.param + 3) - a, e2 = ld3(p.param + 6) - a;
    return 2.0f / sqrtf(len2(cross(e1, e2)));  // triangle: 1 / area
}
RT_HD f3 prim_sample_point(const rt_text_prim &p, float u0, float u1, float u2) {
    if (p.kind == RT_PRIM_ELLIPSOID) {
        const float z = u0 * 2.0f - 1.0f, c = sqrtf(fmaxf(0.0f, 1.0f - z * z));
        float sn, cs;
        rt::sincos_2pi(u1, sn, cs);
        return mk3(p.param[0] * c * cs, p.param[1] * c * sn, p.param[2] * z);
    }
    if (p.kind == RT_PRIM_BOX) {
        const f3 s = ld3(p.param);
        const float wx = s.y * s.z, wy = s.x * s.z, wz = s.x * s.y;
        float pick = u2 * (wx + wy + wz);
        const float a = u0 * 2.0f - 1.0f, b = u1 * 2.0f - 1.0f;
        float side = 1.0f;
        if (pick < wx) {
            if (pick * 2.0f < wx) side = -1.0f;
            return mk3(side * s.x, a * s.y, b * s.z);
        }
        pick -= wx;
        if (pick < wy) {
            if (pick * 2.0f < wy) side = -1.0f;
            return mk3(a * s.x, side * s.y, b * s.z);
        }
        pick -= wy;
        if (pick * 2.0f < wz) side = -1.0f;
        return mk3(a * s.x, b * s.y, side * s.z);
    }
    float u = u0, v = u1;  // triangle_dist::sample, raytracer.h:227-239
    if (u + v > 1.0f) {
        u = 1.0f - u;
        v = 1.0f - v;
    }
    const f3 a = ld3(p.param), e1 = ld3(p.param + 3) - a, e2 = ld3(p.param + 6) - a;
    return a + e1 * v + e2 * u;
}

// solid-angle pdf of direction w from x under "uniform emitter, uniform point on it": every intersection of the
// ray with every emitter counts (raytracer.h:363-375, 79-84)
RT_HD float emitters_pdf(const TextScene &s, f3 x, f3 w) {
    float sum = 0.0f;
    for (uint32_t e = 0; e < s.n_emitters; ++e) {
        const rt_text_prim &p = s.prims[s.emitters[e]];
        const f3 ol = qrot_inv(p.rotation, x - ld3(p.position)), dl = qrot_inv(p.rotation, w);
        float t[2];
        f3 n[2];
        const int cnt = prim_roots(p, ol, dl, t, n);
        for (int k = 0; k < cnt; ++k)
            if (t[k] >= s.eps) {
                const float c = fabsf(dot(dl, n[k]));
                if (c > 0.0f) sum += prim_area_pdf(p, ol + dl * t[k]) * t[k] * t[k] / c;
            }
    }
    return sum / static_cast<float>(s.n_emitters);
}

RT_HD f3 shade_path(const TextScene &s, const rt::RngKey &key, f3 o, f3 d) {
    f3 thr = mk3(1, 1, 1), rad = mk3(0, 0, 0);
    for (uint32_t b = 0; b < s.ray_depth; ++b) {
        const THit h = closest(s, o, d, s.eps);
        if (h.prim < 0) {
            rad = rad + thr * ld3(s.bg);
            break;
        }
        const rt_text_prim &p = s.prims[h.prim];
        rad = rad + thr * ld3(p.emission);
        if (b + 1 == s.ray_depth) break;
        const f3 pos = o + d * h.t;
        const rt::u4 r0 = rt::rng_block(key, b, 0);
        f3 w;
        if (p.material == RT_MAT_METALLIC) {
            w = reflect(d, h.n);
            thr = thr * ld3(p.color);
        } else if (p.material == RT_MAT_DIELECTRIC) {
            const Fresnel f = fresnel(d, h.n, p.ior, h.inside);
            if (rt::u01(r0.x) < f.r) {
                w = f.refl;
            } else {
                w = f.refr;
                if (!h.inside) thr = thr * ld3(p.color);
            }
        } else {
            const bool pick_light = s.n_emitters > 0 && !(rt::u01(r0.x) * 2.0f < 1.0f);
            if (pick_light) {
                uint32_t e = static_cast<uint32_t>(rt::u01(r0.y) * static_cast<float>(s.n_emitters));
                e = e < s.n_emitters ? e : s.n_emitters - 1;
                const rt_text_prim &L = s.prims[s.emitters[e]];
                const rt::u4 r1 = rt::rng_block(key, b, 1);
                const f3 ql = prim_sample_point(L, rt::u01(r0.z), rt::u01(r0.w), rt::u01(r1.x));
                w = normalize(qrot(L.rotation, ql) + ld3(L.position) - pos);
            } else {
                w = rt::cosine_sample(h.n, rt::u01(r0.z), rt::u01(r0.w));
            }
            if (rt::any_nan(w)) break;  // raytracer.h:569-571
            const float c = dot(w, h.n);
            float pdf = rt::cosine_pdf(h.n, w);
            if (s.n_emitters > 0) pdf = 0.5f * (pdf + emitters_pdf(s, pos, w));
            if (!(c > 0.0f) || pdf < s.eps) break;  // raytracer.h:576-586
            thr = thr * ld3(p.color) * (RT_INV_PI * c / pdf);
        }
        o = pos;
        d = w;
    }
    return rad;
}

}  // namespace rtt

#endif  // RT_TEXT_CORE_CUH
