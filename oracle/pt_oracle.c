/*
 * pt_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A CPU restatement, in plain C, of the reference's per-pixel Monte-Carlo integrator
 * (firelion9/raytracing-course-hw-public): every function cites the reference file:line it
 * follows and keeps the reference's float expression order, so that in RNG mode 0
 * (std::minstd_rand drawn in the reference's order, seeded per 256-pixel span) it reproduces
 * the reference's float means BIT FOR BIT.  It is pinned against the unmodified reference by
 * oracle/_ref/ref_tool (built from /root/reference) and the fixtures under tests/golden/.
 * RNG mode 1 keys Philox4x32-10 by (pixel, sample, bounce) exactly like the CUDA path, so the
 * GPU result can be compared path by path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (librt_gpu.so) never does and has no CPU path.
 *
 * Input is the same flattened rt_scene_desc the product's C ABI takes (include/rt_gpu.h), i.e.
 * reference-layout BVH nodes (src/bvh.h:157) and scene.objects-order triangles.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rt_gpu.h"

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* public (test-only) interface                                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct orc_params {
    uint32_t width, height;
    uint32_t samples;       /* divisor and, in mode 0, the sample count */
    uint32_t sample_begin;  /* mode 1 only: renders [sample_begin, sample_end) */
    uint32_t sample_end;    /* 0 => samples */
    uint32_t rng_mode;      /* 0 = minstd_rand in reference order, 1 = Philox keyed (pixel,sample,bounce) */
    uint64_t seed;          /* Philox key (mode 1) */
    int32_t seed_offset;    /* mode 0: span seed = span index + seed_offset (0 = the reference) */
    uint32_t n_threads;     /* 0 => 1 */
} orc_params;

typedef struct orc_stats {
    uint64_t samples, extension_rays, light_pdf_rays, shades;
    uint64_t nodes_visited, box_tests, tri_tests;      /* scene BVH, per reference traversal */
    uint64_t light_nodes_visited, light_box_tests, light_tri_tests;
} orc_stats;

/* ------------------------------------------------------------------------------------------ */
/* vector helpers: the generated vec3 operators (src/generated/vectors.generated.inline.h)     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }   /* vec * scl */
static inline v3 smul(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }   /* scl * vec */
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float len2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline float len(v3 a) { return sqrtf(len2(a)); }
static inline v3 norm(v3 a) { return divs(a, len(a)); }                         /* geometry.h:31-34 */
static inline v3 crs(v3 a, v3 b) {                                               /* geometry.h:18-24 */
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float det(v3 a, v3 b, v3 c) { return dot(a, crs(b, c)); }          /* geometry.h:26-29 */
static inline v3 transform3(v3 l, v3 x, v3 y, v3 z) {                            /* geometry.h:355-359 */
    return add(add(smul(l.x, x), smul(l.y, y)), smul(l.z, z));
}
static inline v3 reflect(v3 normal, v3 in_dir) {                                 /* geometry.h:36-40 */
    return sub(in_dir, muls(smul(2, normal), dot(in_dir, normal)));
}
static inline float std_min(float a, float b) { return (b < a) ? b : a; }        /* std::min */
static inline float std_max(float a, float b) { return (a < b) ? b : a; }        /* std::max */
static inline float pow2f(float x) { return x * x; }                             /* raytracer.h:24 */
static inline float pow5f(float x) {                                             /* pow<5>, raytracer.h:28-38 */
    float x2 = x * x;
    float x4 = x2 * x2;
    return x * (x4 * 1.0f);
}
static const float PI_F = 3.14159265358979323846f;

/* ------------------------------------------------------------------------------------------ */
/* RNG                                                                                          */
/* ------------------------------------------------------------------------------------------ */
/* std::minstd_rand = linear_congruential_engine<uint_fast32_t, 48271, 0, 2147483647> */
static inline uint32_t minstd_seed(int32_t s) {
    /* seed(s): the value is converted to uint_fast32_t (64-bit) first, then reduced mod m */
    uint64_t v = (uint64_t)(int64_t)s;
    uint64_t r = v % 2147483647u;
    return r == 0 ? 1u : (uint32_t)r;
}
static inline uint32_t minstd_next(uint32_t *state) {
    *state = (uint32_t)(((uint64_t)*state * 48271u) % 2147483647u);
    return *state;
}
/* std::generate_canonical<float, 24>(minstd_rand) in libstdc++ 13: one draw, divided by the float
 * nearest to 2147483646 (= 2^31), clamped below 1. */
static inline float minstd_canonical(uint32_t *state) {
    float sum = (float)(minstd_next(state) - 1u);
    float ret = sum / 2147483648.0f;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret;
}
/* std::uniform_int_distribution<int>(0, n-1)(minstd_rand): libstdc++ "downscaling" branch */
static inline uint32_t minstd_int(uint32_t *state, uint32_t n) {
    const uint64_t urngrange = 2147483645u; /* max - min */
    const uint64_t uerange = n;
    const uint64_t scaling = urngrange / uerange;
    const uint64_t past = uerange * scaling;
    uint64_t ret;
    do {
        ret = (uint64_t)minstd_next(state) - 1u;
    } while (ret >= past);
    return (uint32_t)(ret / scaling);
}

/* Philox4x32-10 (Salmon et al. 2011) */
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline float u01_from_bits(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

/* Draw sites of one bounce, in the order the reference consumes them (Appendix B.11):
 * in mode 0 every call is the next minstd draw; in mode 1 the site selects a fixed Philox lane:
 *   block (pixel, sample, 2*bounce  ): [0] alpha coin  [1] strategy coin  [2] vndf u1 | selector  [3] vndf u2 | light index
 *   block (pixel, sample, 2*bounce+1): [0],[1] cosine (z, phi) | light (u, v)
 *   block (pixel, sample, 0xFFFFFFFF): [0],[1] pixel jitter */
typedef struct rng_t {
    int mode;
    uint32_t state; /* mode 0 */
    uint32_t key[2];
    uint32_t pixel, sample, bounce;
    uint32_t blk0[4], blk1[4];
    int have0, have1;
} rng_t;

static void rng_begin_bounce(rng_t *r, uint32_t bounce) {
    r->bounce = bounce;
    r->have0 = r->have1 = 0;
}
static inline uint32_t rng_lane(rng_t *r, int block, int lane) {
    if (block == 0) {
        if (!r->have0) {
            uint32_t c[4] = {r->pixel, r->sample, 2u * r->bounce, 0u};
            philox4x32_10(c, r->key, r->blk0);
            r->have0 = 1;
        }
        return r->blk0[lane];
    }
    if (!r->have1) {
        uint32_t c[4] = {r->pixel, r->sample, 2u * r->bounce + 1u, 0u};
        philox4x32_10(c, r->key, r->blk1);
        r->have1 = 1;
    }
    return r->blk1[lane];
}
/* uniform_real_distribution<float>(0,1) */
static inline float rng_u01(rng_t *r, int block, int lane) {
    if (r->mode == 0) return minstd_canonical(&r->state) * (1.0f - 0.0f) + 0.0f;
    return u01_from_bits(rng_lane(r, block, lane));
}
static inline float rng_uniform(rng_t *r, int block, int lane, float a, float b) {
    if (r->mode == 0) return minstd_canonical(&r->state) * (b - a) + a;
    return u01_from_bits(rng_lane(r, block, lane)) * (b - a) + a;
}
static inline uint32_t rng_int(rng_t *r, int block, int lane, uint32_t n) {
    if (r->mode == 0) return minstd_int(&r->state, n);
    uint32_t i = (uint32_t)(u01_from_bits(rng_lane(r, block, lane)) * (float)n);
    return i < n ? i : n - 1;
}

/* ------------------------------------------------------------------------------------------ */
/* scene access                                                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct ctx_t {
    const rt_scene_desc *sc;
    float gamma_lut[256]; /* powf(k/255, 2.2f): Texture::sample's per-texel pow, geometry.h:525-527,561 */
    uint32_t width, height;
    float tan_half_x, tan_half_y;
} ctx_t;

typedef struct { v3 start, dir; } ray_t;

static inline v3 tri_vert(const rt_scene_desc *sc, uint32_t id, int v) {
    const float *p = sc->tri_pos + (size_t)id * 9 + v * 3;
    return V(p[0], p[1], p[2]);
}
static inline v3 ray_at(ray_t r, float t) { return add(r.start, muls(r.dir, t)); } /* geometry.h:369-371 */

/* ------------------------------------------------------------------------------------------ */
/* intersection (src/bvh.h)                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int has; v3 xs; uint32_t obj; } isect_t; /* intersection_res, bvh.h:31-33 */

/* intersect_ray_triangle + intersect(ray, triangle, min_dst), bvh.h:36-65 */
static inline int intersect_tri(ray_t ray, v3 a, v3 b, v3 c, float min_dst, v3 *out) {
    v3 av = sub(b, a);
    v3 au = sub(c, a);
    v3 at = neg(ray.dir);
    v3 y = sub(ray.start, a);
    v3 xs = divs(V(det(y, au, at), det(av, y, at), det(av, au, y)), det(av, au, at));
    if (xs.x >= 0 && xs.y >= 0 && xs.x + xs.y <= 1 && xs.z >= min_dst) {
        *out = xs;
        return 1;
    }
    return 0;
}

/* intersect(ray, aabb, min_dst), bvh.h:137-152; max_component/min_component are
 * std::max_element / std::min_element over {x,y,z} (geometry.h:42-50) */
static inline int intersect_box(ray_t ray, const rt_bvh_node *nd, float min_dst, float *t_out) {
    v3 i1 = vdiv(sub(V(nd->bmin[0], nd->bmin[1], nd->bmin[2]), ray.start), ray.dir);
    v3 i2 = vdiv(sub(V(nd->bmax[0], nd->bmax[1], nd->bmax[2]), ray.start), ray.dir);
    v3 lo = V(std_min(i1.x, i2.x), std_min(i1.y, i2.y), std_min(i1.z, i2.z));
    v3 hi = V(std_max(i1.x, i2.x), std_max(i1.y, i2.y), std_max(i1.z, i2.z));
    float t_min = lo.x;
    if (t_min < lo.y) t_min = lo.y;
    if (t_min < lo.z) t_min = lo.z;
    float t_max = hi.x;
    if (hi.y < t_max) t_max = hi.y;
    if (hi.z < t_max) t_max = hi.z;
    if (t_min <= t_max && t_max >= min_dst) {
        *t_out = std_max(t_min, min_dst);
        return 1;
    }
    return 0;
}

/* update_intersection, bvh.h:123-135 (max_dst is always INFINITY at the call sites) */
static inline void update_intersection(isect_t *res, const isect_t *in) {
    if (!in->has) return;
    float t = in->xs.z;
    if (t > INFINITY) return;
    if (!res->has || res->xs.z > t) *res = *in;
}

/* BVH::intersect_ray(ray, min_dst, node_id), bvh.h:195-235 */
static isect_t bvh_intersect(const rt_scene_desc *sc, const rt_bvh_desc *bvh, ray_t ray, float min_dst,
                             uint32_t node_id, orc_stats *st) {
    isect_t intr;
    intr.has = 0;
    const rt_bvh_node *node = &bvh->nodes[node_id];
    st->nodes_visited++;
    for (uint32_t k = node->obj_begin; k < node->obj_end; ++k) {
        uint32_t id = bvh->objects[k];
        isect_t cand;
        cand.obj = id;
        st->tri_tests++;
        cand.has = intersect_tri(ray, tri_vert(sc, id, 0), tri_vert(sc, id, 1), tri_vert(sc, id, 2), min_dst, &cand.xs);
        update_intersection(&intr, &cand);
    }
    float d_left = 0, d_right = 0;
    int has_l = 0, has_r = 0;
    if (node->left_child != RT_NO_CHILD) {
        st->box_tests++;
        has_l = intersect_box(ray, &bvh->nodes[node->left_child], min_dst, &d_left);
    }
    if (node->right_child != RT_NO_CHILD) {
        st->box_tests++;
        has_r = intersect_box(ray, &bvh->nodes[node->right_child], min_dst, &d_right);
    }
    if (has_l && has_r) {
        uint32_t id1 = node->left_child, id2 = node->right_child;
        if (d_left > d_right) {
            uint32_t ti = id1; id1 = id2; id2 = ti;
            float tf = d_left; d_left = d_right; d_right = tf;
        }
        isect_t sub1 = bvh_intersect(sc, bvh, ray, min_dst, id1, st);
        update_intersection(&intr, &sub1);
        if (!intr.has || intr.xs.z > d_right) {
            isect_t sub2 = bvh_intersect(sc, bvh, ray, min_dst, id2, st);
            update_intersection(&intr, &sub2);
        }
    } else {
        if (has_l) {
            isect_t s1 = bvh_intersect(sc, bvh, ray, min_dst, node->left_child, st);
            update_intersection(&intr, &s1);
        }
        if (has_r) {
            isect_t s2 = bvh_intersect(sc, bvh, ray, min_dst, node->right_child, st);
            update_intersection(&intr, &s2);
        }
    }
    return intr;
}

/* ------------------------------------------------------------------------------------------ */
/* textures and materials (src/geometry.h:517-631)                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float r, g, b, a; } c4;

static inline float wrap_repeat(float x) { /* geometry.h:517-519: std::fmod(float,int) promotes to double */
    return (float)fmod(fmod((double)x, 1.0) + 1.0, 1.0);
}
static inline uint32_t mod_inc(uint32_t x, uint32_t mod) { return x == mod - 1 ? 0 : x + 1; } /* geometry.h:521-523 */

/* Texture::sample, geometry.h:545-575. tex < 0 selects the built-in 1x1 defaults. gamma_on
 * = (gamma == 2.2f); gamma 1.0 is the identity because powf(x, 1) == x. */
static c4 tex_sample(const ctx_t *cx, int32_t tex, float u, float v, int gamma_on, int normal_default) {
    const rt_scene_desc *sc = cx->sc;
    if (tex < 0) {
        c4 w = {1, 1, 1, 1};              /* WHITE_TEXTURE, geometry.h:601 */
        c4 n = {0.5f, 0.5f, 1, 0};        /* NORMAL_UP, geometry.h:602 */
        return normal_default ? n : w;
    }
    const rt_texture *t = &sc->textures[tex];
    const uint8_t *px0 = sc->texels + t->offset;
    if ((uint64_t)t->width * t->height == 1) { /* data.size() == 1: returned without gamma, geometry.h:548-550 */
        c4 r = {px0[0] / 255.0f, px0[1] / 255.0f, px0[2] / 255.0f, px0[3] / 255.0f};
        return r;
    }
    float tx = wrap_repeat(u) * (float)t->width;
    float ty = wrap_repeat(v) * (float)t->height;
    int px = (int)tx;
    int py = (int)ty;
    float dx = tx - (float)px;
    float dy = ty - (float)py;
    uint32_t x0 = (uint32_t)px, y0 = (uint32_t)py;
    uint32_t x1 = mod_inc(x0, t->width), y1 = mod_inc(y0, t->height);
    const uint64_t n_tex = (uint64_t)t->width * t->height;
    uint64_t idx[4] = {x0 + (uint64_t)y0 * t->width, x0 + (uint64_t)y1 * t->width,
                       x1 + (uint64_t)y0 * t->width, x1 + (uint64_t)y1 * t->width};
    c4 ps[4];
    for (int k = 0; k < 4; ++k) {
        /* wrap_repeat can round up to 1.0f; the reference then reads one row further (or out of
         * bounds). Keep in-bounds texels identical and make the out-of-bounds case defined. */
        const uint8_t *p = px0 + (idx[k] % n_tex) * 4;
        if (gamma_on) {
            ps[k].r = cx->gamma_lut[p[0]];
            ps[k].g = cx->gamma_lut[p[1]];
            ps[k].b = cx->gamma_lut[p[2]];
        } else {
            ps[k].r = p[0] / 255.0f;
            ps[k].g = p[1] / 255.0f;
            ps[k].b = p[2] / 255.0f;
        }
        ps[k].a = p[3] / 255.0f;
    }
    /* (1 - dx) * ((1 - dy) * ps[0][0] + dy * ps[0][1]) + dx * ((1 - dy) * ps[1][0] + dy * ps[1][1]) */
    c4 res;
    const float wx0 = 1 - dx, wy0 = 1 - dy;
#define BIL(ch) res.ch = wx0 * (wy0 * ps[0].ch + dy * ps[1].ch) + dx * (wy0 * ps[2].ch + dy * ps[3].ch)
    BIL(r); BIL(g); BIL(b); BIL(a);
#undef BIL
    return res;
}

/* ray_intersection_info, bvh.h:18-29 */
typedef struct hit_t {
    v3 normal, shading_normal;
    float t;
    uint32_t obj;
    int is_inside;
    c4 color;
    v3 emission;
    float metallic, roughness, ior;
} hit_t;

static inline v3 interop3(float bx, float by, const float *vals) { /* triangle::interop, geometry.h:497-502 */
    float w0 = 1 - bx - by;
    v3 a = V(vals[0], vals[1], vals[2]), b = V(vals[3], vals[4], vals[5]), c = V(vals[6], vals[7], vals[8]);
    return add(add(muls(a, w0), muls(b, bx)), muls(c, by));
}

/* to_intersection_info, bvh.h:80-121 */
static hit_t to_intersection_info(const ctx_t *cx, const isect_t *intr, ray_t ray) {
    const rt_scene_desc *sc = cx->sc;
    const uint32_t id = intr->obj;
    float b = intr->xs.x, c = intr->xs.y, t = intr->xs.z;
    v3 A = tri_vert(sc, id, 0), B = tri_vert(sc, id, 1), C = tri_vert(sc, id, 2);
    v3 normal = norm(crs(sub(B, A), sub(C, A))); /* triangle::normal, geometry.h:477-479 */
    int is_inside = dot(normal, ray.dir) > 0;
    v3 smooth = norm(interop3(b, c, sc->tri_normals + (size_t)id * 9));
    if (dot(normal, smooth) < 0) smooth = neg(smooth);
    const float *uvs = sc->tri_uv + (size_t)id * 6;
    float w0 = 1 - b - c;
    float tu = uvs[0] * w0 + uvs[2] * b + uvs[4] * c;
    float tv = uvs[1] * w0 + uvs[3] * b + uvs[5] * c;
    v3 tangent;
    if (sc->tri_tangents) {
        tangent = norm(interop3(b, c, sc->tri_tangents + (size_t)id * 9));
    } else {
        const float one[9] = {1, 0, 0, 1, 0, 0, 1, 0, 0};
        tangent = norm(interop3(b, c, one));
    }
    v3 bitangent = crs(smooth, tangent);
    const rt_material *m = &sc->materials[sc->tri_material[id]];
    /* material::normal_at -> Texture::sample_normal, geometry.h:577-582,628-630 */
    c4 n01 = tex_sample(cx, m->normal_tex, tu, tv, 0, 1);
    v3 nl = norm(V(n01.r * 2 - 1, n01.g * 2 - 1, n01.b * 2 - 1));
    v3 shading = norm(transform3(nl, tangent, bitangent, smooth));
    /* metallic_roughness_at, color_at, emission_at: geometry.h:615-626 */
    c4 mr = tex_sample(cx, m->metallic_roughness_tex, tu, tv, 0, 0);
    c4 ct = tex_sample(cx, m->color_tex, tu, tv, 1, 0);
    c4 et = tex_sample(cx, m->emissive_tex, tu, tv, 1, 0);
    hit_t h;
    h.normal = is_inside ? neg(normal) : normal;
    h.shading_normal = is_inside ? neg(shading) : shading;
    h.t = t;
    h.obj = id;
    h.is_inside = is_inside;
    h.color.r = m->color[0] * ct.r;
    h.color.g = m->color[1] * ct.g;
    h.color.b = m->color[2] * ct.b;
    h.color.a = m->color[3] * ct.a;
    h.emission = V(m->emission[0] * et.r, m->emission[1] * et.g, m->emission[2] * et.b);
    h.metallic = m->metallic * mr.b;
    h.roughness = m->roughness * mr.g;
    h.ior = m->ior;
    return h;
}

/* ------------------------------------------------------------------------------------------ */
/* sampling distributions and BRDF (src/raytracer.h)                                            */
/* ------------------------------------------------------------------------------------------ */
/* VNDF_dist::choose_local_x, raytracer.h:208-219 */
static v3 choose_local_x(v3 n) {
    v3 res = V(1, 1, 1);
    if (fabsf(n.x) > 0.5f) res.x -= dot(res, n) / n.x;
    else if (fabsf(n.y) > 0.5f) res.y -= dot(res, n) / n.y;
    else res.z -= dot(res, n) / n.z;
    return norm(res);
}

/* VNDF_dist::sample, raytracer.h:140-173 (`roughness` member holds alpha = max(r, 0.04)^2) */
static v3 vndf_sample(rng_t *rng, float alpha, v3 in_dir, v3 normal) {
    v3 nx = choose_local_x(normal);
    v3 ny = crs(normal, nx);
    v3 v = neg(norm(V(dot(nx, in_dir), dot(ny, in_dir), dot(normal, in_dir))));
    v3 vh = norm(mul(V(alpha, alpha, 1), v));
    float lensq = vh.x * vh.x + vh.y * vh.y;
    v3 T1 = lensq > 0 ? divs(V(-vh.y, vh.x, 0), sqrtf(lensq)) : V(1, 0, 0);
    v3 T2 = crs(vh, T1);
    float r = sqrtf(rng_u01(rng, 0, 2));
    float phi = 2.0f * PI_F * rng_u01(rng, 0, 3);
    float t1 = r * cosf(phi);
    float t2 = r * sinf(phi);
    float s = 0.5f * (1.0f + vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - pow2f(t1)) + s * t2;
    v3 nh = transform3(V(t1, t2, sqrtf(std_max(0.0f, 1.0f - pow2f(t1) - pow2f(t2)))), T1, T2, vh);
    v3 ne = norm(V(alpha * nh.x, alpha * nh.y, std_max(0.0f, nh.z)));
    v3 res_n = norm(transform3(ne, nx, ny, normal));
    return reflect(res_n, in_dir);
}

/* VNDF_dist::pdf, raytracer.h:175-206 */
static float vndf_pdf(float alpha, float eps, v3 in_dir, v3 normal, v3 dir) {
    v3 nx = choose_local_x(normal);
    v3 ny = crs(normal, nx);
    v3 v = neg(V(dot(nx, in_dir), dot(ny, in_dir), dot(normal, in_dir)));
    v3 nv = norm(sub(dir, in_dir)); /* halfway, raytracer.h:131-134 */
    v3 n = V(dot(nx, nv), dot(ny, nv), dot(normal, nv));
    float vdn = dot(v, n);
    if (vdn <= 0) return 0;
    float ax = v.x * alpha, ay = v.y * alpha;
    float lambda = (-1 + sqrtf(1 + (ax * ax + ay * ay) / pow2f(v.z))) / 2;
    float g1 = 1 / (1 + lambda);
    v3 nd = vdiv(n, V(alpha, alpha, 1));
    float dn = 1 / PI_F / alpha / alpha / pow2f(len2(nd));
    float dv = g1 * vdn * dn / std_max(eps, v.z);
    return dv / 4 / vdn;
}

/* sphere_uniform_dist::sample + cosine_dist::sample, raytracer.h:94-105,116-121 */
static v3 cosine_sample(rng_t *rng, v3 normal) {
    float z = rng_uniform(rng, 1, 0, -1.0f, 1.0f);
    float co_z = sqrtf(std_max(0.0f, 1 - z * z));
    float phi = rng_uniform(rng, 1, 1, 0.0f, 2 * PI_F);
    v3 s = V(co_z * cosf(phi), co_z * sinf(phi), z);
    return norm(add(normal, s));
}
static float cosine_pdf(v3 normal, v3 dir) { return std_max(dot(normal, dir) / PI_F, 0.0f); } /* raytracer.h:123-128 */

/* bvh_mix_dist::sample -> triangle_dist::sample, raytracer.h:227-239,355-361 */
static v3 light_sample(rng_t *rng, const rt_scene_desc *sc, v3 x) {
    uint32_t k = rng_int(rng, 0, 3, sc->light_bvh.n_objects);
    uint32_t id = sc->light_bvh.objects[k];
    float u = rng_u01(rng, 1, 0);
    float v = rng_u01(rng, 1, 1);
    if (u + v > 1) {
        u = 1 - u;
        v = 1 - v;
    }
    v3 A = tri_vert(sc, id, 0), B = tri_vert(sc, id, 1), C = tri_vert(sc, id, 2);
    v3 p = add(add(A, muls(sub(B, A), v)), muls(sub(C, A), u));
    return norm(sub(p, x));
}

/* BVH::foreach_intersection with the lambda of bvh_mix_dist::pdf inlined, bvh.h:237-260,
 * raytracer.h:79-84,255-261,363-375 */
static void light_pdf_visit(const rt_scene_desc *sc, ray_t ray, float min_dst, uint32_t node_id, float *res,
                            orc_stats *st) {
    const rt_bvh_desc *bvh = &sc->light_bvh;
    const rt_bvh_node *node = &bvh->nodes[node_id];
    st->light_nodes_visited++;
    for (uint32_t k = node->obj_begin; k < node->obj_end; ++k) {
        uint32_t id = bvh->objects[k];
        v3 A = tri_vert(sc, id, 0), B = tri_vert(sc, id, 1), C = tri_vert(sc, id, 2);
        v3 xs;
        st->light_tri_tests++;
        if (intersect_tri(ray, A, B, C, min_dst, &xs)) {
            v3 y = ray_at(ray, xs.z);
            v3 cr = crs(sub(B, A), sub(C, A));
            v3 normal_y = norm(cr);
            v3 dir = norm(sub(y, ray.start));
            float mult = len2(sub(ray.start, y)) / fabsf(dot(dir, normal_y));
            float square = len(cr) / 2;
            *res += mult / square;
        }
    }
    float d;
    if (node->left_child != RT_NO_CHILD) {
        st->light_box_tests++;
        if (intersect_box(ray, &bvh->nodes[node->left_child], min_dst, &d))
            light_pdf_visit(sc, ray, min_dst, node->left_child, res, st);
    }
    if (node->right_child != RT_NO_CHILD) {
        st->light_box_tests++;
        if (intersect_box(ray, &bvh->nodes[node->right_child], min_dst, &d))
            light_pdf_visit(sc, ray, min_dst, node->right_child, res, st);
    }
}
static float light_pdf(const rt_scene_desc *sc, v3 x, v3 dir, orc_stats *st) {
    float res = 0;
    ray_t ray = {x, dir};
    st->light_pdf_rays++;
    if (sc->light_bvh.root != RT_NO_CHILD) light_pdf_visit(sc, ray, sc->eps, sc->light_bvh.root, &res, st);
    return res / (float)(size_t)sc->light_bvh.n_objects;
}

static inline float heaviside(float x) { return x > 0 ? 1 : 0; } /* raytracer.h:264-266 */

/* specular_brdf, raytracer.h:273-293 */
static float specular_brdf(float alpha, v3 in_dir, v3 out_dir, v3 normal) {
    v3 h = norm(sub(out_dir, in_dir));
    float ndh = dot(normal, h);
    float d = pow2f(alpha) * heaviside(ndh) / PI_F / pow2f(pow2f(ndh) * (pow2f(alpha) - 1) + 1);
    float ndo = dot(normal, out_dir);
    float ndi = dot(normal, neg(in_dir));
    float div1 = fabsf(ndo) + sqrtf(pow2f(alpha) + (1 - pow2f(alpha)) * pow2f(ndo));
    float div2 = fabsf(ndi) + sqrtf(pow2f(alpha) + (1 - pow2f(alpha)) * pow2f(ndi));
    float v = heaviside(dot(h, out_dir)) * heaviside(dot(h, neg(in_dir))) / div1 / div2;
    return v * d;
}

/* pbr_brdf = (1-m) dielectric_brdf + m metallic_brdf, raytracer.h:295-343 */
static v3 pbr_brdf(const ctx_t *cx, v3 in_dir, v3 out_dir, const hit_t *h) {
    v3 res = V(0, 0, 0);
    const float alpha = pow2f(std_max(h->roughness, cx->sc->min_roughness));
    v3 color = V(h->color.r, h->color.g, h->color.b);
    if (h->metallic < 1) {
        float spec = specular_brdf(alpha, in_dir, out_dir, h->shading_normal);
        float vdh = dot(neg(in_dir), norm(sub(out_dir, in_dir)));
        float f0 = pow2f((1 - h->ior) / (1 + h->ior));
        float fr = f0 + (1 - f0) * pow5f(1 - fabsf(vdh));
        v3 base = divs(color, PI_F);
        v3 layer = V(spec, spec, spec);
        v3 d = add(muls(base, 1 - fr), muls(layer, fr));
        res = add(res, smul(1 - h->metallic, d));
    }
    if (h->metallic > 0) {
        float spec = specular_brdf(alpha, in_dir, out_dir, h->shading_normal);
        float vdh = dot(neg(in_dir), norm(sub(out_dir, in_dir)));
        float p5 = pow5f(1 - fabsf(vdh));
        /* conductor_fresnel: bsdf * (f0 + (1 - f0) * pow5), raytracer.h:267-271 */
        v3 f = V(color.x + (1 - color.x) * p5, color.y + (1 - color.y) * p5, color.z + (1 - color.z) * p5);
        v3 mtl = mul(V(spec, spec, spec), f);
        res = add(res, smul(h->metallic, mtl));
    }
    return res;
}

/* ------------------------------------------------------------------------------------------ */
/* integrator (src/raytracer.h:540-627)                                                         */
/* ------------------------------------------------------------------------------------------ */
static v3 trace_ray(const ctx_t *cx, rng_t *rng, ray_t ray, unsigned max_depth, uint32_t bounce, orc_stats *st);

/* shade, raytracer.h:555-591 */
static v3 shade(const ctx_t *cx, rng_t *rng, ray_t ray, const hit_t *h, unsigned max_depth, uint32_t bounce,
                orc_stats *st) {
    const rt_scene_desc *sc = cx->sc;
    st->shades++;
    rng_begin_bounce(rng, bounce);
    v3 pos = ray_at(ray, h->t);
    if (!(rng_u01(rng, 0, 0) <= h->color.a)) {
        ray_t next = {pos, ray.dir};
        return trace_ray(cx, rng, next, max_depth, bounce + 1, st);
    }
    const float alpha = pow2f(std_max(h->roughness, sc->min_roughness));
    v3 dir;
    if (rng_u01(rng, 0, 1) <= sc->vndf_factor) {
        dir = vndf_sample(rng, alpha, ray.dir, h->shading_normal);
    } else if (sc->light_bvh.n_objects == 0) {
        dir = cosine_sample(rng, h->normal); /* dir_dist = cosine_dist, raytracer.h:449-453 */
    } else {
        /* mix_dist{cosine, bvh_mix}::sample, raytracer.h:383-392 */
        uint32_t which = rng_int(rng, 0, 2, 2);
        dir = which == 0 ? cosine_sample(rng, h->normal) : light_sample(rng, sc, pos);
    }
    if (isnan(dir.x) || isnan(dir.y) || isnan(dir.z)) return h->emission;
    float vndf_p = vndf_pdf(alpha, sc->eps, ray.dir, h->shading_normal, dir);
    float mis_p;
    if (sc->light_bvh.n_objects == 0) {
        mis_p = cosine_pdf(h->normal, dir);
    } else { /* mix_dist::pdf, raytracer.h:395-407 */
        float res = 0;
        res += cosine_pdf(h->normal, dir);
        res += light_pdf(sc, pos, dir, st);
        mis_p = res / 2;
    }
    float p = sc->vndf_factor * vndf_p + (1 - sc->vndf_factor) * mis_p;
    if (p < sc->eps) return h->emission;
    v3 scl = muls(divs(pbr_brdf(cx, ray.dir, dir, h), p), std_max(0.0f, dot(dir, h->shading_normal)));
    if (len2(scl) == 0.0f) return h->emission;
    ray_t next = {pos, dir};
    v3 clr = mul(trace_ray(cx, rng, next, max_depth, bounce + 1, st), scl);
    return add(h->emission, clr);
}

/* trace_ray + cast_ray, raytracer.h:540-553,593-605 */
static v3 trace_ray(const ctx_t *cx, rng_t *rng, ray_t ray, unsigned max_depth, uint32_t bounce, orc_stats *st) {
    const rt_scene_desc *sc = cx->sc;
    if (max_depth == 0) return V(0, 0, 0);
    st->extension_rays++;
    isect_t res;
    res.has = 0;
    if (sc->scene_bvh.root != RT_NO_CHILD) res = bvh_intersect(sc, &sc->scene_bvh, ray, sc->eps, sc->scene_bvh.root, st);
    if (res.has) {
        hit_t h = to_intersection_info(cx, &res, ray);
        return shade(cx, rng, ray, &h, max_depth - 1, bounce, st);
    }
    /* Scene::bg_at, scene.h:83-89.  At HEAD `bg` is the 1x1 white texture (USE_ENV_MAP = false, config.h:36):
     * bg_color * (1,1,1).  With an environment map: bg_color * bg.sample({x, y}, 2.2f).rgb() where
     *   x = 0.5 + 0.5 * atan2(dir.z, dir.x) / pi_v<float>,  y = 0.5 - asin(dir.y) / pi_v<float>
     * — the literals 0.5 are doubles: x is evaluated in double from the float atan2 (0.5 * atan2 promotes first), y
     * divides float by float and only then promotes for the subtraction; both are rounded to float on assignment. */
    v3 bgc = V(sc->bg_color[0], sc->bg_color[1], sc->bg_color[2]);
    if (sc->env_texture == 0) return mul(bgc, V(1, 1, 1));
    {
        const float pi_f = 3.14159265358979323846f;
        float x = (float)(0.5 + 0.5 * (double)atan2f(ray.dir.z, ray.dir.x) / (double)pi_f);
        float y = (float)(0.5 - (double)(asinf(ray.dir.y) / pi_f));
        c4 c = tex_sample(cx, (int32_t)sc->env_texture - 1, x, y, 1, 0);
        return mul(bgc, V(c.r, c.g, c.b));
    }
}

/* gen_ray (jittered), raytracer.h:527-538 */
static ray_t gen_ray_jitter(const ctx_t *cx, float ox, float oy, int x, int y) {
    const rt_camera *cam = &cx->sc->camera;
    float a = (2 * ((float)x + ox) / (float)cx->width - 1) * cx->tan_half_x;
    float b = (2 * ((float)y + oy) / (float)cx->height - 1) * cx->tan_half_y;
    v3 right = V(cam->right[0], cam->right[1], cam->right[2]);
    v3 up = V(cam->up[0], cam->up[1], cam->up[2]);
    v3 fwd = V(cam->forward[0], cam->forward[1], cam->forward[2]);
    v3 dir = norm(add(sub(smul(a, right), smul(b, up)), smul(1, fwd)));
    ray_t r = {V(cam->position[0], cam->position[1], cam->position[2]), dir};
    return r;
}
/* gen_ray (pixel centre), raytracer.h:516-525: the scalar factors are evaluated in double */
static ray_t gen_ray_centre(const ctx_t *cx, int x, int y) {
    const rt_camera *cam = &cx->sc->camera;
    float a = (float)((2 * (x + 0.5) / cx->width - 1) * (double)cx->tan_half_x);
    float b = (float)((2 * (y + 0.5) / cx->height - 1) * (double)cx->tan_half_y);
    v3 right = V(cam->right[0], cam->right[1], cam->right[2]);
    v3 up = V(cam->up[0], cam->up[1], cam->up[2]);
    v3 fwd = V(cam->forward[0], cam->forward[1], cam->forward[2]);
    v3 dir = norm(add(sub(smul(a, right), smul(b, up)), smul(1, fwd)));
    ray_t r = {V(cam->position[0], cam->position[1], cam->position[2]), dir};
    return r;
}

static void ctx_init(ctx_t *cx, const rt_scene_desc *sc, uint32_t w, uint32_t h) {
    cx->sc = sc;
    cx->width = w;
    cx->height = h;
    for (int k = 0; k < 256; ++k) cx->gamma_lut[k] = powf((float)k / 255.0f, 2.2f);
    /* tan(fov_x/2) and Camera::fov_y, scene.h:69-71: float overloads (stb_image.h pulls <math.h>) */
    cx->tan_half_x = tanf(sc->camera.fov_x / 2);
    float fov_y = atanf(tanf(sc->camera.fov_x / 2) * (float)h / (float)w) * 2;
    cx->tan_half_y = tanf(fov_y / 2);
}

/* ------------------------------------------------------------------------------------------ */
/* drivers                                                                                      */
/* ------------------------------------------------------------------------------------------ */
#define SPAN_SIZE 256u /* config.h:13 */

typedef struct job_t {
    ctx_t cx;
    const orc_params *p;
    float *out;
    atomic_uint next_span;
    uint32_t span_count;
    orc_stats total;
    pthread_mutex_t lock;
} job_t;

static void stats_add(orc_stats *a, const orc_stats *b) {
    uint64_t *pa = (uint64_t *)a;
    const uint64_t *pb = (const uint64_t *)b;
    for (size_t i = 0; i < sizeof(orc_stats) / sizeof(uint64_t); ++i) pa[i] += pb[i];
}

/* render_pixel + sanitize_nans, raytracer.h:607-627, inside the span loop of run_raytracer :646-659 */
static void *worker(void *arg) {
    job_t *job = (job_t *)arg;
    const orc_params *p = job->p;
    const ctx_t *cx = &job->cx;
    const uint32_t n_pix = p->width * p->height;
    const uint32_t s_begin = p->rng_mode == 0 ? 0 : p->sample_begin;
    const uint32_t s_end = p->rng_mode == 0 ? p->samples : (p->sample_end ? p->sample_end : p->samples);
    orc_stats st;
    memset(&st, 0, sizeof st);
    for (;;) {
        uint32_t span = atomic_fetch_add(&job->next_span, 1);
        if (span >= job->span_count) break;
        rng_t rng;
        memset(&rng, 0, sizeof rng);
        rng.mode = (int)p->rng_mode;
        rng.state = minstd_seed((int32_t)span + p->seed_offset);
        rng.key[0] = (uint32_t)p->seed;
        rng.key[1] = (uint32_t)(p->seed >> 32);
        uint32_t begin = SPAN_SIZE * span, end = begin + SPAN_SIZE < n_pix ? begin + SPAN_SIZE : n_pix;
        for (uint32_t pix = begin; pix < end; ++pix) {
            int x = (int)(pix % p->width), y = (int)(pix / p->width);
            v3 res = V(0, 0, 0);
            for (uint32_t s = s_begin; s < s_end; ++s) {
                rng.pixel = pix;
                rng.sample = s;
                float ox, oy;
                if (rng.mode == 0) {
                    ox = rng_u01(&rng, 0, 0);
                    oy = rng_u01(&rng, 0, 0);
                } else {
                    uint32_t c[4] = {pix, s, 0xFFFFFFFFu, 0u}, o[4];
                    philox4x32_10(c, rng.key, o);
                    ox = u01_from_bits(o[0]);
                    oy = u01_from_bits(o[1]);
                }
                ray_t ray = gen_ray_jitter(cx, ox, oy, x, y);
                v3 c = trace_ray(cx, &rng, ray, cx->sc->ray_depth, 0, &st);
                if (isnan(c.x)) c.x = 0;
                if (isnan(c.y)) c.y = 0;
                if (isnan(c.z)) c.z = 0;
                res = add(res, c);
                st.samples++;
            }
            res = divs(res, (float)p->samples);
            job->out[(size_t)pix * 3 + 0] = res.x;
            job->out[(size_t)pix * 3 + 1] = res.y;
            job->out[(size_t)pix * 3 + 2] = res.z;
        }
    }
    pthread_mutex_lock(&job->lock);
    stats_add(&job->total, &st);
    pthread_mutex_unlock(&job->lock);
    return NULL;
}

ORC_EXPORT int orc_render(const rt_scene_desc *sc, const orc_params *p, float *rgb_mean, orc_stats *stats) {
    if (!sc || !p || !rgb_mean || p->width == 0 || p->height == 0 || p->samples == 0) return -1;
    job_t job;
    memset(&job, 0, sizeof job);
    ctx_init(&job.cx, sc, p->width, p->height);
    job.p = p;
    job.out = rgb_mean;
    atomic_init(&job.next_span, 0);
    job.span_count = (p->width * p->height + SPAN_SIZE - 1) / SPAN_SIZE;
    pthread_mutex_init(&job.lock, NULL);
    if (sc->ray_depth == 0) { /* run_raytracer returns early, raytracer.h:630 */
        memset(rgb_mean, 0, (size_t)p->width * p->height * 3 * sizeof(float));
    } else {
        uint32_t nt = p->n_threads ? p->n_threads : 1;
        if (nt > 256) nt = 256;
        pthread_t th[256];
        for (uint32_t i = 0; i < nt; ++i) pthread_create(&th[i], NULL, worker, &job);
        for (uint32_t i = 0; i < nt; ++i) pthread_join(th[i], NULL);
    }
    if (stats) *stats = job.total;
    pthread_mutex_destroy(&job.lock);
    return 0;
}

/* Primary ids with pixel-centre rays (gen_ray(camera,x,y) + cast_ray); hitinfo (nullable) gets
 * 18 floats per pixel in the layout of ref_tool's `hitinfo` command. */
ORC_EXPORT int orc_primary_ids(const rt_scene_desc *sc, uint32_t width, uint32_t height, int32_t *ids, float *hitinfo) {
    if (!sc || !ids || width == 0 || height == 0) return -1;
    ctx_t cx;
    ctx_init(&cx, sc, width, height);
    orc_stats st;
    memset(&st, 0, sizeof st);
    if (hitinfo) memset(hitinfo, 0, (size_t)width * height * 18 * sizeof(float));
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            size_t p = (size_t)y * width + x;
            ray_t ray = gen_ray_centre(&cx, (int)x, (int)y);
            isect_t res;
            res.has = 0;
            if (sc->scene_bvh.root != RT_NO_CHILD)
                res = bvh_intersect(sc, &sc->scene_bvh, ray, sc->eps, sc->scene_bvh.root, &st);
            ids[p] = res.has ? (int32_t)res.obj : -1;
            if (res.has && hitinfo) {
                hit_t h = to_intersection_info(&cx, &res, ray);
                float *o = hitinfo + p * 18;
                o[0] = h.t;
                o[1] = h.normal.x; o[2] = h.normal.y; o[3] = h.normal.z;
                o[4] = h.shading_normal.x; o[5] = h.shading_normal.y; o[6] = h.shading_normal.z;
                o[7] = h.color.r; o[8] = h.color.g; o[9] = h.color.b; o[10] = h.color.a;
                o[11] = h.emission.x; o[12] = h.emission.y; o[13] = h.emission.z;
                o[14] = h.metallic; o[15] = h.roughness; o[16] = h.is_inside ? 1.0f : 0.0f; o[17] = h.ior;
            }
        }
    return 0;
}

/* Image::set_pixel -> convert_color, image.h:40-82 */
ORC_EXPORT void orc_tonemap_rgb8(const float *rgb, size_t n_pixels, uint8_t *out) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    const float inv_gamma = 1 / 2.2f;
    for (size_t i = 0; i < n_pixels * 3; ++i) {
        float x = rgb[i];
        float m = (x * (a * x + b)) / (x * (c * x + d) + e);
        float v = powf(m, inv_gamma) * 255;
        /* std::clamp(v, 0, 255) then std::round, image.h:66-69 */
        float cl = (v < 0.0f) ? 0.0f : (255.0f < v) ? 255.0f : v;
        out[i] = (cl != cl) ? 0 : (uint8_t)roundf(cl);
    }
}

/* Same layout as `ref_tool rng`: n x {u01, u(-1,1), u(0,2pi), int(0..6), int(0..0), int(0..1)} */
ORC_EXPORT void orc_rng_stream(int32_t seed, uint32_t n, float *out) {
    uint32_t st = minstd_seed(seed);
    for (uint32_t i = 0; i < n; ++i) {
        out[i * 6 + 0] = minstd_canonical(&st) * (1.0f - 0.0f) + 0.0f;
        out[i * 6 + 1] = minstd_canonical(&st) * (1.0f - -1.0f) + -1.0f;
        out[i * 6 + 2] = minstd_canonical(&st) * (2 * PI_F - 0.0f) + 0.0f;
        out[i * 6 + 3] = (float)minstd_int(&st, 7);
        out[i * 6 + 4] = (float)minstd_int(&st, 1);
        out[i * 6 + 5] = (float)minstd_int(&st, 2);
    }
}

ORC_EXPORT void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
ORC_EXPORT float orc_u01(uint32_t bits) { return u01_from_bits(bits); }
