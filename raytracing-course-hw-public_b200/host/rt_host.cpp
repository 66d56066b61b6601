// rt_host.cpp — host-side helpers of include/rt_host.h (no GPU code, no integrator).
//
// The BVH build is an independent restatement of the reference's host build
// (src/bvh.h:262-393) on index arrays: it reproduces the reference's tree node for node
// (same std::sort comparator outcomes, same cost expression with its quirks, same
// pre-order node numbering), so that traversal order and therefore tie-breaking match.
#include "rt_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// RTSC container
// ---------------------------------------------------------------------------------------------
struct RtscHeader {
    char magic[8];  // "RTSC0001"
    uint32_t abi_version;
    uint32_t n_tris;
    rt_camera camera;  // 13 floats
    float bg_color[3];
    float eps, min_roughness, vndf_factor;
    uint32_t ray_depth;
    uint32_t n_materials;
    uint32_t n_textures;
    uint32_t has_tangents;
    uint32_t env_texture;  // was padding: 0 in older files = constant sky
    uint64_t texel_bytes;
    uint32_t scene_n_nodes, scene_root, scene_n_objects;
    uint32_t light_n_nodes, light_root, light_n_objects;
};
static_assert(sizeof(RtscHeader) == 144, "RTSC header layout is part of the file format");
static_assert(sizeof(rt_bvh_node) == 40, "rt_bvh_node mirrors BVHNode (bvh.h:157)");
static_assert(sizeof(rt_material) == 56, "rt_material layout");
static_assert(sizeof(rt_texture) == 16, "rt_texture layout");

constexpr size_t kAlign = 16;
inline size_t pad(size_t n) { return (n + kAlign - 1) / kAlign * kAlign; }

struct Section {
    const void *ptr;
    size_t bytes;
};

std::vector<Section> sections_of(const rt_scene_desc *s) {
    const size_t n = s->n_tris;
    return {
        {s->tri_pos, n * 9 * sizeof(float)},
        {s->tri_normals, n * 9 * sizeof(float)},
        {s->tri_uv, n * 6 * sizeof(float)},
        {s->tri_tangents, s->tri_tangents ? n * 9 * sizeof(float) : 0},
        {s->tri_material, n * sizeof(uint32_t)},
        {s->materials, s->n_materials * sizeof(rt_material)},
        {s->textures, s->n_textures * sizeof(rt_texture)},
        {s->texels, (size_t)s->texel_bytes},
        {s->scene_bvh.nodes, s->scene_bvh.n_nodes * sizeof(rt_bvh_node)},
        {s->scene_bvh.objects, s->scene_bvh.n_objects * sizeof(uint32_t)},
        {s->light_bvh.nodes, s->light_bvh.n_nodes * sizeof(rt_bvh_node)},
        {s->light_bvh.objects, s->light_bvh.n_objects * sizeof(uint32_t)},
    };
}

int validate_bvh(const rt_bvh_desc &b, uint32_t n_tris) {
    if (b.root == RT_NO_CHILD) return b.n_objects == 0 ? RT_OK : RT_ERR_BAD_SCENE;
    if (b.root >= b.n_nodes) return RT_ERR_BAD_SCENE;
    if ((b.n_nodes && !b.nodes) || (b.n_objects && !b.objects)) return RT_ERR_BAD_SCENE;
    for (uint32_t i = 0; i < b.n_nodes; ++i) {
        const rt_bvh_node &nd = b.nodes[i];
        if (nd.left_child != RT_NO_CHILD && nd.left_child >= b.n_nodes) return RT_ERR_BAD_SCENE;
        if (nd.right_child != RT_NO_CHILD && nd.right_child >= b.n_nodes) return RT_ERR_BAD_SCENE;
        if (nd.obj_begin > nd.obj_end || nd.obj_end > b.n_objects) return RT_ERR_BAD_SCENE;
    }
    for (uint32_t i = 0; i < b.n_objects; ++i)
        if (b.objects[i] >= n_tris) return RT_ERR_BAD_SCENE;
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// BVH build
// ---------------------------------------------------------------------------------------------
struct Box {
    float lo[3] = {INFINITY, INFINITY, INFINITY};      // aabb default, geometry.h:380-381
    float hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    void extend(const Box &o) {                         // geometry.h:398-401
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], o.lo[k]);
            hi[k] = std::max(hi[k], o.hi[k]);
        }
    }
    // aabb::surface_area, geometry.h:419-421: 2*dot(d, d.yxz) -- NOT the true area; kept as is.
    float surface_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2 * (dx * dy + dy * dx + dz * dz);
    }
};

struct Builder {
    const std::vector<Box> &tri_box;
    const std::vector<float> &center;  // n*3, triangle::center geometry.h:485-487
    std::vector<uint32_t> &objs;
    std::vector<rt_bvh_node> nodes;
    std::vector<float> pref, suf;
    uint32_t min_node_size;

    Box box_of(uint32_t b, uint32_t e) const {  // bounding_box_of, bvh.h:315-321
        Box r;
        for (uint32_t i = b; i < e; ++i) r.extend(tri_box[objs[i]]);
        return r;
    }

    // split_node, bvh.h:268-313. Returns the split position in [b, e]; e means "no split".
    uint32_t split(uint32_t b, uint32_t e, const Box &box) {
        const float dx = box.hi[0] - box.lo[0], dy = box.hi[1] - box.lo[1], dz = box.hi[2] - box.lo[2];
        const int axis = (dx >= dy && dx >= dz) ? 0 : (dy >= dz ? 1 : 2);
        const float *c = center.data();
        std::sort(objs.begin() + b, objs.begin() + e,
                  [c, axis](uint32_t l, uint32_t r) { return c[l * 3 + axis] < c[r * 3 + axis]; });
        const uint32_t n = e - b;
        pref.clear();
        suf.clear();
        Box acc;
        pref.push_back(acc.surface_area());
        for (uint32_t i = 0; i < n; ++i) {
            acc.extend(tri_box[objs[b + i]]);
            pref.push_back(acc.surface_area());
        }
        acc = Box();
        suf.push_back(acc.surface_area());
        for (uint32_t i = n; i-- > 0;) {
            acc.extend(tri_box[objs[b + i]]);
            suf.push_back(acc.surface_area());
        }
        uint32_t best = e;
        float best_score = (float)(size_t)n * acc.surface_area();
        for (uint32_t i = 1; i < n; ++i) {
            // bvh.h:303-304: left weight i with the area of i+1 objects (sic)
            const float score = (float)(int)i * pref[i + 1] + (float)(size_t)(n - i) * suf[n - i];
            if (score < best_score) {
                best_score = score;
                best = b + i;
            }
        }
        return best;
    }

    // build_node, bvh.h:323-366 (pre-order numbering: parent, left subtree, right subtree)
    uint32_t build(uint32_t b, uint32_t e, const Box &box, uint32_t depth_left) {
        auto leaf = [&]() {
            rt_bvh_node nd;
            std::memcpy(nd.bmin, box.lo, sizeof nd.bmin);
            std::memcpy(nd.bmax, box.hi, sizeof nd.bmax);
            nd.left_child = nd.right_child = RT_NO_CHILD;
            nd.obj_begin = b;
            nd.obj_end = e;
            nodes.push_back(nd);
            return (uint32_t)nodes.size() - 1;
        };
        if (depth_left == 0) return leaf();
        const uint32_t mid = split(b, e, box);
        const uint32_t nl = mid - b, nr = e - mid;
        if (nl == 0 || nr == 0 || (nl < min_node_size && nr < min_node_size)) return leaf();
        rt_bvh_node nd;
        std::memcpy(nd.bmin, box.lo, sizeof nd.bmin);
        std::memcpy(nd.bmax, box.hi, sizeof nd.bmax);
        nd.left_child = nd.right_child = RT_NO_CHILD;
        nd.obj_begin = nd.obj_end = 0;
        const uint32_t idx = (uint32_t)nodes.size();
        nodes.push_back(nd);
        const uint32_t l = build(b, mid, box_of(b, mid), depth_left - 1);
        const uint32_t r = build(mid, e, box_of(mid, e), depth_left - 1);
        nodes[idx].left_child = l;
        nodes[idx].right_child = r;
        return idx;
    }
};

inline float aces(float x) {  // Image::aces_tonemap, image.h:51-59
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return (x * (a * x + b)) / (x * (c * x + d) + e);
}

}  // namespace

extern "C" {

int rt_scene_validate(const rt_scene_desc *s) {
    if (!s || s->abi_version != RT_GPU_ABI_VERSION) return RT_ERR_INVALID_ARG;
    if (s->n_tris && (!s->tri_pos || !s->tri_normals || !s->tri_uv || !s->tri_material)) return RT_ERR_BAD_SCENE;
    if (s->n_materials && !s->materials) return RT_ERR_BAD_SCENE;
    if (s->n_textures && (!s->textures || !s->texels)) return RT_ERR_BAD_SCENE;
    if (s->env_texture > s->n_textures) return RT_ERR_BAD_SCENE;
    for (uint32_t i = 0; i < s->n_tris; ++i)
        if (s->tri_material[i] >= s->n_materials) return RT_ERR_BAD_SCENE;
    for (uint32_t i = 0; i < s->n_materials; ++i) {
        const int32_t t[4] = {s->materials[i].color_tex, s->materials[i].emissive_tex,
                              s->materials[i].metallic_roughness_tex, s->materials[i].normal_tex};
        for (int32_t v : t)
            if (v < -1 || v >= (int32_t)s->n_textures) return RT_ERR_BAD_SCENE;
    }
    for (uint32_t i = 0; i < s->n_textures; ++i) {
        const rt_texture &t = s->textures[i];
        if (t.width == 0 || t.height == 0) return RT_ERR_BAD_SCENE;
        if (t.offset + (uint64_t)t.width * t.height * 4 > s->texel_bytes) return RT_ERR_BAD_SCENE;
    }
    if (int rc = validate_bvh(s->scene_bvh, s->n_tris)) return rc;
    if (int rc = validate_bvh(s->light_bvh, s->n_tris)) return rc;
    return RT_OK;
}

int rt_scene_save(const rt_scene_desc *s, const char *path) {
    if (!s || !path) return RT_ERR_INVALID_ARG;
    RtscHeader h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.magic, "RTSC0001", 8);
    h.abi_version = s->abi_version;
    h.n_tris = s->n_tris;
    h.camera = s->camera;
    std::memcpy(h.bg_color, s->bg_color, sizeof h.bg_color);
    h.eps = s->eps;
    h.min_roughness = s->min_roughness;
    h.vndf_factor = s->vndf_factor;
    h.ray_depth = s->ray_depth;
    h.n_materials = s->n_materials;
    h.n_textures = s->n_textures;
    h.has_tangents = s->tri_tangents ? 1u : 0u;
    h.env_texture = s->env_texture;
    h.texel_bytes = s->texel_bytes;
    h.scene_n_nodes = s->scene_bvh.n_nodes;
    h.scene_root = s->scene_bvh.root;
    h.scene_n_objects = s->scene_bvh.n_objects;
    h.light_n_nodes = s->light_bvh.n_nodes;
    h.light_root = s->light_bvh.root;
    h.light_n_objects = s->light_bvh.n_objects;
    FILE *f = std::fopen(path, "wb");
    if (!f) return RT_ERR_INVALID_ARG;
    static const char zeros[kAlign] = {0};
    bool ok = std::fwrite(&h, sizeof h, 1, f) == 1;
    for (const Section &sec : sections_of(s)) {
        if (sec.bytes && ok) ok = std::fwrite(sec.ptr, 1, sec.bytes, f) == sec.bytes;
        const size_t p = pad(sec.bytes) - sec.bytes;
        if (p && ok) ok = std::fwrite(zeros, 1, p, f) == p;
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? RT_OK : RT_ERR_INVALID_ARG;
}

int rt_scene_load(const char *path, rt_scene_desc **out) {
    if (!path || !out) return RT_ERR_INVALID_ARG;
    *out = nullptr;
    FILE *f = std::fopen(path, "rb");
    if (!f) return RT_ERR_INVALID_ARG;
    RtscHeader h;
    if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, "RTSC0001", 8) != 0) {
        std::fclose(f);
        return RT_ERR_BAD_SCENE;
    }
    rt_scene_desc d;
    std::memset(&d, 0, sizeof d);
    d.abi_version = h.abi_version;
    d.n_tris = h.n_tris;
    d.camera = h.camera;
    std::memcpy(d.bg_color, h.bg_color, sizeof d.bg_color);
    d.eps = h.eps;
    d.min_roughness = h.min_roughness;
    d.vndf_factor = h.vndf_factor;
    d.ray_depth = h.ray_depth;
    d.n_materials = h.n_materials;
    d.n_textures = h.n_textures;
    d.env_texture = h.env_texture;
    d.texel_bytes = h.texel_bytes;
    d.scene_bvh.n_nodes = h.scene_n_nodes;
    d.scene_bvh.root = h.scene_root;
    d.scene_bvh.n_objects = h.scene_n_objects;
    d.light_bvh.n_nodes = h.light_n_nodes;
    d.light_bvh.root = h.light_root;
    d.light_bvh.n_objects = h.light_n_objects;
    // dummy non-null tangents pointer so sections_of() sizes the section
    d.tri_tangents = h.has_tangents ? reinterpret_cast<const float *>(&d) : nullptr;
    std::vector<Section> secs = sections_of(&d);
    size_t total = pad(sizeof(rt_scene_desc));
    for (const Section &sec : secs) total += pad(sec.bytes);
    char *block = static_cast<char *>(std::malloc(total));
    if (!block) {
        std::fclose(f);
        return RT_ERR_OOM;
    }
    const void *ptrs[12];
    size_t off = pad(sizeof(rt_scene_desc));
    bool ok = true;
    for (size_t i = 0; i < secs.size(); ++i) {
        const size_t padded = pad(secs[i].bytes);
        ptrs[i] = secs[i].bytes ? block + off : nullptr;
        if (padded && ok) ok = std::fread(block + off, 1, padded, f) == padded;
        off += padded;
    }
    std::fclose(f);
    if (!ok) {
        std::free(block);
        return RT_ERR_BAD_SCENE;
    }
    d.tri_pos = static_cast<const float *>(ptrs[0]);
    d.tri_normals = static_cast<const float *>(ptrs[1]);
    d.tri_uv = static_cast<const float *>(ptrs[2]);
    d.tri_tangents = static_cast<const float *>(ptrs[3]);
    d.tri_material = static_cast<const uint32_t *>(ptrs[4]);
    d.materials = static_cast<const rt_material *>(ptrs[5]);
    d.textures = static_cast<const rt_texture *>(ptrs[6]);
    d.texels = static_cast<const uint8_t *>(ptrs[7]);
    d.scene_bvh.nodes = static_cast<const rt_bvh_node *>(ptrs[8]);
    d.scene_bvh.objects = static_cast<const uint32_t *>(ptrs[9]);
    d.light_bvh.nodes = static_cast<const rt_bvh_node *>(ptrs[10]);
    d.light_bvh.objects = static_cast<const uint32_t *>(ptrs[11]);
    std::memcpy(block, &d, sizeof d);
    rt_scene_desc *res = reinterpret_cast<rt_scene_desc *>(block);
    if (int rc = rt_scene_validate(res)) {
        std::free(block);
        return rc;
    }
    *out = res;
    return RT_OK;
}

void rt_scene_free(rt_scene_desc *scene) { std::free(scene); }

int rt_host_build_bvh(const float *tri_pos, uint32_t n_tris, const uint8_t *select, uint32_t min_node_size,
                      uint32_t max_depth, rt_bvh_build *out) {
    if (!out || (n_tris && !tri_pos)) return RT_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof *out);
    if (n_tris == 0) {  // bvh.h:373-376
        out->root = RT_NO_CHILD;
        return RT_OK;
    }
    std::vector<Box> tri_box(n_tris);
    std::vector<float> center((size_t)n_tris * 3);
    std::vector<uint32_t> objs;
    objs.reserve(n_tris);
    for (uint32_t i = 0; i < n_tris; ++i) {
        const float *p = tri_pos + (size_t)i * 9;
        Box &bx = tri_box[i];
        for (int v = 0; v < 3; ++v)
            for (int k = 0; k < 3; ++k) {
                bx.lo[k] = std::min(bx.lo[k], p[v * 3 + k]);
                bx.hi[k] = std::max(bx.hi[k], p[v * 3 + k]);
            }
        for (int k = 0; k < 3; ++k) center[(size_t)i * 3 + k] = (p[k] + p[3 + k] + p[6 + k]) / 3;
        if (!select || select[i]) objs.push_back(i);
    }
    Builder b{tri_box, center, objs, {}, {}, {}, min_node_size};
    b.pref.reserve(objs.size() + 1);
    b.suf.reserve(objs.size() + 1);
    const uint32_t root = b.build(0, (uint32_t)objs.size(), b.box_of(0, (uint32_t)objs.size()), max_depth);
    out->n_nodes = (uint32_t)b.nodes.size();
    out->root = root;
    out->n_objects = (uint32_t)objs.size();
    out->nodes = static_cast<rt_bvh_node *>(std::malloc(std::max<size_t>(1, b.nodes.size()) * sizeof(rt_bvh_node)));
    out->objects = static_cast<uint32_t *>(std::malloc(std::max<size_t>(1, objs.size()) * sizeof(uint32_t)));
    if (!out->nodes || !out->objects) {
        rt_host_free_bvh(out);
        return RT_ERR_OOM;
    }
    std::memcpy(out->nodes, b.nodes.data(), b.nodes.size() * sizeof(rt_bvh_node));
    std::memcpy(out->objects, objs.data(), objs.size() * sizeof(uint32_t));
    return RT_OK;
}

void rt_host_free_bvh(rt_bvh_build *bvh) {
    if (!bvh) return;
    std::free(bvh->nodes);
    std::free(bvh->objects);
    bvh->nodes = nullptr;
    bvh->objects = nullptr;
}

int rt_text_scene_parse(const char *path, rt_text_scene **out) {
    if (!path || !out) return RT_ERR_INVALID_ARG;
    *out = nullptr;
    std::FILE *f = std::fopen(path, "r");
    if (!f) return RT_ERR_INVALID_ARG;
    rt_text_scene sc;
    std::memset(&sc, 0, sizeof sc);
    sc.abi_version = RT_GPU_ABI_VERSION;
    sc.ray_depth = 1;  // Scene defaults, scene.h:76-77
    sc.samples = 1;
    sc.eps = 1e-4f;
    sc.camera.right[0] = sc.camera.up[1] = 1.0f;
    sc.camera.forward[2] = -1.0f;
    sc.camera.fov_x = 1.5708f;
    std::vector<rt_text_prim> prims;
    std::vector<rt_text_light> lights;
    bool has_samples = false, has_whitted = false;
    int rc = RT_OK;
    char line[1024];
    enum { NONE, PRIM, LIGHT } cur = NONE;
    auto floats = [](const char *s, float *dst, int n) {
        for (int i = 0; i < n; ++i) {
            char *end = nullptr;
            dst[i] = std::strtof(s, &end);
            if (end == s) return false;
            s = end;
        }
        return true;
    };
    while (rc == RT_OK && std::fgets(line, sizeof line, f)) {
        char cmd[64];
        int used = 0;
        if (std::sscanf(line, "%63s%n", cmd, &used) != 1) continue;  // blank line
        const char *args = line + used;
        const std::string c(cmd);
        float v[9];
        bool ok = true;
        if (c == "DIMENSIONS") { ok = floats(args, v, 2); sc.width = (uint32_t)v[0]; sc.height = (uint32_t)v[1]; }
        else if (c == "RAY_DEPTH") { ok = floats(args, v, 1); sc.ray_depth = (uint32_t)v[0]; has_whitted = true; }
        else if (c == "SAMPLES") { ok = floats(args, v, 1); sc.samples = (uint32_t)v[0]; has_samples = true; }
        else if (c == "BG_COLOR") ok = floats(args, sc.bg_color, 3);
        else if (c == "AMBIENT_LIGHT") { ok = floats(args, sc.ambient, 3); has_whitted = true; }
        else if (c == "CAMERA_POSITION") ok = floats(args, sc.camera.position, 3);
        else if (c == "CAMERA_RIGHT") ok = floats(args, sc.camera.right, 3);
        else if (c == "CAMERA_UP") ok = floats(args, sc.camera.up, 3);
        else if (c == "CAMERA_FORWARD") ok = floats(args, sc.camera.forward, 3);
        else if (c == "CAMERA_FOV_X") ok = floats(args, &sc.camera.fov_x, 1);
        else if (c == "NEW_PRIMITIVE") {
            rt_text_prim p;
            std::memset(&p, 0, sizeof p);
            p.kind = RT_PRIM_PLANE;
            p.param[1] = 1.0f;
            p.rotation[3] = 1.0f;
            p.ior = 1.5f;  // material::ior default, geometry.h:609
            prims.push_back(p);
            cur = PRIM;
        } else if (c == "NEW_LIGHT") {
            rt_text_light l;
            std::memset(&l, 0, sizeof l);
            l.attenuation[0] = 1.0f;
            lights.push_back(l);
            cur = LIGHT;
            has_whitted = true;
        } else if (cur == PRIM && c == "PLANE") { prims.back().kind = RT_PRIM_PLANE; ok = floats(args, prims.back().param, 3); }
        else if (cur == PRIM && c == "ELLIPSOID") { prims.back().kind = RT_PRIM_ELLIPSOID; ok = floats(args, prims.back().param, 3); }
        else if (cur == PRIM && c == "BOX") { prims.back().kind = RT_PRIM_BOX; ok = floats(args, prims.back().param, 3); }
        else if (cur == PRIM && c == "TRIANGLE") { prims.back().kind = RT_PRIM_TRIANGLE; ok = floats(args, prims.back().param, 9); }
        else if (cur == PRIM && c == "POSITION") ok = floats(args, prims.back().position, 3);
        else if (cur == PRIM && c == "ROTATION") ok = floats(args, prims.back().rotation, 4);
        else if (cur == PRIM && c == "COLOR") ok = floats(args, prims.back().color, 3);
        else if (cur == PRIM && c == "EMISSION") ok = floats(args, prims.back().emission, 3);
        else if (cur == PRIM && c == "METALLIC") prims.back().material = RT_MAT_METALLIC;
        else if (cur == PRIM && c == "DIELECTRIC") prims.back().material = RT_MAT_DIELECTRIC;
        else if (cur == PRIM && c == "IOR") ok = floats(args, &prims.back().ior, 1);
        else if (cur == LIGHT && c == "LIGHT_INTENSITY") ok = floats(args, lights.back().intensity, 3);
        else if (cur == LIGHT && c == "LIGHT_DIRECTION") { lights.back().kind = RT_LIGHT_DIRECTIONAL; ok = floats(args, lights.back().vec, 3); }
        else if (cur == LIGHT && c == "LIGHT_POSITION") { lights.back().kind = RT_LIGHT_POINT; ok = floats(args, lights.back().vec, 3); }
        else if (cur == LIGHT && c == "LIGHT_ATTENUATION") ok = floats(args, lights.back().attenuation, 3);
        else ok = false;
        if (!ok) rc = RT_ERR_BAD_SCENE;
    }
    std::fclose(f);
    if (rc != RT_OK) return rc;
    if (prims.size() > RT_TEXT_MAX_PRIMS || lights.size() > RT_TEXT_MAX_LIGHTS || sc.ray_depth > RT_TEXT_MAX_DEPTH)
        return RT_ERR_BAD_SCENE;
    sc.shading = has_samples ? RT_SHADE_PATH : (has_whitted ? RT_SHADE_WHITTED : RT_SHADE_FLAT);
    sc.n_prims = (uint32_t)prims.size();
    sc.n_lights = (uint32_t)lights.size();
    const size_t bytes = sizeof(rt_text_scene) + prims.size() * sizeof(rt_text_prim) + lights.size() * sizeof(rt_text_light);
    char *block = static_cast<char *>(std::malloc(bytes));
    if (!block) return RT_ERR_OOM;
    rt_text_prim *pp = reinterpret_cast<rt_text_prim *>(block + sizeof(rt_text_scene));
    rt_text_light *lp = reinterpret_cast<rt_text_light *>(pp + prims.size());
    if (!prims.empty()) std::memcpy(pp, prims.data(), prims.size() * sizeof(rt_text_prim));
    if (!lights.empty()) std::memcpy(lp, lights.data(), lights.size() * sizeof(rt_text_light));
    sc.prims = pp;
    sc.lights = lp;
    std::memcpy(block, &sc, sizeof sc);
    *out = reinterpret_cast<rt_text_scene *>(block);
    return RT_OK;
}

void rt_text_scene_free(rt_text_scene *scene) { std::free(scene); }

void rt_host_tonemap_rgb8(const float *rgb_mean, size_t n_pixels, uint8_t *rgb8) {
    const float inv_gamma = 1 / 2.2f;  // image.h:49,63
    for (size_t i = 0; i < n_pixels * 3; ++i) {
        const float v = std::pow(aces(rgb_mean[i]), inv_gamma) * 255;
        // discretize_channel, image.h:66-69; NaN (never produced by sanitised means) maps to 0
        const float c = v != v ? 0.0f : std::clamp(v, 0.0f, 255.0f);
        rgb8[i] = static_cast<uint8_t>(std::round(c));
    }
}

int rt_host_write_ppm(const char *path, const uint8_t *rgb8, uint32_t width, uint32_t height) {
    FILE *f = std::fopen(path, "wb");
    if (!f) return RT_ERR_INVALID_ARG;
    std::fprintf(f, "P6\n%u %u\n255\n", width, height);  // Image::write, image.h:34-38
    const size_t n = (size_t)width * height * 3;
    const bool ok = std::fwrite(rgb8, 1, n, f) == n;
    return (std::fclose(f) == 0 && ok) ? RT_OK : RT_ERR_INVALID_ARG;
}

}  // extern "C"
