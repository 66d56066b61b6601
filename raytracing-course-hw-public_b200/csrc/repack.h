// repack.h — host-side re-packing of an rt_scene_desc (reference layout) into the device layout of
// rt_types.h.  Pure C++ (no CUDA), header-only; used by rt_gpu_upload_scene and by the host unit
// test of the device math (tests/hostcheck).
#ifndef RT_REPACK_H
#define RT_REPACK_H

#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "rt_gpu.h"
#include "rt_types.h"

namespace rt {

struct PackedBvh {
    std::vector<DNode> nodes;
    std::vector<DTri> tris;       // BVH object order
    std::vector<uint32_t> order;  // BVH position -> scene.objects index
    int32_t root = RT_LINK_NONE;
    uint32_t max_depth = 0;
};

struct PackedScene {
    PackedBvh scene, light;
    std::vector<DAttr> attrs;
    std::vector<DTangent> tangents;  // empty when every tangent is (1,0,0)
    std::vector<DLight> light_extra;
    std::vector<DMat> materials;
    std::vector<DTex> textures;
    std::vector<uint32_t> texels;
    float gamma_lut[256];  // powf(k/255, 2.2f): the per-texel pow of Texture::sample (geometry.h:525-527,561)
};

namespace detail {

inline void set_box(DNode &n, bool left, const float *lo, const float *hi) {
    if (left) {
        n.lminx = lo[0]; n.lminy = lo[1]; n.lminz = lo[2];
        n.lmaxx = hi[0]; n.lmaxy = hi[1]; n.lmaxz = hi[2];
    } else {
        n.rminx = lo[0]; n.rminy = lo[1]; n.rminz = lo[2];
        n.rmaxx = hi[0]; n.rmaxy = hi[1]; n.rmaxz = hi[2];
    }
}

// Returns the link of reference node `ref_id` (inner index or ~first_tri), or RT_LINK_NONE for an
// empty leaf.  `rc` collects structural errors.
inline int32_t pack_node(const rt_bvh_desc &src, uint32_t ref_id, PackedBvh &out, uint32_t depth, int &rc) {
    const rt_bvh_node &nd = src.nodes[ref_id];
    if (depth > out.max_depth) out.max_depth = depth;
    const bool has_children = nd.left_child != RT_NO_CHILD || nd.right_child != RT_NO_CHILD;
    if (!has_children) {
        if (nd.obj_begin >= nd.obj_end) return RT_LINK_NONE;
        out.tris[nd.obj_end - 1].id_last |= RT_LAST_BIT;
        return ~static_cast<int32_t>(nd.obj_begin);
    }
    if (nd.obj_begin < nd.obj_end || depth >= RT_STACK_SIZE) {
        rc = RT_ERR_BAD_SCENE;  // inner node with own objects / deeper than bvh.h:371 allows
        return RT_LINK_NONE;
    }
    const int32_t idx = static_cast<int32_t>(out.nodes.size());
    out.nodes.emplace_back();
    const float inf = std::numeric_limits<float>::infinity();
    const float empty_lo[3] = {inf, inf, inf}, empty_hi[3] = {-inf, -inf, -inf};
    int32_t links[2];
    const uint32_t child[2] = {nd.left_child, nd.right_child};
    for (int c = 0; c < 2; ++c) {
        if (child[c] == RT_NO_CHILD) {
            links[c] = RT_LINK_NONE;
            set_box(out.nodes[idx], c == 0, empty_lo, empty_hi);
        } else {
            links[c] = pack_node(src, child[c], out, depth + 1, rc);
            const rt_bvh_node &ch = src.nodes[child[c]];
            if (links[c] == RT_LINK_NONE)
                set_box(out.nodes[idx], c == 0, empty_lo, empty_hi);
            else
                set_box(out.nodes[idx], c == 0, ch.bmin, ch.bmax);
        }
    }
    out.nodes[idx].left = links[0];
    out.nodes[idx].right = links[1];
    out.nodes[idx].pad0 = out.nodes[idx].pad1 = 0;
    return idx;
}

}  // namespace detail

inline int pack_bvh(const rt_scene_desc &sc, const rt_bvh_desc &src, PackedBvh &out) {
    out.nodes.clear();
    out.tris.assign(src.n_objects, DTri());
    out.order.assign(src.objects, src.objects + src.n_objects);
    out.root = RT_LINK_NONE;
    out.max_depth = 0;
    for (uint32_t k = 0; k < src.n_objects; ++k) {
        const uint32_t id = src.objects[k];
        const float *p = sc.tri_pos + static_cast<size_t>(id) * 9;
        DTri &t = out.tris[k];
        t.ax = p[0]; t.ay = p[1]; t.az = p[2];
        t.e1x = p[3] - p[0]; t.e1y = p[4] - p[1]; t.e1z = p[5] - p[2];
        t.e2x = p[6] - p[0]; t.e2y = p[7] - p[1]; t.e2z = p[8] - p[2];
        t.pad0 = t.pad1 = 0.0f;
        t.pad2[0] = t.pad2[1] = t.pad2[2] = t.pad2[3] = 0.0f;
        t.id_last = id;
    }
    if (src.root == RT_NO_CHILD || src.n_objects == 0) return RT_OK;
    int rc = RT_OK;
    out.root = detail::pack_node(src, src.root, out, 0, rc);
    return rc;
}

inline int pack_scene(const rt_scene_desc &sc, PackedScene &out) {
    if (int rc = pack_bvh(sc, sc.scene_bvh, out.scene)) return rc;
    if (int rc = pack_bvh(sc, sc.light_bvh, out.light)) return rc;

    const uint32_t n = sc.scene_bvh.n_objects;
    out.attrs.resize(n);
    bool any_tangent = false;
    if (sc.tri_tangents)
        for (size_t i = 0; i < static_cast<size_t>(sc.n_tris) * 3 && !any_tangent; ++i) {
            const float *t = sc.tri_tangents + i * 3;
            any_tangent = !(t[0] == 1.0f && t[1] == 0.0f && t[2] == 0.0f);
        }
    out.tangents.clear();
    if (any_tangent) out.tangents.resize(n);
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t id = sc.scene_bvh.objects[k];
        const float *nn = sc.tri_normals + static_cast<size_t>(id) * 9;
        const float *uv = sc.tri_uv + static_cast<size_t>(id) * 6;
        DAttr &a = out.attrs[k];
        a.n0x = nn[0]; a.n0y = nn[1]; a.n0z = nn[2];
        a.n1x = nn[3]; a.n1y = nn[4]; a.n1z = nn[5];
        a.n2x = nn[6]; a.n2y = nn[7]; a.n2z = nn[8];
        a.uv0x = uv[0]; a.uv0y = uv[1]; a.uv1x = uv[2]; a.uv1y = uv[3]; a.uv2x = uv[4]; a.uv2y = uv[5];
        a.material = sc.tri_material[id];
        if (any_tangent) {
            const float *t = sc.tri_tangents + static_cast<size_t>(id) * 9;
            DTangent &d = out.tangents[k];
            d.t0x = t[0]; d.t0y = t[1]; d.t0z = t[2];
            d.t1x = t[3]; d.t1y = t[4]; d.t1z = t[5];
            d.t2x = t[6]; d.t2y = t[7]; d.t2z = t[8];
            d.pad0 = d.pad1 = d.pad2 = 0.0f;
        }
    }

    out.light_extra.resize(sc.light_bvh.n_objects);
    for (uint32_t k = 0; k < sc.light_bvh.n_objects; ++k) {
        const DTri &t = out.light.tris[k];
        // crs(v, u), triangle::normal / square (geometry.h:477-483)
        const float cx = t.e1y * t.e2z - t.e1z * t.e2y;
        const float cy = t.e1z * t.e2x - t.e1x * t.e2z;
        const float cz = t.e1x * t.e2y - t.e1y * t.e2x;
        const float l = std::sqrt(cx * cx + cy * cy + cz * cz);
        out.light_extra[k] = {cx / l, cy / l, cz / l, l / 2};
    }

    out.materials.resize(sc.n_materials);
    for (uint32_t i = 0; i < sc.n_materials; ++i) {
        const rt_material &m = sc.materials[i];
        DMat &d = out.materials[i];
        std::memcpy(d.color, m.color, sizeof d.color);
        std::memcpy(d.emission, m.emission, sizeof d.emission);
        d.roughness = m.roughness;
        d.metallic = m.metallic;
        d.ior = m.ior;
        d.color_tex = m.color_tex;
        d.emissive_tex = m.emissive_tex;
        d.mr_tex = m.metallic_roughness_tex;
        d.normal_tex = m.normal_tex;
        d.pad0 = d.pad1 = 0;
    }

    out.textures.resize(sc.n_textures);
    out.texels.resize(sc.texel_bytes / 4);
    if (sc.texel_bytes) std::memcpy(out.texels.data(), sc.texels, out.texels.size() * 4);
    for (uint32_t i = 0; i < sc.n_textures; ++i) {
        if (sc.textures[i].offset % 4) return RT_ERR_BAD_SCENE;
        out.textures[i] = {static_cast<uint32_t>(sc.textures[i].offset / 4), sc.textures[i].width,
                           sc.textures[i].height, 0};
    }
    for (int k = 0; k < 256; ++k) out.gamma_lut[k] = std::pow(static_cast<float>(k) / 255.0f, 2.2f);
    return RT_OK;
}

// Fill the pointer-free part of a DScene; pointers are set by the caller (device or host arrays).
inline void fill_scene_constants(const rt_scene_desc &sc, const PackedScene &p, DScene &d) {
    std::memset(&d, 0, sizeof d);
    d.scene.root = p.scene.root;
    d.scene.n_tris = static_cast<uint32_t>(p.scene.tris.size());
    d.light.root = p.light.root;
    d.light.n_tris = static_cast<uint32_t>(p.light.tris.size());
    d.n_lights = sc.light_bvh.n_objects;
    d.ray_depth = sc.ray_depth;
    d.eps = sc.eps;
    d.min_roughness = sc.min_roughness;
    d.vndf_factor = sc.vndf_factor;
    for (int k = 0; k < 3; ++k) {
        d.bg[k] = sc.bg_color[k];
        d.cam_pos[k] = sc.camera.position[k];
        d.cam_right[k] = sc.camera.right[k];
        d.cam_up[k] = sc.camera.up[k];
        d.cam_fwd[k] = sc.camera.forward[k];
    }
    d.fov_x = sc.camera.fov_x;
}

}  // namespace rt

#endif  // RT_REPACK_H
