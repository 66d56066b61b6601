// flatten_ref.hpp — reference structures -> rt_scene_desc (the host half of the drop-in).
//
// Include AFTER the reference's own headers (scene.h, bvh.h, raytracer.h), which are found by
// include path (-I<reference>/src) and are never copied into this repository.  A reference
// maintainer adds exactly this translation next to src/main.cpp:37; see INTEGRATION.md.
//
// What it walks (reference file:line):
//   Scene::objects / textures / camera / bg_color      src/scene.h:60-90
//   Object{shape, attrs, material}                      src/geometry.h:633-659
//   material (per-object copy -> deduplicated table)    src/geometry.h:604-614
//   Texture (float4 texels that are exactly u8/255)     src/geometry.h:529-599
//   RaytracerStaticContext::{scene_bvh, light_bvh}      src/raytracer.h:434-455
//   BVH{objects, nodes, root}, BVHNode                  src/bvh.h:157-168
#ifndef RT_FLATTEN_REF_HPP
#define RT_FLATTEN_REF_HPP

#include <cmath>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "rt_gpu.h"

namespace rt_flatten {

struct FlatScene {
    std::vector<float> tri_pos, tri_normals, tri_uv, tri_tangents;
    std::vector<uint32_t> tri_material;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    std::vector<uint8_t> texels;
    std::vector<uint32_t> scene_objects, light_objects;
    rt_scene_desc desc;
};

static_assert(sizeof(BVHNode) == sizeof(rt_bvh_node), "rt_bvh_node must mirror BVHNode (bvh.h:157-163)");

inline int texture_index(const Scene &scene, const geometry::Texture *t) {
    if (scene.textures.empty()) return -1;
    const geometry::Texture *b = scene.textures.data();
    if (t >= b && t < b + scene.textures.size()) return static_cast<int>(t - b);
    return -1;  // &WHITE_TEXTURE or &NORMAL_UP (geometry.h:601-602)
}

inline rt_bvh_desc flatten_bvh(const Scene &scene, const BVH &bvh, std::vector<uint32_t> &ids) {
    ids.resize(bvh.objects.size());
    for (size_t i = 0; i < bvh.objects.size(); ++i)
        ids[i] = static_cast<uint32_t>(bvh.objects[i] - scene.objects.data());
    rt_bvh_desc d;
    std::memset(&d, 0, sizeof d);
    d.n_nodes = static_cast<uint32_t>(bvh.nodes.size());
    d.root = bvh.root;
    d.n_objects = static_cast<uint32_t>(ids.size());
    d.nodes = reinterpret_cast<const rt_bvh_node *>(bvh.nodes.data());  // same layout, zero copy
    d.objects = ids.data();
    if (d.n_objects == 0) d.root = RT_NO_CHILD;  // an empty light BVH is a 0-object leaf (bvh.h:343-346)
    return d;
}

// `scene_bvh` (may be null: the library then builds the scene BVH itself) and `light_bvh` must outlive the returned
// FlatScene (BVH nodes are referenced, not copied).
inline void flatten(const Scene &scene, const BVH *scene_bvh, const BVH &light_bvh, FlatScene &out) {
    const size_t n = scene.objects.size();
    out.tri_pos.resize(n * 9);
    out.tri_normals.resize(n * 9);
    out.tri_uv.resize(n * 6);
    out.tri_tangents.resize(n * 9);
    out.tri_material.resize(n);
    out.materials.clear();

    using key_t = std::tuple<float, float, float, float, float, float, float, float, float, float, int, int, int, int>;
    std::map<key_t, uint32_t> seen;
    for (size_t i = 0; i < n; ++i) {
        const geometry::Object &o = scene.objects[i];
        for (int v = 0; v < 3; ++v) {
            for (int k = 0; k < 3; ++k) {
                out.tri_pos[i * 9 + v * 3 + k] = o.shape.vertices[v].val[k];
                out.tri_normals[i * 9 + v * 3 + k] = o.attrs.normals[v].val[k];
                out.tri_tangents[i * 9 + v * 3 + k] = o.attrs.tangents[v].val[k];
            }
            out.tri_uv[i * 6 + v * 2 + 0] = o.attrs.tex_coords[v].val[0];
            out.tri_uv[i * 6 + v * 2 + 1] = o.attrs.tex_coords[v].val[1];
        }
        const geometry::material &m = o.material;
        rt_material rm;
        for (int k = 0; k < 4; ++k) rm.color[k] = m.color.val[k];
        for (int k = 0; k < 3; ++k) rm.emission[k] = m.emission.val[k];
        rm.roughness = m.roughness;
        rm.metallic = m.metallic;
        rm.ior = m.ior;
        rm.color_tex = texture_index(scene, m.color_tex);
        rm.emissive_tex = texture_index(scene, m.emissive_tex);
        rm.metallic_roughness_tex = texture_index(scene, m.metallic_roughness_tex);
        rm.normal_tex = texture_index(scene, m.normal_tex);
        key_t key{rm.color[0], rm.color[1], rm.color[2], rm.color[3], rm.emission[0], rm.emission[1], rm.emission[2],
                  rm.roughness, rm.metallic, rm.ior, rm.color_tex, rm.emissive_tex, rm.metallic_roughness_tex,
                  rm.normal_tex};
        auto it = seen.find(key);
        if (it == seen.end()) {
            it = seen.emplace(key, static_cast<uint32_t>(out.materials.size())).first;
            out.materials.push_back(rm);
        }
        out.tri_material[i] = it->second;
    }

    out.textures.clear();
    out.texels.clear();
    for (const geometry::Texture &t : scene.textures) {
        rt_texture rt;
        rt.width = t.width;
        rt.height = t.height;
        rt.offset = out.texels.size();
        out.textures.push_back(rt);
        for (const geometry::color4 &c : t.data)
            for (int k = 0; k < 4; ++k)
                out.texels.push_back(static_cast<uint8_t>(std::lround(c.val[k] * 255.0f)));  // inverse of geometry.h:593
    }

    // Scene::bg (scene.h:81): the 1x1 WHITE_TEXTURE at HEAD (USE_ENV_MAP = false); a host that loads an environment
    // map into it (main.cpp:29-31) gets it appended as one more texture and referenced by env_texture
    uint32_t env_texture = 0;
    const bool white_1x1 = scene.bg.data.size() == 1 && scene.bg.data[0].val[0] == 1.0f && scene.bg.data[0].val[1] == 1.0f &&
                           scene.bg.data[0].val[2] == 1.0f;
    if (!scene.bg.data.empty() && !white_1x1) {
        rt_texture rt;
        rt.width = scene.bg.width;
        rt.height = scene.bg.height;
        rt.offset = out.texels.size();
        out.textures.push_back(rt);
        for (const geometry::color4 &c : scene.bg.data)
            for (int k = 0; k < 4; ++k) out.texels.push_back(static_cast<uint8_t>(std::lround(c.val[k] * 255.0f)));
        env_texture = static_cast<uint32_t>(out.textures.size());
    }

    rt_scene_desc &d = out.desc;
    std::memset(&d, 0, sizeof d);
    d.abi_version = RT_GPU_ABI_VERSION;
    d.n_tris = static_cast<uint32_t>(n);
    for (int k = 0; k < 3; ++k) {
        d.camera.position[k] = scene.camera.position.val[k];
        d.camera.right[k] = scene.camera.right.val[k];
        d.camera.up[k] = scene.camera.up.val[k];
        d.camera.forward[k] = scene.camera.forward.val[k];
        d.bg_color[k] = scene.bg_color.val[k];
    }
    d.camera.fov_x = scene.camera.fov_x;
    d.eps = EPS;
    d.min_roughness = MIN_ROUGHNESS;
    d.vndf_factor = VNDF_factor;
    d.ray_depth = scene.ray_depth;
    d.n_materials = static_cast<uint32_t>(out.materials.size());
    d.n_textures = static_cast<uint32_t>(out.textures.size());
    d.env_texture = env_texture;
    d.texel_bytes = out.texels.size();
    d.tri_pos = out.tri_pos.data();
    d.tri_normals = out.tri_normals.data();
    d.tri_uv = out.tri_uv.data();
    d.tri_tangents = out.tri_tangents.data();
    d.tri_material = out.tri_material.data();
    d.materials = out.materials.data();
    d.textures = out.textures.data();
    d.texels = out.texels.data();
    if (scene_bvh) {
        d.scene_bvh = flatten_bvh(scene, *scene_bvh, out.scene_objects);
    } else {
        std::memset(&d.scene_bvh, 0, sizeof d.scene_bvh);
        d.scene_bvh.root = RT_NO_CHILD;
    }
    d.light_bvh = flatten_bvh(scene, light_bvh, out.light_objects);
}

// The reference's own context (both trees built by RaytracerStaticContext, raytracer.h:440-447).
inline void flatten(const Scene &scene, const RaytracerStaticContext &ctx, FlatScene &out) {
    flatten(scene, &ctx.scene_bvh, ctx.light_bvh, out);
}

// Run-time form of the reference's compile-time ADD_LIGHT_TRIANGLE (config.h:39-47, false at HEAD): what
// parse_gltf_scene appends under that switch (scene.h:479-498) — one emissive triangle given in camera coordinates
// (LIGHT_TRIANGLE_RELATIVE_POS: 0.1 behind the camera plane), intensity LIGHT_TRIANGLE_INTENSITY, default material.
// A reference built with ADD_LIGHT_TRIANGLE = true needs nothing from this: its loader has appended the object already.
inline void append_light_triangle(Scene &scene) {
    geometry::Object light;
    const geometry::vec3 axes[3] = {scene.camera.right, scene.camera.up, scene.camera.forward};
    geometry::vec3 *corner[3] = {&light.shape.a(), &light.shape.b(), &light.shape.c()};
    for (int v = 0; v < 3; ++v) {
        const auto &rel = LIGHT_TRIANGLE_RELATIVE_POS[v];
        *corner[v] = scene.camera.position + geometry::transform3(geometry::vec3(rel[0], rel[1], rel[2]), axes[0], axes[1], axes[2]);
    }
    light.material.emission = {LIGHT_TRIANGLE_INTENSITY, LIGHT_TRIANGLE_INTENSITY, LIGHT_TRIANGLE_INTENSITY};
    light.attrs.normals.fill(light.shape.normal());
    light.attrs.tex_coords.fill(geometry::vec2(0, 0));
    light.attrs.tangents.fill(geometry::vec3(1, 0, 0));
    scene.objects.push_back(light);
}

// raytracer.h:444-447 alone: the light BVH, whose object order the light sampler indexes.
inline BVH build_light_bvh(const Scene &scene) {
    return BVH::build(std::span(scene.objects),
                      [](const geometry::Object &obj) { return obj.material.emission != geometry::color3{0, 0, 0}; });
}

}  // namespace rt_flatten

#endif  // RT_FLATTEN_REF_HPP
