/*
 * rt_host.h — host-side helpers around the rt_gpu C ABI (plain C++ inside, C ABI outside).
 *
 *  - rt_scene_save / rt_scene_load: "RTSC" binary container of an rt_scene_desc, the exchange
 *    format between a reference-hosted flattener, the oracle tools, the CLI and Python.
 *  - rt_host_build_bvh: own restatement of the reference's host BVH build
 *    (BVH::build / build_node / split_node, src/bvh.h:262-393) for hosts that do not link the
 *    reference headers; produces the same rt_bvh_node array the reference would.
 *  - rt_host_tonemap_rgb8 / rt_host_write_ppm: Image::set_pixel + Image::write
 *    (src/image.h:34-82) for hosts without the reference's Image class.
 *
 * None of this touches the GPU and none of it is an integrator: there is no CPU render path here.
 */
#ifndef RT_HOST_H
#define RT_HOST_H

#include "rt_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RTSC container ---------------------------------------------------------------------- */
int rt_scene_save(const rt_scene_desc *scene, const char *path);
/* Allocates one block holding the descriptor and all arrays; free with rt_scene_free. */
int rt_scene_load(const char *path, rt_scene_desc **out);
void rt_scene_free(rt_scene_desc *scene);
/* Structural validation used by rt_gpu_upload_scene as well (ids in range, node links sane). */
int rt_scene_validate(const rt_scene_desc *scene);

/* ---- BVH build (src/bvh.h:262-393) -------------------------------------------------------- */
typedef struct rt_bvh_build {
    uint32_t n_nodes;
    uint32_t root;
    uint32_t n_objects;
    uint32_t _pad;
    rt_bvh_node *nodes;   /* malloc'ed, free with rt_host_free_bvh */
    uint32_t *objects;
} rt_bvh_build;

/* tri_pos: n_tris*9 floats. select: optional n_tris bytes, non-zero = triangle takes part
 * (the `pred` of BVH::build, bvh.h:370); NULL = all. min_node_size=4, max_depth=64 are the
 * reference defaults (bvh.h:371). n_tris == 0 gives root = RT_NO_CHILD (bvh.h:373-376). */
int rt_host_build_bvh(const float *tri_pos, uint32_t n_tris, const uint8_t *select,
                      uint32_t min_node_size, uint32_t max_depth, rt_bvh_build *out);
void rt_host_free_bvh(rt_bvh_build *bvh);

/* ---- course text scenes (sample_data/scene-*.txt; PARITY UNPINNED, see rt_gpu.h) ---------------------------
 * Grammar (one command per line): DIMENSIONS, RAY_DEPTH, SAMPLES, BG_COLOR, AMBIENT_LIGHT,
 * CAMERA_POSITION/RIGHT/UP/FORWARD, CAMERA_FOV_X, NEW_PRIMITIVE {PLANE|ELLIPSOID|BOX|TRIANGLE, POSITION, ROTATION,
 * COLOR, METALLIC, DIELECTRIC, IOR, EMISSION}, NEW_LIGHT {LIGHT_INTENSITY, LIGHT_DIRECTION | LIGHT_POSITION,
 * LIGHT_ATTENUATION}.  Shading: SAMPLES present -> RT_SHADE_PATH; else any of NEW_LIGHT / RAY_DEPTH /
 * AMBIENT_LIGHT -> RT_SHADE_WHITTED; else RT_SHADE_FLAT.  Unknown commands are an error (RT_ERR_BAD_SCENE).
 * One malloc block; free with rt_text_scene_free. */
int rt_text_scene_parse(const char *path, rt_text_scene **out);
void rt_text_scene_free(rt_text_scene *scene);

/* ---- image (src/image.h:34-82) ------------------------------------------------------------ */
void rt_host_tonemap_rgb8(const float *rgb_mean, size_t n_pixels, uint8_t *rgb8);
int rt_host_write_ppm(const char *path, const uint8_t *rgb8, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif

#endif /* RT_HOST_H */
