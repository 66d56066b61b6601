// rt_types.h — device-side scene layout (what rt_gpu_upload_scene builds from an rt_scene_desc).
//
// The reference keeps a pointer-linked AoS (Object* -> material -> Texture*, 208 B per triangle,
// 40 B nodes that hold their own box; src/bvh.h:157-168, src/geometry.h:604-659).  For the GPU
// everything is re-packed into 16-byte-aligned records fetched with 128-bit loads:
//
//   DNode  64 B  one per INNER node of the binary tree: both children's boxes (full precision) + child
//                links.  The reference tests both child boxes at the parent (bvh.h:205-212), so one
//                fetch replaces two 40 B node fetches.  Host-side intermediate (and host checks) only.
//   QNode  32 B  (host checks only; the device traversed it before QNode4) the same inner node with both child boxes quantised to 8 bits per plane on a per-node
//                power-of-two grid (conservative: the decoded box always contains the exact one): one
//                node visit is ONE 32-byte sector.  Divergent addresses cost the L1 data pipe one
//                sector per cycle, so bytes per node visit matter ("k_extend").
//   QNode4 64 B  two binary levels collapsed: the 2..4 grandchildren's boxes on one grid + 4 links;
//                what k_extend traverses.
//   DTri   64 B  triangle in BVH order: a, (b-a), (c-a) + scene.objects id + end-of-leaf flag, padded to
//                64 B so that it is two 256-bit loads (LDG.E.256 on sm_100a) inside one 128 B line.
//   DAttr  64 B  per-vertex normals + uv + material id, BVH order (read once per shade).
//   DMat   64 B  deduplicated material.
//
// A child link >= 0 is an inner-node index; < 0 is a leaf: ~link = index of its first DTri, the
// leaf ends at the first DTri whose `id_last` has bit 31 set.
#ifndef RT_TYPES_H
#define RT_TYPES_H

#include <stdint.h>

#define RT_STACK_SIZE 64        /* BVH::build max_depth = 64, bvh.h:371 */
#define RT_LINK_NONE 0x7FFFFFFF /* empty BVH */
#define RT_EXT_STACK_CAP 96     /* traversal-stack entries per ray in k_extend (shared + local part); the upload rejects a
                                   4-wide tree whose worst-case need (repack.h, stack_need4) exceeds it */
#define RT_LAST_BIT 0x80000000u

struct alignas(64) DNode {
    // lo/hi of the left (l) and right (r) child boxes
    float lminx, lminy, lminz, lmaxx;
    float lmaxy, lmaxz, rminx, rminy;
    float rminz, rmaxx, rmaxy, rmaxz;
    int32_t left, right;  // links
    int32_t pad0, pad1;
};

// Grid: plane(axis, q) = org[axis] + q * cell[axis], q in [0, 255].
//   org[axis]  = the float whose bits are the whole word org[axis] (so the low byte, which holds the
//                cell exponent, is part of the origin's mantissa; the packer rounds such that this
//                value is <= the exact minimum); bit 8 is always 0.
//   cell[axis] = 2^(e - 127), e = org[axis] & 0xFF (a float exponent field).
//   q[axis] bytes (low to high): left.min, left.max, right.min, right.max along that axis — one word per
//   axis, so that ONE byte permute with a per-ray selector (swap min/max where the ray runs in the negative
//   direction) turns it into (left.near, left.far, right.near, right.far) and the slab test needs no pairwise
//   min/max.
struct alignas(32) QNode {
    uint32_t org[3];
    uint32_t q[3];
    int32_t left, right;  // links
};

// 4-wide quantised node (collapse of two levels of the binary tree): same grid as QNode, per axis one word with the
// four children's min planes and one with their max planes.  Absent children (a wide node has 2..4) carry the
// link of the "null leaf" (one degenerate triangle at the end of the triangle array) and an inverted box.
struct alignas(64) QNode4 {
    uint32_t org[3];
    uint32_t lo[3];   // per axis: min-plane bytes of children 0..3
    uint32_t hi[3];   // per axis: max-plane bytes of children 0..3
    int32_t link[4];
    uint32_t pad[3];
};

// 8-wide quantised node (compressed wide BVH in the manner of Ylitie, Karras & Laine 2017, on this library's grid
// encoding): 96 B = three 32-byte sectors.
//   sector 0: grid origin words (as QNode), imask (bit s: the child in slot s is an inner node), index of the first
//             inner child (inner children are consecutive in the node array, in slot order), index of the first
//             triangle of the node's leaf children (consecutive in the triangle array, in slot order; the last
//             triangle of each leaf child carries RT_LAST_BIT), counts: 2 bits per slot = triangles of the leaf
//             child in that slot (1..3; 0 for an inner or empty slot)
//   sector 1: children in slots 0..3: per axis the min-plane bytes and the max-plane bytes (six words)
//   sector 2: children in slots 4..7
// Slots are assigned by octant (slot bit k set = the child lies on the positive side along axis k), so that
// `slot ^ octant-of-the-ray` orders the children front to back without any distance sort; empty slots carry an
// inverted box.  The traversal stack holds (first child / first triangle, hit mask) groups instead of one entry per
// child.  A BVH in this format has its triangles re-ordered (PackedBvh::order follows).
struct alignas(32) QNode8 {
    uint32_t org[3];
    uint32_t imask;
    uint32_t child_base;
    uint32_t tri_base;
    uint32_t counts;
    uint32_t pad;
    uint32_t g0[6], pad0[2];  // slots 0..3: lo x, lo y, lo z, hi x, hi y, hi z
    uint32_t g1[6], pad1[2];  // slots 4..7
};

struct alignas(64) DTri {
    float ax, ay, az;
    uint32_t id_last;  // scene.objects index | RT_LAST_BIT on the last triangle of a leaf
    float e1x, e1y, e1z, pad0;  // b - a  (triangle::v, geometry.h:473)
    float e2x, e2y, e2z, pad1;  // c - a  (triangle::u, geometry.h:475)
    float pad2[4];
};

struct alignas(64) DAttr {
    float n0x, n0y, n0z, uv0x;
    float n1x, n1y, n1z, uv0y;
    float n2x, n2y, n2z, uv1x;
    float uv1y, uv2x, uv2y;
    uint32_t material;
};

struct alignas(16) DTangent {  // only when some tangent differs from (1,0,0)
    float t0x, t0y, t0z, t1x;
    float t1y, t1z, t2x, t2y;
    float t2z, pad0, pad1, pad2;
};

struct alignas(64) DMat {
    float color[4];
    float emission[3];
    float roughness;
    float metallic, ior;
    int32_t color_tex, emissive_tex;
    int32_t mr_tex, normal_tex;
    int32_t pad0, pad1;
};

struct alignas(16) DTex {
    uint32_t offset;  // in texels (uint32 RGBA8) from the texel pool
    uint32_t width, height;
    uint32_t pad;
};

struct alignas(16) DLight {  // light triangle extras, light-BVH order
    float nx, ny, nz;  // triangle::normal (geometry.h:477)
    float area;        // triangle::square (geometry.h:481)
};

struct DBvh {
    const DNode *nodes;    // full-precision nodes (light BVH traversal, host checks); may be null on the device
    const QNode *qnodes;   // quantised 2-wide nodes
    const QNode4 *qnodes4; // quantised 4-wide nodes (collapsed tree) and their root link
    int32_t root4;
    // The same nodes as k_extend reads them (device only): bytes 0..31 of node i at q4lo + 32 i, bytes 32..63 at
    // q4lo + q4_hi_off + 32 i.  With 64-byte records the first 256-bit load of a node step only ever touches the L1
    // sector banks 0 and 2 and the second one banks 1 and 3; at a 32-byte stride each load spreads over all four,
    // and a divergent warp's load costs fewer data-pipe wavefronts (tools/l1_probe.cu: -23 % when the nodes hit the
    // L1, -10 % when they come from the L2).  The scene and the light BVH share one allocation and one q4_hi_off.
    const char *q4lo;
    uint32_t q4_hi_off;
    const QNode8 *qnodes8; // 8-wide nodes; node 0 is the root (n_nodes8 == 0: empty BVH)
    uint32_t n_nodes8;
    const DTri *tris;
    int32_t root;  // link; RT_LINK_NONE when empty
    uint32_t n_tris;
};

struct DScene {
    DBvh scene;
    DBvh light;
    const DAttr *attrs;         // scene-BVH order
    const DTangent *tangents;   // scene-BVH order or nullptr
    const DLight *light_extra;  // light-BVH order
    const DTri *light_sample;   // emissive triangles in the HOST light BVH's object order: bvh_mix_dist::sample
                                // draws a uniform index into that list (raytracer.h:355-361), so the sampling
                                // order stays the reference's even when the light BVH itself is rebuilt
    const DMat *materials;
    const DTex *textures;
    const uint32_t *texels;
    uint32_t n_lights;
    uint32_t ray_depth;
    float eps, min_roughness, vndf_factor;
    float bg[3];
    int32_t env_tex;  // -1: constant sky; else the equirectangular environment map (Scene::bg, scene.h:81)
    float cam_pos[3], cam_right[3], cam_up[3], cam_fwd[3];
    float fov_x;
};

#endif  // RT_TYPES_H
