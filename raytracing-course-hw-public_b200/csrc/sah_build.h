// sah_build.h — the library's own scene-BVH builder (host C++, multi-threaded binned SAH).
//
// Why the tree the host passes in is not traversed as is: the reference's builder (BVH::build, src/bvh.h:262-393)
// minimises a surface-area heuristic whose `aabb::surface_area` is 2*(2*dx*dy + dz*dz) (src/geometry.h:419-421,
// not an area), tries only the longest axis and weights the left side with the area of i+1 objects
// (src/bvh.h:303).  Closest hits do not depend on the tree, only the work per ray does: on the 260k-triangle
// bench scene a true-SAH tree over the same triangles needs 28.7 node steps + 3.3 triangle tests per extension
// ray instead of 33.5 + 9.1 (counted with the oracle's traversal counters).  SURVEY.md 8(f)-2 lists replacing
// BVH::build as the next row after the integrator; ids, not tree shape, are what parity compares.
//
// Algorithm: top-down, 32 centroid bins per axis, all three axes, cost = Ct*SA(node) + sum n_side*SA(side)
// with Ct = 1 triangle test, leaf when not splitting is cheaper and the node holds <= kMaxLeaf triangles;
// object-median split when binning cannot separate the centroids or when the depth budget (RT_STACK_SIZE)
// is nearly used up.  Sub-trees are independent: the top levels hand their left child to a new thread.
// Output is the reference's node format (rt_bvh_node, mirror of BVHNode) so that everything downstream
// (pack_bvh -> DNode / QNode) is shared with host-provided trees.
#ifndef RT_SAH_BUILD_H
#define RT_SAH_BUILD_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "rt_gpu.h"
#include "rt_types.h"

namespace rt {

struct BuiltBvh {
    std::vector<rt_bvh_node> nodes;  // sparse: slots a sub-tree did not need keep left_child == right_child == NO_CHILD, 0 objects
    std::vector<uint32_t> objects;
    uint32_t root = RT_NO_CHILD;
    rt_bvh_desc desc() const {
        rt_bvh_desc d;
        std::memset(&d, 0, sizeof d);
        d.n_nodes = static_cast<uint32_t>(nodes.size());
        d.root = objects.empty() ? RT_NO_CHILD : root;
        d.n_objects = static_cast<uint32_t>(objects.size());
        d.nodes = nodes.data();
        d.objects = objects.data();
        return d;
    }
};

namespace sah {

constexpr int kBins = 32;
constexpr uint32_t kMaxLeaf = 8;
// in triangle tests; 0.5 / 1 / 1.5 measured equal on the B200 (k_extend 86.9 ms), 2 -> 91.6, 3 -> 95.9
// (the 8-wide build, whose node step costs 1.5x the 4-wide one, does not want larger leaves either:
// 1 / 2 / 3 / 5 -> 94.4 / 100.2 / 113.7 / 115.8 ms of k_extend8 per 128 spp)
#ifndef RT_SAH_TRAVERSAL_COST
#define RT_SAH_TRAVERSAL_COST 1.0f
#endif
constexpr float kTraversalCost = RT_SAH_TRAVERSAL_COST;
constexpr uint32_t kParallelMin = 8192;  // sub-trees at least this large may get their own thread
constexpr int kParallelDepth = 5;        // ... down to this depth (<= 32 threads)

// fn(begin, end) over [0, n) on up to 16 threads
template <class F> inline void parallel_chunks(uint32_t n, F fn) {
    const unsigned hw = std::thread::hardware_concurrency();
    const uint32_t nt = std::max(1u, std::min(std::min(hw ? hw : 1u, 16u), n / 16384u));
    if (nt <= 1) {
        fn(0u, n);
        return;
    }
    std::vector<std::thread> th;
    const uint32_t chunk = (n + nt - 1) / nt;
    for (uint32_t t = 1; t < nt; ++t) th.emplace_back([=] { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk)); });
    fn(0u, std::min(n, chunk));
    for (auto &x : th) x.join();
}

struct Box3 {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::numeric_limits<float>::infinity();
            hi[k] = -std::numeric_limits<float>::infinity();
        }
    }
    void grow(const float *l, const float *h) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = l[k] < lo[k] ? l[k] : lo[k];
            hi[k] = h[k] > hi[k] ? h[k] : hi[k];
        }
    }
    void grow(const Box3 &b) { grow(b.lo, b.hi); }
    float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;  // half the surface area; only ratios matter
    }
};

struct Prim {  // 40 bytes, one per triangle
    float lo[3], hi[3], c[3];
    uint32_t id;
};

struct Ctx {
    std::vector<Prim> prims;  // permuted in place
    std::vector<Prim> tmp;    // scatter target of the parallel partition near the root
    std::vector<rt_bvh_node> nodes;
};

inline void make_leaf(Ctx &cx, uint32_t slot, uint32_t b, uint32_t e, const Box3 &box) {
    rt_bvh_node &nd = cx.nodes[slot];
    std::memcpy(nd.bmin, box.lo, sizeof nd.bmin);
    std::memcpy(nd.bmax, box.hi, sizeof nd.bmax);
    nd.left_child = nd.right_child = RT_NO_CHILD;
    nd.obj_begin = b;
    nd.obj_end = e;
}

// Builds the sub-tree of prims [b, e) into node slots [slot, slot + 2*(e-b) - 1).
inline void build_range(Ctx &cx, uint32_t slot, uint32_t b, uint32_t e, const Box3 &box, const Box3 &cbox, int depth) {
    const uint32_t n = e - b;
    Prim *P = cx.prims.data();
    if (n <= 1) {
        make_leaf(cx, slot, b, e, box);
        return;
    }
    // depth budget: an object-median tree over n primitives is ceil(log2 n) deep
    int log2n = 0;
    while ((1u << log2n) < n) ++log2n;
    const bool force_median = depth + log2n + 2 >= RT_STACK_SIZE - 2;

    uint32_t mid = b;
    Box3 lbox, rbox, lcb, rcb;
    bool have_split = false;
    if (!force_median) {
        float best_cost = std::numeric_limits<float>::infinity();
        int best_axis = -1, best_bin = -1;
        // one pass over the primitives fills the bins of all three axes; large ranges near the root are binned by
        // several threads into private bins that are merged afterwards
        struct Bins {
            Box3 box[3][kBins];
            uint32_t cnt[3][kBins];
            void reset(int nbins) {
                for (int a = 0; a < 3; ++a)
                    for (int i = 0; i < nbins; ++i) {
                        box[a][i].reset();
                        cnt[a][i] = 0;
                    }
            }
        };
        const int nb = n <= 32 ? 8 : kBins;  // small ranges: fewer bins, the sweep over empty bins would dominate
        float cmin[3], scale[3];
        bool use[3];
        for (int a = 0; a < 3; ++a) {
            const float cext = cbox.hi[a] - cbox.lo[a];
            use[a] = cext > 0.0f;
            cmin[a] = cbox.lo[a];
            scale[a] = use[a] ? static_cast<float>(nb) / cext : 0.0f;
        }
        auto bin_range = [&](uint32_t rb, uint32_t re, Bins &bins) {
            bins.reset(nb);
            for (uint32_t i = rb; i < re; ++i)
                for (int a = 0; a < 3; ++a) {
                    if (!use[a]) continue;
                    int k = static_cast<int>((P[i].c[a] - cmin[a]) * scale[a]);
                    k = k < 0 ? 0 : (k >= nb ? nb - 1 : k);
                    bins.box[a][k].grow(P[i].lo, P[i].hi);
                    ++bins.cnt[a][k];
                }
        };
        Bins local_bins;  // 2.7 KB on the stack (recursion depth <= RT_STACK_SIZE)
        Bins *bins = &local_bins;
        if (n >= 4 * kParallelMin && depth < kParallelDepth) {
            const unsigned hw = std::thread::hardware_concurrency();
            const uint32_t nt = std::max(1u, std::min(hw ? hw : 1u, std::min(16u >> depth, n / kParallelMin)));
            std::vector<std::unique_ptr<Bins>> part(nt);
            std::vector<std::thread> th;
            const uint32_t chunk = (n + nt - 1) / nt;
            for (uint32_t t = 0; t < nt; ++t) {
                part[t].reset(new Bins);
                const uint32_t rb = b + std::min(n, t * chunk), re = b + std::min(n, (t + 1) * chunk);
                if (t + 1 < nt) th.emplace_back([&bin_range, rb, re, &part, t] { bin_range(rb, re, *part[t]); });
                else bin_range(rb, re, *part[t]);
            }
            for (auto &x : th) x.join();
            bins->reset(nb);
            for (uint32_t t = 0; t < nt; ++t)
                for (int a = 0; a < 3; ++a)
                    for (int i = 0; i < nb; ++i)
                        if (part[t]->cnt[a][i]) {
                            bins->box[a][i].grow(part[t]->box[a][i]);
                            bins->cnt[a][i] += part[t]->cnt[a][i];
                        }
        } else {
            bin_range(b, e, *bins);
        }
        for (int axis = 0; axis < 3; ++axis) {
            if (!use[axis]) continue;
            const Box3 *bx = bins->box[axis];
            const uint32_t *cnt = bins->cnt[axis];
            float right_area[kBins];
            uint32_t right_cnt[kBins];
            Box3 acc;
            acc.reset();
            uint32_t c = 0;
            for (int i = nb - 1; i > 0; --i) {
                if (cnt[i]) acc.grow(bx[i]);
                c += cnt[i];
                right_area[i] = c ? acc.area() : 0.0f;
                right_cnt[i] = c;
            }
            acc.reset();
            c = 0;
            for (int i = 0; i < nb - 1; ++i) {  // split after bin i
                if (cnt[i]) acc.grow(bx[i]);
                c += cnt[i];
                if (c == 0 || right_cnt[i + 1] == 0) continue;
                const float cost = static_cast<float>(c) * acc.area() + static_cast<float>(right_cnt[i + 1]) * right_area[i + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = axis;
                    best_bin = i;
                }
            }
        }
        const float node_area = box.area();
        const float leaf_cost = static_cast<float>(n) * node_area;
        if (best_axis >= 0 && (n > kMaxLeaf || kTraversalCost * node_area + best_cost < leaf_cost)) {
            const int axis = best_axis;
            const float pmin = cmin[axis], pscale = scale[axis];
            lbox.reset(); rbox.reset(); lcb.reset(); rcb.reset();
            uint32_t i = b, j = e;
            uint32_t ntp = 1;
            if (n >= 4 * kParallelMin && depth < kParallelDepth) {  // (hardware_concurrency() is a system call: only here)
                const unsigned hw_p = std::thread::hardware_concurrency();
                ntp = std::max(1u, std::min(hw_p ? hw_p : 1u, std::min(16u >> depth, n / kParallelMin)));
            }
            if (ntp > 1) {
                // large ranges near the root: count + bound per chunk, scatter into cx.tmp at the chunks' offsets, copy back
                // (the serial in-place partition of the top four levels was the builder's longest serial stretch)
                struct Part {
                    uint32_t left = 0;
                    Box3 lb, rb, lc, rc;
                };
                std::vector<Part> part(ntp);
                const uint32_t chunk = (n + ntp - 1) / ntp;
                auto bin_of = [=](const Prim &pr) {
                    int k = static_cast<int>((pr.c[axis] - pmin) * pscale);
                    return k < 0 ? 0 : (k >= nb ? nb - 1 : k);
                };
                auto run = [&](auto fn) {
                    std::vector<std::thread> th;
                    for (uint32_t t = 1; t < ntp; ++t) th.emplace_back([&fn, t] { fn(t); });
                    fn(0u);
                    for (auto &x : th) x.join();
                };
                run([&](uint32_t t) {
                    Part pt;  // thread-local; stored once at the end (neighbouring Parts share cache lines)
                    pt.lb.reset(); pt.rb.reset(); pt.lc.reset(); pt.rc.reset();
                    const uint32_t cb = b + std::min(n, t * chunk), ce = b + std::min(n, (t + 1) * chunk);
                    for (uint32_t q = cb; q < ce; ++q) {
                        if (bin_of(P[q]) <= best_bin) {
                            ++pt.left;
                            pt.lb.grow(P[q].lo, P[q].hi);
                            pt.lc.grow(P[q].c, P[q].c);
                        } else {
                            pt.rb.grow(P[q].lo, P[q].hi);
                            pt.rc.grow(P[q].c, P[q].c);
                        }
                    }
                    part[t] = pt;
                });
                uint32_t total_left = 0;
                for (uint32_t t = 0; t < ntp; ++t) total_left += part[t].left;
                std::vector<uint32_t> loff(ntp), roff(ntp);
                uint32_t lo_acc = b, ro_acc = b + total_left;
                for (uint32_t t = 0; t < ntp; ++t) {
                    const uint32_t cb = b + std::min(n, t * chunk), ce = b + std::min(n, (t + 1) * chunk);
                    loff[t] = lo_acc;
                    roff[t] = ro_acc;
                    lo_acc += part[t].left;
                    ro_acc += (ce - cb) - part[t].left;
                    if (part[t].left) { lbox.grow(part[t].lb); lcb.grow(part[t].lc); }
                    if (part[t].left < ce - cb) { rbox.grow(part[t].rb); rcb.grow(part[t].rc); }
                }
                Prim *T = cx.tmp.data();
                run([&](uint32_t t) {
                    const uint32_t cb = b + std::min(n, t * chunk), ce = b + std::min(n, (t + 1) * chunk);
                    uint32_t l = loff[t], r = roff[t];
                    for (uint32_t q = cb; q < ce; ++q) {
                        if (bin_of(P[q]) <= best_bin) T[l++] = P[q];
                        else T[r++] = P[q];
                    }
                });
                run([&](uint32_t t) {
                    const uint32_t cb = b + std::min(n, t * chunk), ce = b + std::min(n, (t + 1) * chunk);
                    std::memcpy(P + cb, T + cb, static_cast<size_t>(ce - cb) * sizeof(Prim));
                });
                i = b + total_left;
                j = i;
            }
            while (i < j) {  // in-place partition by bin index
                int k = static_cast<int>((P[i].c[axis] - pmin) * pscale);
                k = k < 0 ? 0 : (k >= nb ? nb - 1 : k);
                if (k <= best_bin) {
                    lbox.grow(P[i].lo, P[i].hi);
                    lcb.grow(P[i].c, P[i].c);
                    ++i;
                } else {
                    --j;
                    std::swap(P[i], P[j]);
                    rbox.grow(P[j].lo, P[j].hi);
                    rcb.grow(P[j].c, P[j].c);
                }
            }
            mid = i;
            have_split = mid > b && mid < e;
        } else if (n <= kMaxLeaf) {
            make_leaf(cx, slot, b, e, box);
            return;
        }
    }
    if (!have_split) {  // object median along the widest centroid axis (or by index when all centroids coincide)
        int axis = 0;
        float ext = -1.0f;
        for (int k = 0; k < 3; ++k)
            if (cbox.hi[k] - cbox.lo[k] > ext) {
                ext = cbox.hi[k] - cbox.lo[k];
                axis = k;
            }
        if (!(ext > 0.0f) && n <= kMaxLeaf) {
            make_leaf(cx, slot, b, e, box);
            return;
        }
        mid = b + n / 2;
        if (ext > 0.0f)
            std::nth_element(P + b, P + mid, P + e, [axis](const Prim &l, const Prim &r) { return l.c[axis] < r.c[axis]; });
        lbox.reset(); rbox.reset(); lcb.reset(); rcb.reset();
        for (uint32_t i = b; i < mid; ++i) {
            lbox.grow(P[i].lo, P[i].hi);
            lcb.grow(P[i].c, P[i].c);
        }
        for (uint32_t i = mid; i < e; ++i) {
            rbox.grow(P[i].lo, P[i].hi);
            rcb.grow(P[i].c, P[i].c);
        }
    }
    rt_bvh_node &nd = cx.nodes[slot];
    std::memcpy(nd.bmin, box.lo, sizeof nd.bmin);
    std::memcpy(nd.bmax, box.hi, sizeof nd.bmax);
    nd.obj_begin = nd.obj_end = 0;
    const uint32_t lslot = slot + 1, rslot = slot + 2 * (mid - b);
    nd.left_child = lslot;
    nd.right_child = rslot;
    if (n >= kParallelMin && depth < kParallelDepth) {
        std::thread t([&cx, lslot, b, mid, lbox, lcb, depth]() { build_range(cx, lslot, b, mid, lbox, lcb, depth + 1); });
        build_range(cx, rslot, mid, e, rbox, rcb, depth + 1);
        t.join();
    } else {
        build_range(cx, lslot, b, mid, lbox, lcb, depth + 1);
        build_range(cx, rslot, mid, e, rbox, rcb, depth + 1);
    }
}

}  // namespace sah

// Builds a BVH over the triangles `ids` (indices into tri_pos, 9 floats per triangle).  `scratch`: working arrays a caller
// that builds repeatedly keeps alive, so that their memory is allocated and touched only once.
inline void build_sah_bvh(const float *tri_pos, const uint32_t *ids, uint32_t n, BuiltBvh &out, sah::Ctx *scratch = nullptr) {
    out.nodes.clear();
    out.objects.clear();
    out.root = RT_NO_CHILD;
    if (n == 0) return;
    sah::Ctx local;
    sah::Ctx &cx = scratch ? *scratch : local;
    cx.prims.resize(n);
    cx.tmp.resize(n);
    sah::Box3 box, cbox;
    box.reset();
    cbox.reset();
    {  // per-triangle boxes and centroids, in parallel; the scene box is merged from the chunks' boxes
        std::mutex mu;
        sah::parallel_chunks(n, [&](uint32_t b, uint32_t e) {
            sah::Box3 lb, lc;
            lb.reset();
            lc.reset();
            for (uint32_t i = b; i < e; ++i) {
                const uint32_t id = ids ? ids[i] : i;
                const float *p = tri_pos + static_cast<size_t>(id) * 9;
                sah::Prim &pr = cx.prims[i];
                for (int k = 0; k < 3; ++k) {
                    const float a = p[k], b2 = p[3 + k], c = p[6 + k];
                    pr.lo[k] = std::min(a, std::min(b2, c));
                    pr.hi[k] = std::max(a, std::max(b2, c));
                    pr.c[k] = 0.5f * (pr.lo[k] + pr.hi[k]);
                }
                pr.id = id;
                lb.grow(pr.lo, pr.hi);
                lc.grow(pr.c, pr.c);
            }
            std::lock_guard<std::mutex> g(mu);
            box.grow(lb);
            cbox.grow(lc);
        });
    }
    rt_bvh_node blank;
    std::memset(&blank, 0, sizeof blank);
    blank.left_child = blank.right_child = RT_NO_CHILD;
    cx.nodes.resize(static_cast<size_t>(2) * n - 1);
    {
        rt_bvh_node *nodes = cx.nodes.data();
        sah::parallel_chunks(2 * n - 1, [=](uint32_t b, uint32_t e) {
            for (uint32_t i = b; i < e; ++i) nodes[i] = blank;
        });
    }
    sah::build_range(cx, 0, 0, n, box, cbox, 0);
    out.nodes.swap(cx.nodes);
    out.objects.resize(n);
    {
        uint32_t *obj = out.objects.data();
        const sah::Prim *prims = cx.prims.data();
        sah::parallel_chunks(n, [=](uint32_t b, uint32_t e) {
            for (uint32_t i = b; i < e; ++i) obj[i] = prims[i].id;
        });
    }
    out.root = 0;
}

}  // namespace rt

#endif  // RT_SAH_BUILD_H
