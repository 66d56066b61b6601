// wide_study.cpp — host-side experiment (not part of the product, not part of the tests): how many node visits
// do different wide-BVH traversal schemes need on the bench scene's path-traced ray distribution?
//   A  4-wide, children sorted by entry distance, (link, t) stack with distance culling at pop  (what k_extend does)
//   B  N-wide (4 or 8), octant-ordered slots, (node, hit-mask) stack without distances (Ylitie et al. 2017 style):
//      a stale entry is only culled when its own children are tested against the shrunken best_t
// Built by tools/wide_study.py into build/; prints counts per ray.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "pt_core.cuh"
#include "repack.h"
#include "rt_gpu.h"

using namespace rt;

namespace {
struct WChild {
    float lo[3], hi[3];
    int32_t link;  // >= 0 wide node, < 0 leaf (~first tri), RT_LINK_NONE empty
};
struct WNode {
    WChild ch[8];
    int n;
    int ax0, ax1;  // 4-wide: the two axes the slot bits refer to
};
struct Item {
    int32_t link;
    float lo[3], hi[3];
};
float area(const Item &c) {
    const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
    return dx * dy + dy * dz + dz * dx;
}
void kids(const DNode &n, Item out[2]) {
    out[0] = Item{n.left, {n.lminx, n.lminy, n.lminz}, {n.lmaxx, n.lmaxy, n.lmaxz}};
    out[1] = Item{n.right, {n.rminx, n.rminy, n.rminz}, {n.rmaxx, n.rmaxy, n.rmaxz}};
}
int32_t collapse(const std::vector<DNode> &nodes, int32_t link, int width, std::vector<WNode> &out) {
    if (link < 0) return link;
    Item it[8];
    int n = 2;
    kids(nodes[link], it);
    while (n < width) {
        int best = -1;
        float ba = -1;
        for (int i = 0; i < n; ++i)
            if (it[i].link >= 0 && area(it[i]) > ba) ba = area(it[i]), best = i;
        if (best < 0) break;
        Item two[2];
        kids(nodes[it[best].link], two);
        it[best] = two[0];
        it[n++] = two[1];
    }
    // slot assignment: greedy on dot(child centre - node centre, slot direction); slot s bit k set = positive side
    float c[3] = {0, 0, 0}, lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = it[0].lo[k];
        hi[k] = it[0].hi[k];
        for (int i = 1; i < n; ++i) lo[k] = std::min(lo[k], it[i].lo[k]), hi[k] = std::max(hi[k], it[i].hi[k]);
        c[k] = 0.5f * (lo[k] + hi[k]);
    }
    const int slots = width;
    // 4 slots: the two axes along which the child centres spread most
    int ax0 = 0, ax1 = 1;
    {
        float spread[3];
        for (int k = 0; k < 3; ++k) {
            float mn = 1e30f, mx = -1e30f;
            for (int i = 0; i < n; ++i) {
                const float cc = 0.5f * (it[i].lo[k] + it[i].hi[k]);
                mn = std::min(mn, cc), mx = std::max(mx, cc);
            }
            spread[k] = mx - mn;
        }
        int order[3] = {0, 1, 2};
        std::sort(order, order + 3, [&](int a, int b) { return spread[a] > spread[b]; });
        ax0 = order[0];
        ax1 = order[1];
    }
    int slot_of[8];
    bool used[8] = {false}, done[8] = {false};
    for (int round = 0; round < n; ++round) {
        float bestv = -1e30f;
        int bi = -1, bs = -1;
        for (int i = 0; i < n; ++i) {
            if (done[i]) continue;
            for (int s = 0; s < slots; ++s) {
                if (used[s]) continue;
                float v = 0;
                for (int k = 0; k < 3; ++k) {
                    const float cc = 0.5f * (it[i].lo[k] + it[i].hi[k]) - c[k];
                    int bit;
                    if (slots == 8) bit = (s >> k) & 1;
                    else if (k == ax0) bit = s & 1;
                    else if (k == ax1) bit = (s >> 1) & 1;
                    else continue;
                    v += bit ? cc : -cc;
                }
                if (v > bestv) bestv = v, bi = i, bs = s;
            }
        }
        done[bi] = true;
        used[bs] = true;
        slot_of[bi] = bs;
    }
    const int32_t idx = (int32_t)out.size();
    out.emplace_back();
    WNode w;
    w.n = slots;
    w.ax0 = ax0;
    w.ax1 = ax1;
    for (int s = 0; s < 8; ++s) w.ch[s].link = RT_LINK_NONE;
    for (int i = 0; i < n; ++i) {
        WChild &cdst = w.ch[slot_of[i]];
        std::memcpy(cdst.lo, it[i].lo, 12);
        std::memcpy(cdst.hi, it[i].hi, 12);
        cdst.link = it[i].link;
    }
    out[idx] = w;
    for (int s = 0; s < slots; ++s)
        if (out[idx].ch[s].link != RT_LINK_NONE && out[idx].ch[s].link >= 0) {
            const int32_t l = collapse(nodes, out[idx].ch[s].link, width, out);
            out[idx].ch[s].link = l;
        }
    return idx;
}


// ---- cost-optimal 8-wide collapse (Ylitie et al. 2017, sec. 4.1, without leaf merging): minimises the summed box
// area of the wide nodes.  C[n][i-1] = cost of representing the binary subtree n by at most i children of its parent.
struct DpCollapse {
    const std::vector<DNode> &nodes;
    std::vector<float> C;       // 8 per inner node
    std::vector<uint8_t> split; // [n][i-1]: slots given to the left subtree when n is distributed over i slots (0: take C[n][i-2])
    std::vector<float> area_of; // box area of each inner node (from its parent's record); root: 0
    int W;
    DpCollapse(const std::vector<DNode> &nd, int width = 8) : nodes(nd), C(nd.size() * 8, 0.0f), split(nd.size() * 8, 0), area_of(nd.size(), 0.0f), W(width) {}
    float cost(int32_t link, int i) const { return link < 0 ? 0.0f : C[(size_t)link * 8 + (i - 1)]; }
    void solve(int32_t n) {
        if (n < 0) return;
        Item k[2];
        kids(nodes[n], k);
        for (int c = 0; c < 2; ++c)
            if (k[c].link >= 0) {
                area_of[k[c].link] = area(k[c]);
                solve(k[c].link);
            }
        // dist[j] = best split of j slots (j = 2..8) over the two subtrees
        float dist[9];
        uint8_t arg[9];
        for (int j = 2; j <= W; ++j) {
            dist[j] = 1e30f;
            arg[j] = 1;
            for (int a = 1; a < j; ++a) {
                const float v = cost(k[0].link, std::min(a, W - 1)) + cost(k[1].link, std::min(j - a, W - 1));
                if (v < dist[j]) dist[j] = v, arg[j] = (uint8_t)a;
            }
        }
        float *c = &C[(size_t)n * 8];
        uint8_t *sp = &split[(size_t)n * 8];
        c[0] = area_of[n] + dist[W];  // n is a wide node of its own
        sp[0] = arg[W];
        for (int i = 2; i <= W; ++i) {
            if (dist[i] < c[i - 2]) c[i - 1] = dist[i], sp[i - 1] = arg[i];
            else c[i - 1] = c[i - 2], sp[i - 1] = 0;
        }
    }
    // children of the wide node that n's subtree contributes when it may use i slots
    void expand(int32_t link, const Item &self, int i, std::vector<Item> &out) const {
        if (link < 0) {
            out.push_back(self);
            return;
        }
        while (i > 1 && split[(size_t)link * 8 + (i - 1)] == 0) --i;
        if (i == 1) {
            out.push_back(self);  // stays an inner child: its own wide node
            return;
        }
        const int a = split[(size_t)link * 8 + (i - 1)];
        Item k[2];
        kids(nodes[link], k);
        expand(k[0].link, k[0], std::min(a, W - 1), out);
        expand(k[1].link, k[1], std::min(i - a, W - 1), out);
    }
};
int32_t collapse_dp(const DpCollapse &dp, int32_t link, std::vector<WNode> &out);

int32_t collapse_dp(const DpCollapse &dp, int32_t link, std::vector<WNode> &out) {
    if (link < 0) return link;
    const std::vector<DNode> &nodes = dp.nodes;
    const int width = dp.W;
    std::vector<Item> items;
    Item k[2];
    kids(nodes[link], k);
    const int a = dp.split[(size_t)link * 8 + 0];
    dp.expand(k[0].link, k[0], std::min(a, width - 1), items);
    dp.expand(k[1].link, k[1], std::min(width - a, width - 1), items);
    Item it[8];
    const int n = (int)items.size();
    for (int i = 0; i < n; ++i) it[i] = items[i];
    // slot assignment: greedy on dot(child centre - node centre, slot direction); slot s bit k set = positive side
    float c[3] = {0, 0, 0}, lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = it[0].lo[k];
        hi[k] = it[0].hi[k];
        for (int i = 1; i < n; ++i) lo[k] = std::min(lo[k], it[i].lo[k]), hi[k] = std::max(hi[k], it[i].hi[k]);
        c[k] = 0.5f * (lo[k] + hi[k]);
    }
    const int slots = width;
    // 4 slots: the two axes along which the child centres spread most
    int ax0 = 0, ax1 = 1;
    {
        float spread[3];
        for (int k = 0; k < 3; ++k) {
            float mn = 1e30f, mx = -1e30f;
            for (int i = 0; i < n; ++i) {
                const float cc = 0.5f * (it[i].lo[k] + it[i].hi[k]);
                mn = std::min(mn, cc), mx = std::max(mx, cc);
            }
            spread[k] = mx - mn;
        }
        int order[3] = {0, 1, 2};
        std::sort(order, order + 3, [&](int a, int b) { return spread[a] > spread[b]; });
        ax0 = order[0];
        ax1 = order[1];
    }
    int slot_of[8];
    bool used[8] = {false}, done[8] = {false};
    for (int round = 0; round < n; ++round) {
        float bestv = -1e30f;
        int bi = -1, bs = -1;
        for (int i = 0; i < n; ++i) {
            if (done[i]) continue;
            for (int s = 0; s < slots; ++s) {
                if (used[s]) continue;
                float v = 0;
                for (int k = 0; k < 3; ++k) {
                    const float cc = 0.5f * (it[i].lo[k] + it[i].hi[k]) - c[k];
                    int bit;
                    if (slots == 8) bit = (s >> k) & 1;
                    else if (k == ax0) bit = s & 1;
                    else if (k == ax1) bit = (s >> 1) & 1;
                    else continue;
                    v += bit ? cc : -cc;
                }
                if (v > bestv) bestv = v, bi = i, bs = s;
            }
        }
        done[bi] = true;
        used[bs] = true;
        slot_of[bi] = bs;
    }
    const int32_t idx = (int32_t)out.size();
    out.emplace_back();
    WNode w;
    w.n = slots;
    w.ax0 = ax0;
    w.ax1 = ax1;
    for (int s = 0; s < 8; ++s) w.ch[s].link = RT_LINK_NONE;
    for (int i = 0; i < n; ++i) {
        WChild &cdst = w.ch[slot_of[i]];
        std::memcpy(cdst.lo, it[i].lo, 12);
        std::memcpy(cdst.hi, it[i].hi, 12);
        cdst.link = it[i].link;
    }
    out[idx] = w;
    for (int s = 0; s < slots; ++s)
        if (out[idx].ch[s].link != RT_LINK_NONE && out[idx].ch[s].link >= 0) {
            const int32_t l = collapse_dp(dp, out[idx].ch[s].link, out);
            out[idx].ch[s].link = l;
        }
    return idx;
}

struct Cnt {
    uint64_t rays = 0, nodes = 0, tris = 0, pops = 0, culled = 0, leaves = 0, pushes = 0;
    std::vector<uint32_t> per_node;  // optional: visits per wide node
};

bool tri_hit(const DBvh &bvh, uint32_t k, f3 o, f3 d, float eps, Hit &best) {
    const char *p = reinterpret_cast<const char *>(bvh.tris + k);
    const f4 t0 = ld4(p), t1 = ld4(p + 16), t2 = ld4(p + 32);
    float t, b, c;
    if (tri_test(mk3(t0.x, t0.y, t0.z), mk3(t1.x, t1.y, t1.z), mk3(t2.x, t2.y, t2.z), o, d, eps, t, b, c) && t < best.t) {
        best.t = t;
        best.b = b;
        best.c = c;
        best.tri = (int32_t)k;
    }
    return (f2u(t0.w) & RT_LAST_BIT) != 0;
}

// scheme B
Hit trav_oct(const std::vector<WNode> &W, int32_t root, int width, const DBvh &bvh, f3 o, f3 d, float eps, Cnt &c, bool by_dist) {
    Hit best;
    best.t = INFINITY;
    best.b = best.c = 0;
    best.tri = -1;
    ++c.rays;
    const f3 idir = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const float dd[3] = {d.x, d.y, d.z};
    auto inv_of = [&](const WNode &n) {
        if (width == 8) return (d.x < 0 ? 0 : 1) | (d.y < 0 ? 0 : 2) | (d.z < 0 ? 0 : 4);
        return (dd[n.ax0] < 0 ? 0 : 1) | (dd[n.ax1] < 0 ? 0 : 2);
    };
    struct E {
        int32_t node;
        uint32_t mask;
    };
    E st[256];
    int sp = 0;
    if (root < 0) {
        for (uint32_t k = (uint32_t)~root;; ++k) {
            ++c.tris;
            if (tri_hit(bvh, k, o, d, eps, best)) break;
        }
        return best;
    }
    int32_t cur = root;
    for (;;) {
        ++c.nodes;
        const WNode &n = W[cur];
        uint32_t mask = 0;
        float dist[8];
        for (int s = 0; s < width; ++s) {
            const WChild &ch = n.ch[s];
            if (ch.link == RT_LINK_NONE) continue;
            const float t = slab(ch.lo[0], ch.lo[1], ch.lo[2], ch.hi[0], ch.hi[1], ch.hi[2], o, idir, eps);
            if (t >= 0.0f && t < best.t) {
                mask |= 1u << s;
                dist[s] = t;
            }
        }
        // leaves first (Ylitie: triangles of the node are intersected before descending), in priority order
        E e{cur, mask};
        for (;;) {
            // next child by priority
            int pick = -1;
            if (!by_dist) {
                int bestp = -1;
                for (int s = 0; s < width; ++s)
                    if (e.mask >> s & 1) {
                        const int pr = (s ^ inv_of(W[e.node])) & (width - 1);
                        if (pr > bestp) bestp = pr, pick = s;
                    }
            } else {
                float bd = INFINITY;
                for (int s = 0; s < width; ++s)
                    if ((e.mask >> s & 1) && dist[s] < bd) bd = dist[s], pick = s;
            }
            if (pick < 0) break;
            e.mask &= ~(1u << pick);
            const int32_t l = W[e.node].ch[pick].link;
            if (l < 0) {
                ++c.leaves;
                for (uint32_t k = (uint32_t)~l;; ++k) {
                    ++c.tris;
                    if (tri_hit(bvh, k, o, d, eps, best)) break;
                }
                continue;
            }
            if (e.mask) {
                st[sp++] = e;
                ++c.pushes;
            }
            cur = l;
            goto next_node;
        }
        // pop
        for (;;) {
            if (sp == 0) return best;
            E &t = st[sp - 1];
            ++c.pops;
            int pick = -1, bestp = -1;
            for (int s = 0; s < width; ++s)
                if (t.mask >> s & 1) {
                    const int pr = (s ^ inv_of(W[t.node])) & (width - 1);
                    if (pr > bestp) bestp = pr, pick = s;
                }
            t.mask &= ~(1u << pick);
            const int32_t node = t.node;
            if (!t.mask) --sp;
            const int32_t l = W[node].ch[pick].link;
            if (l < 0) {
                ++c.leaves;
                for (uint32_t k = (uint32_t)~l;; ++k) {
                    ++c.tris;
                    if (tri_hit(bvh, k, o, d, eps, best)) break;
                }
                continue;
            }
            cur = l;
            break;
        }
    next_node:;
    }
}

// scheme A on the same wide nodes: sorted by distance, (link, t) stack, cull at pop
Hit trav_sorted(const std::vector<WNode> &W, int32_t root, int width, const DBvh &bvh, f3 o, f3 d, float eps, Cnt &c, int mode = 0) {
    Hit best;
    best.t = INFINITY;
    best.b = best.c = 0;
    best.tri = -1;
    ++c.rays;
    const f3 idir = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int32_t sl[512];
    float stt[512];
    int sp = 0;
    int32_t link = root;
    for (;;) {
        if (link >= 0) {
            ++c.nodes;
            if (!c.per_node.empty()) ++c.per_node[link];
            const WNode &n = W[link];
            float dist[8];
            int32_t ls[8];
            int m = 0;
            for (int s = 0; s < width; ++s) {
                const WChild &ch = n.ch[s];
                if (ch.link == RT_LINK_NONE) continue;
                const float t = slab(ch.lo[0], ch.lo[1], ch.lo[2], ch.hi[0], ch.hi[1], ch.hi[2], o, idir, eps);
                if (t >= 0.0f && t < best.t) dist[m] = t, ls[m++] = ch.link;
            }
            if (mode == 0) {  // full sort
                for (int i = 1; i < m; ++i)
                    for (int j = i; j > 0 && dist[j] < dist[j - 1]; --j) std::swap(dist[j], dist[j - 1]), std::swap(ls[j], ls[j - 1]);
            } else if (m > 1) {  // only the nearest child is found (three compare-exchanges: (0,1) (2,3) (0,2)); the rest keeps its order
                int best = 0;
                for (int i = 1; i < m; ++i)
                    if (dist[i] < dist[best]) best = i;
                std::swap(dist[0], dist[best]);
                std::swap(ls[0], ls[best]);
                if (mode == 2 && m > 2) {  // ... plus the farthest to the bottom of the pushes
                    int far = 1;
                    for (int i = 2; i < m; ++i)
                        if (dist[i] > dist[far]) far = i;
                    std::swap(dist[m - 1], dist[far]);
                    std::swap(ls[m - 1], ls[far]);
                }
            }
            for (int i = m - 1; i >= 1; --i) sl[sp] = ls[i], stt[sp++] = dist[i], ++c.pushes;
            if (m > 0) {
                link = ls[0];
                continue;
            }
        } else {
            ++c.leaves;
            for (uint32_t k = (uint32_t)~link;; ++k) {
                ++c.tris;
                if (tri_hit(bvh, k, o, d, eps, best)) break;
            }
        }
        for (;;) {
            if (sp == 0) return best;
            --sp;
            ++c.pops;
            if (stt[sp] < best.t) {
                link = sl[sp];
                break;
            }
            ++c.culled;
        }
    }
}

void report(const char *name, const Cnt &c) {
    const double r = (double)c.rays;
    std::printf("%-34s nodes/ray %6.2f  leaves %5.2f  tris %5.2f  pushes %5.2f  pops %5.2f  culled pops %5.2f\n", name, c.nodes / r,
                c.leaves / r, c.tris / r, c.pushes / r, c.pops / r, c.culled / r);
}
}  // namespace

// Child boxes of a wide-node array replaced by what an 8-bit-per-plane encoding would decode to (conservative:
// every decoded box contains the exact one).  mode 0: the library's grid (power-of-two cell, 1/64-cell margin, origin
// with 14 mantissa bits); mode 1: cell = extent / 255 exactly (a float scale per axis), origin = the node's minimum;
// mode 2: like 0 without the margin; mode 3: 16 bits per plane on the power-of-two grid.
std::vector<WNode> requant(const std::vector<WNode> &W, int mode) {
    std::vector<WNode> out = W;
    for (WNode &n : out) {
        if (mode >= 4) {  // 4: exact boxes moved outward by one ulp; 5: by 1e-5 of the node's extent
            for (int k = 0; k < 3; ++k) {
                float mn = 1e30f, mx = -1e30f;
                for (int s = 0; s < 8; ++s)
                    if (n.ch[s].link != RT_LINK_NONE) mn = std::min(mn, n.ch[s].lo[k]), mx = std::max(mx, n.ch[s].hi[k]);
                for (int s = 0; s < 8; ++s) {
                    WChild &c = n.ch[s];
                    if (c.link == RT_LINK_NONE) continue;
                    if (mode == 4) c.lo[k] = std::nextafterf(c.lo[k], -INFINITY), c.hi[k] = std::nextafterf(c.hi[k], INFINITY);
                    else c.lo[k] -= 1e-5f * (mx - mn), c.hi[k] += 1e-5f * (mx - mn);
                }
            }
            continue;
        }
        for (int k = 0; k < 3; ++k) {
            double mn = 1e300, mx = -1e300;
            for (int s = 0; s < 8; ++s)
                if (n.ch[s].link != RT_LINK_NONE) mn = std::min<double>(mn, n.ch[s].lo[k]), mx = std::max<double>(mx, n.ch[s].hi[k]);
            if (mn > mx) continue;
            const double ext = mx - mn;
            const double steps = mode == 3 ? 65535.0 : 255.0;
            double cell, org, margin = (mode == 0) ? 1.0 / 64.0 : 0.0;
            if (mode == 1) {
                cell = ext > 0 ? ext / steps * (1.0 + 1e-6) : 1e-30;
                org = mn;
            } else {
                int e = ext > 0 ? (int)std::ceil(std::log2(ext / steps * (1.0 + 2.0 * margin / steps + 1e-9))) : -100;
                cell = std::ldexp(1.0, e);
                org = mn - margin * cell;
                if (mode == 0) {  // origin rounded down to 14 mantissa bits (the low 9 bits of the word hold the exponent)
                    float of = (float)org;
                    if ((double)of > org) of = std::nextafterf(of, -INFINITY);
                    uint32_t b;
                    std::memcpy(&b, &of, 4);
                    if (of >= 0) b &= ~0x1FFu; else b = (b | 0x1FFu);
                    std::memcpy(&of, &b, 4);
                    org = of;
                    while ((mx - org) / cell + margin > steps) cell *= 2.0;
                }
            }
            for (int s = 0; s < 8; ++s) {
                WChild &c = n.ch[s];
                if (c.link == RT_LINK_NONE) continue;
                const double a = std::floor((c.lo[k] - org) / cell - margin), z = std::ceil((c.hi[k] - org) / cell + margin);
                c.lo[k] = (float)(org + a * cell);
                if ((double)c.lo[k] > org + a * cell) c.lo[k] = std::nextafterf(c.lo[k], -INFINITY);
                c.hi[k] = (float)(org + z * cell);
                if ((double)c.hi[k] < org + z * cell) c.hi[k] = std::nextafterf(c.hi[k], INFINITY);
            }
        }
    }
    return out;
}

extern "C" int wide_study(const rt_scene_desc *sc, uint32_t w, uint32_t h, uint32_t spp) {
    PackedScene p;
    if (int rc = pack_scene(*sc, p, true)) return rc;
    DScene d;
    fill_scene_constants(*sc, p, d);
    d.scene.nodes = p.scene.nodes.data();
    d.scene.qnodes = p.scene.qnodes.data();
    d.scene.qnodes4 = p.scene.qnodes4.data();
    d.light.qnodes4 = p.light.qnodes4.data();
    d.light.qnodes = p.light.qnodes.data();
    d.scene.tris = p.scene.tris.data();
    d.light.nodes = p.light.nodes.data();
    d.light.tris = p.light.tris.data();
    d.light_sample = p.light_sample.data();
    d.attrs = p.attrs.data();
    d.tangents = p.tangents.empty() ? nullptr : p.tangents.data();
    d.light_extra = p.light_extra.data();
    d.materials = p.materials.data();
    d.textures = p.textures.data();
    d.texels = p.texels.data();
    std::vector<WNode> W4, W8;
    const int32_t r4 = collapse(p.scene.nodes, p.scene.root, 4, W4), r8 = collapse(p.scene.nodes, p.scene.root, 8, W8);
    std::vector<WNode> W8d;
    DpCollapse dp(p.scene.nodes);
    dp.solve(p.scene.root);
    const int32_t r8d = collapse_dp(dp, p.scene.root, W8d);
    std::vector<WNode> W4d;
    DpCollapse dp4(p.scene.nodes, 4);
    dp4.solve(p.scene.root);
    const int32_t r4d = collapse_dp(dp4, p.scene.root, W4d);
    std::printf("binary inner nodes %zu, 4-wide greedy %zu, cost-optimal %zu; 8-wide greedy %zu, cost-optimal %zu\n", p.scene.nodes.size(),
                W4.size(), W4d.size(), W8.size(), W8d.size());
    Camera c;
    c.pos = mk3(d.cam_pos[0], d.cam_pos[1], d.cam_pos[2]);
    c.right = mk3(d.cam_right[0], d.cam_right[1], d.cam_right[2]);
    c.up = mk3(d.cam_up[0], d.cam_up[1], d.cam_up[2]);
    c.fwd = mk3(d.cam_fwd[0], d.cam_fwd[1], d.cam_fwd[2]);
    c.tan_half_x = tanf(d.fov_x / 2);
    c.tan_half_y = tanf(atanf(tanf(d.fov_x / 2) * (float)h / (float)w));
    c.inv_w2 = 2.0f / (float)w;
    c.inv_h2 = 2.0f / (float)h;
    const std::vector<WNode> W4q0 = requant(W4, 0), W4q1 = requant(W4, 1), W4q2 = requant(W4, 2), W4q3 = requant(W4, 3), W4q4 = requant(W4, 4), W4q5 = requant(W4, 5);
    Cnt aq0, aq1, aq2, aq3, aq4, aq5;
    Cnt a4, b4, b4d, a8, b8, b8d, q4, b8o, a4o, a4m, a4m2;
    a4.per_node.assign(W4.size(), 0);
    aq5.per_node.assign(W4.size(), 0);
    uint64_t mism = 0;
    for (uint32_t pix = 0; pix < w * h; ++pix)
        for (uint32_t s = 0; s < spp; ++s) {
            const RngKey key{pix, s, 7u, 0u};
            const u4 j = rng_jitter(key);
            f3 o = c.pos;
            f3 dir = camera_dir(c, (float)(pix % w) + u01(j.x), (float)(pix / w) + u01(j.y));
            f3 thr = mk3(1, 1, 1), rad = mk3(0, 0, 0);
            for (uint32_t b = 0; b < d.ray_depth; ++b) {
                uint32_t st = 0;
                const Hit hit = closest_hit_q4(d.scene, o, dir, d.eps, &st);
                q4.nodes += st;
                ++q4.rays;
                const Hit h1 = trav_sorted(W4, r4, 4, d.scene, o, dir, d.eps, a4);
                const Hit h2 = trav_oct(W4, r4, 4, d.scene, o, dir, d.eps, b4, false);
                trav_oct(W4, r4, 4, d.scene, o, dir, d.eps, b4d, true);
                trav_sorted(W8, r8, 8, d.scene, o, dir, d.eps, a8);
                const Hit h3 = trav_oct(W8, r8, 8, d.scene, o, dir, d.eps, b8, false);
                trav_oct(W8, r8, 8, d.scene, o, dir, d.eps, b8d, true);
                trav_oct(W8d, r8d, 8, d.scene, o, dir, d.eps, b8o, false);
                trav_sorted(W4d, r4d, 4, d.scene, o, dir, d.eps, a4o);
                trav_sorted(W4, r4, 4, d.scene, o, dir, d.eps, a4m, 1);
                trav_sorted(W4, r4, 4, d.scene, o, dir, d.eps, a4m2, 2);
                trav_sorted(W4q0, r4, 4, d.scene, o, dir, d.eps, aq0);
                trav_sorted(W4q1, r4, 4, d.scene, o, dir, d.eps, aq1);
                trav_sorted(W4q2, r4, 4, d.scene, o, dir, d.eps, aq2);
                trav_sorted(W4q3, r4, 4, d.scene, o, dir, d.eps, aq3);
                trav_sorted(W4q4, r4, 4, d.scene, o, dir, d.eps, aq4);
                trav_sorted(W4q5, r4, 4, d.scene, o, dir, d.eps, aq5);
                mism += (h1.tri != hit.tri) + (h2.tri != hit.tri) + (h3.tri != hit.tri);
                uint32_t lr = 0;
                if (!shade_bounce(d, p.gamma_lut, key, b, b + 1 == d.ray_depth, hit, o, dir, thr, rad, lr)) break;
            }
        }
    std::printf("rays %llu, hit mismatches vs the quantised 4-wide traversal %llu\n", (unsigned long long)q4.rays, (unsigned long long)mism);
    report("k_extend's 4-wide (quantised)", q4);
    report("A4 sorted + distance stack", a4);
    report("B4 octant order, mask stack", b4);
    report("B4 distance order, mask stack", b4d);
    report("A8 sorted + distance stack", a8);
    report("B8 octant order, mask stack", b8);
    report("B8 distance order, mask stack", b8d);
    report("B8 cost-optimal collapse, octant", b8o);
    report("A4 cost-optimal collapse, sorted", a4o);
    report("A4 nearest first, rest unsorted", a4m);
    report("A4 nearest first, farthest last", a4m2);
    report("A4 boxes as the library quantises", aq0);
    report("A4 8 bit, exact cell = extent/255", aq1);
    report("A4 8 bit, power-of-two, no margin", aq2);
    report("A4 16 bit planes, power-of-two", aq3);
    report("A4 exact boxes + 1 ulp outward", aq4);
    report("A4 exact boxes + 1e-5 extent", aq5);
    {   // where do the extra visits of the slightly widened boxes happen?
        std::vector<int> depth(W4.size(), 0), order;
        std::vector<int32_t> st{r4};
        while (!st.empty()) {
            const int32_t n = st.back();
            st.pop_back();
            for (int s = 0; s < 4; ++s)
                if (W4[n].ch[s].link != RT_LINK_NONE && W4[n].ch[s].link >= 0) depth[W4[n].ch[s].link] = depth[n] + 1, st.push_back(W4[n].ch[s].link);
        }
        uint64_t by_depth_a[32] = {0}, by_depth_b[32] = {0};
        for (size_t i = 0; i < W4.size(); ++i) by_depth_a[std::min(depth[i], 31)] += a4.per_node[i], by_depth_b[std::min(depth[i], 31)] += aq5.per_node[i];
        for (int dd = 0; dd < 14; ++dd) std::printf("depth %2d: exact %8llu  widened %8llu\n", dd, (unsigned long long)by_depth_a[dd], (unsigned long long)by_depth_b[dd]);
        size_t worst = 0;
        for (size_t i = 0; i < W4.size(); ++i)
            if ((int64_t)aq5.per_node[i] - a4.per_node[i] > (int64_t)aq5.per_node[worst] - a4.per_node[worst]) worst = i;
        std::printf("node %zu depth %d: %u -> %u visits\n", worst, depth[worst], a4.per_node[worst], aq5.per_node[worst]);
        for (int s = 0; s < 4; ++s) {
            const WChild &c = W4[worst].ch[s];
            if (c.link != RT_LINK_NONE) std::printf("  child %d link %d lo %.6f %.6f %.6f hi %.6f %.6f %.6f\n", s, c.link, c.lo[0], c.lo[1], c.lo[2], c.hi[0], c.hi[1], c.hi[2]);
        }
    }
    return 0;
}
