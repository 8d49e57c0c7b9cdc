// rt_bvh.hpp — host-side construction of the EXACT culling structure used by kernel variant 3.
//
// Why this exists. The reference's KD traversal (KDTree.cpp:31-69) enters BOTH children of every
// node whose box the ray touches, never orders them and never stops early, and its build leaves
// 30-170-triangle leaves and 3-5x duplicated references: per ray that is ~130 box tests and ~580
// triangle tests on the pond scene for an answer that depends on a handful of them. The result of
// that traversal, however, has a closed form:
//
//     t*  =  min { t(T) : triangle T is hit by the ray (Triangle::getIntersection, unchanged
//                         arithmetic) and T lies in at least one leaf whose whole ancestor chain
//                         of boxes passes AABB::intersects for this ray }
//
// (t(T) does not depend on which leaf holds the reference; ties between different triangles go to
// the last such leaf in depth-first order, first reference inside it.) Variant 3 evaluates exactly
// that set expression, but finds the candidate triangles with a conventional bounding-volume
// hierarchy over the mesh's distinct triangles instead of walking the reference tree:
//   * the BVH only PRUNES — boxes are padded and tested conservatively in fp32, so it can return
//     extra candidates but never lose a triangle the reference would hit;
//   * every candidate is tested with the reference's own triangle arithmetic (rt::triangle_t);
//   * a candidate that would become the new minimum must be REACHABLE in the reference tree.
//     Almost always its hit parameter lies strictly inside the slab interval of a leaf that holds
//     it (with a margin far above rounding), which implies every ancestor test passes; otherwise the
//     ancestor chain is re-tested with the reference's fp64 slab arithmetic (rt::slab_hit).
// The arrays built here are derived data: they never cross the C ABI and the host KD-tree stays the
// single source of truth (tests compare variant 3 with the reference-order variants bit for bit).
#ifndef HAI719_RT_BVH_HPP
#define HAI719_RT_BVH_HPP
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>
#include "hai719_rt.h"
#include "rt_pack.hpp"

namespace rt {

struct Accel {
    // BVH, 4 float4 per node: {c0lo.xyz, c0hi.x} {c0hi.yz, c1lo.xy} {c1lo.z, c1hi.xyz} {bits child0, bits child1, -, -}
    // child >= 0: inner node index; child < 0: leaf, -(child+1) = first*8 + count (count 1..7) into bvh_tris
    std::vector<float4> nodes;
    std::vector<uint32_t> tris;          // representative (first) leaf-ref index of each distinct triangle, in BVH leaf order
    std::vector<int32_t> mesh_root;      // per mesh: root node index, -1 = no hierarchy
    // the same hierarchy with 4-wide nodes (collapse4 below), 8 float4 per node: lo.x lo.y lo.z hi.x hi.y hi.z of the four
    // children (one float4 per coordinate), their codes, one spare; unused slots have an inverted box and an empty leaf
    std::vector<float4> nodes4;
    std::vector<int32_t> mesh_root4;     // per mesh: root node index in nodes4, -1 = no hierarchy
    // per mesh: range of `tris` holding the triangles that are tested for EVERY ray (barycentric test too
    // ill-conditioned to bound the region it accepts, see build_accel)
    std::vector<uint32_t> always_first, always_count;
    // reference-tree bookkeeping for the reachability test
    std::vector<uint32_t> ref_next;      // per leaf ref: next leaf ref (ascending) of the same triangle, 0xFFFFFFFF = none
    std::vector<uint32_t> ref_leaf;      // per leaf ref: node index (shared node array) of the leaf holding it
    std::vector<uint32_t> node_parent;   // per node of the shared node array: parent index, 0xFFFFFFFF for a mesh's synthetic root
};

// Builder knobs, measured on B200 (profiles/r01_notes.md). Any choice gives the same pixels: the hierarchies only cull.
//   SAH over all three axes instead of the widest one: config 3 73.4 -> 68.3 ms (2-spp frame), configs 4/5 -1 %
//   leaves of <= 2 triangles (1: config 3 +8 %, 4: config 4 +8 %); leaves of ONE sphere/square (config 2 -7 %: a sphere test
//   costs as much as two box tests, so a box that rules it out pays); 32 bins (-1 % against 16)
#ifndef RT_BVH_SAH_AXES
#define RT_BVH_SAH_AXES 3
#endif
#ifndef RT_BVH_LEAF_TRIS
#define RT_BVH_LEAF_TRIS 2
#endif
#ifndef RT_BVH_KAPPA_MAX
#define RT_BVH_KAPPA_MAX 4e5   /* conditioning beyond which a triangle is tested for every ray instead of being boxed (build_accel) */
#endif
#ifndef RT_BVH_LEAF_ANALYTIC
#define RT_BVH_LEAF_ANALYTIC 1
#endif
namespace bvh_detail {
struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const float *p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return (dx < 0 || dy < 0 || dz < 0) ? 0.f : 2.f * (dx * dy + dy * dz + dz * dx);
    }
};
struct Prim { Box box; float c[3]; uint32_t ref; float coef = 0.f; };

inline int32_t leaf_code(uint32_t first, uint32_t count) { return -(int32_t)(first * 8u + count) - 1; }

// The barycentric denominator the device stores for a triangle with the (already scaled) float corners c0, c1, c2 — precompute_triangle's
// arithmetic replayed on the host: float edges, float dot products summed left to right, d00 * d11 - d01 * d01, every operation rounded to
// float and none fused (volatile). tests/test_hostsim.py compares it bit for bit with precompute_triangle on every triangle of the assets.
inline float stored_denominator(const float *c0, const float *c1, const float *c2) {
    const volatile float e0x = c1[0] - c0[0], e0y = c1[1] - c0[1], e0z = c1[2] - c0[2];
    const volatile float e1x = c2[0] - c0[0], e1y = c2[1] - c0[1], e1z = c2[2] - c0[2];
    const float e0[3] = {e0x, e0y, e0z}, e1[3] = {e1x, e1y, e1z};
    auto fdot = [](const float *a, const float *b) { const volatile float x = a[0] * b[0], y = a[1] * b[1], z = a[2] * b[2]; const volatile float xy = x + y; return (float)(xy + z); };
    const float f00 = fdot(e0, e0), f01 = fdot(e0, e1), f11 = fdot(e1, e1);
    const volatile float pa = f00 * f11, pb = f01 * f01;
    return pa - pb;
}

// returns a child code (>= 0 inner node, < 0 leaf). `median` forces balanced splits (depth <= log2 n),
// used when the SAH tree came out deeper than the device's traversal stack.
template <class A>
inline int32_t build(std::vector<Prim> &prims, size_t begin, size_t end, A &out, bool median, int depth, int &max_depth, size_t LEAF) {
    const size_t n = end - begin;
    max_depth = std::max(max_depth, depth);
    Box cb; cb.reset();
    for (size_t i = begin; i < end; ++i) cb.grow(prims[i].c);
    int axis = 0;
    float ext = -1.f;
    for (int a = 0; a < 3; ++a) if (cb.hi[a] - cb.lo[a] > ext) { ext = cb.hi[a] - cb.lo[a]; axis = a; }
    size_t mid = begin;
    if (n > LEAF && median) {
        mid = begin + n / 2;
        std::nth_element(prims.begin() + begin, prims.begin() + mid, prims.begin() + end,
                         [&](const Prim &a, const Prim &b) { return a.c[axis] < b.c[axis]; });
    } else if (n > LEAF && ext > 0.f) {
        // binned SAH: RT_BVH_SAH_AXES = 1 widest centroid axis only, 3 = the best split over all three axes
        const int NB = 32, NBMAX = NB;
        float best = FLT_MAX; int bs = -1, best_axis = axis;
        for (int a = 0; a < 3; ++a) {
            if (RT_BVH_SAH_AXES == 1 && a != axis) continue;
            const float ea = cb.hi[a] - cb.lo[a];
            if (!(ea > 0.f)) continue;
            Box bb[NBMAX]; int bc[NBMAX];
            for (int b = 0; b < NB; ++b) { bb[b].reset(); bc[b] = 0; }
            const float k = NB * (1.f - 1e-6f) / ea;
            for (size_t i = begin; i < end; ++i) {
                int b = (int)((prims[i].c[a] - cb.lo[a]) * k);
                b = std::max(0, std::min(NB - 1, b));
                bb[b].grow(prims[i].box); bc[b]++;
            }
            float la[NBMAX], ra[NBMAX]; int lc[NBMAX], rc[NBMAX];
            Box acc; acc.reset(); int cnt = 0;
            for (int b = 0; b < NB; ++b) { acc.grow(bb[b]); cnt += bc[b]; la[b] = acc.area(); lc[b] = cnt; }
            acc.reset(); cnt = 0;
            for (int b = NB - 1; b >= 0; --b) { acc.grow(bb[b]); cnt += bc[b]; ra[b] = acc.area(); rc[b] = cnt; }
            for (int b = 0; b + 1 < NB; ++b) {
                if (lc[b] == 0 || rc[b + 1] == 0) continue;
                const float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
                if (cost < best) { best = cost; bs = b; best_axis = a; }
            }
        }
        if (bs >= 0) {
            const int a = best_axis;
            const float k = NB * (1.f - 1e-6f) / (cb.hi[a] - cb.lo[a]);
            auto it = std::partition(prims.begin() + begin, prims.begin() + end, [&](const Prim &p) {
                int b = (int)((p.c[a] - cb.lo[a]) * k);
                b = std::max(0, std::min(NB - 1, b));
                return b <= bs;
            });
            mid = (size_t)(it - prims.begin());
        }
    }
    if (n <= LEAF || ((mid == begin || mid == end) && n <= 7)) {
        const uint32_t first = (uint32_t)out.tris.size();
        for (size_t i = begin; i < end; ++i) out.tris.push_back(prims[i].ref);
        return leaf_code(first, (uint32_t)n);
    }
    if (mid == begin || mid == end) {   // all centroids equal but too many for one leaf: split by count
        mid = begin + n / 2;
    }
    const size_t id = out.nodes.size() / 4;
    out.nodes.resize(out.nodes.size() + 4);
    Box b0, b1; b0.reset(); b1.reset();
    float k0 = 0.f, k1 = 0.f;   // largest per-primitive padding coefficient below each child (0 for triangles)
    for (size_t i = begin; i < mid; ++i) { b0.grow(prims[i].box); k0 = std::max(k0, prims[i].coef); }
    for (size_t i = mid; i < end; ++i) { b1.grow(prims[i].box); k1 = std::max(k1, prims[i].coef); }
    const int32_t c0 = build(prims, begin, mid, out, median, depth + 1, max_depth, LEAF);
    const int32_t c1 = build(prims, mid, end, out, median, depth + 1, max_depth, LEAF);
    out.nodes[4 * id + 0] = make_float4(b0.lo[0], b0.lo[1], b0.lo[2], b0.hi[0]);
    out.nodes[4 * id + 1] = make_float4(b0.hi[1], b0.hi[2], b1.lo[0], b1.lo[1]);
    out.nodes[4 * id + 2] = make_float4(b1.lo[2], b1.hi[0], b1.hi[1], b1.hi[2]);
    out.nodes[4 * id + 3] = make_float4(u2f((uint32_t)c0), u2f((uint32_t)c1), k0, k1);
    return (int32_t)id;
}
}  // namespace bvh_detail

// Collapse a binary subtree (child code `code` of Accel::nodes) into 4-wide nodes: a node adopts its two children, then
// replaces the inner child with the largest box by that child's two children until it has four (or only leaves are left).
// Returns the code of the subtree in nodes4 (>= 0 node index, < 0 the same leaf code); depth = levels of 4-wide nodes.
inline int32_t collapse4(const std::vector<float4> &bin, int32_t code, std::vector<float4> &out4, int depth, int &max_depth) {
    if (code < 0) return code;
    max_depth = std::max(max_depth, depth + 1);
    struct Child { int32_t code; float lo[3], hi[3]; };
    auto children = [&](int32_t n, Child &c0, Child &c1) {
        const float4 a = bin[4 * (size_t)n], b = bin[4 * (size_t)n + 1], c = bin[4 * (size_t)n + 2], k = bin[4 * (size_t)n + 3];
        c0.lo[0] = a.x; c0.lo[1] = a.y; c0.lo[2] = a.z; c0.hi[0] = a.w; c0.hi[1] = b.x; c0.hi[2] = b.y;
        c1.lo[0] = b.z; c1.lo[1] = b.w; c1.lo[2] = c.x; c1.hi[0] = c.y; c1.hi[1] = c.z; c1.hi[2] = c.w;
        c0.code = (int32_t)f2u(k.x); c1.code = (int32_t)f2u(k.y);
    };
    auto area = [](const Child &c) {
        const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
        return (dx < 0 || dy < 0 || dz < 0) ? -1.f : dx * dy + dy * dz + dz * dx;
    };
    Child ch[4];
    int n = 2;
    children(code, ch[0], ch[1]);
    while (n < 4) {
        int best = -1;
        for (int i = 0; i < n; ++i) if (ch[i].code >= 0 && (best < 0 || area(ch[i]) > area(ch[best]))) best = i;
        if (best < 0) break;
        Child a, b;
        children(ch[best].code, a, b);
        ch[best] = a;
        ch[n++] = b;
    }
    const size_t id = out4.size() / 8;
    out4.resize(out4.size() + 8);
    float lo[3][4], hi[3][4];
    int32_t codes[4];
    for (int i = 0; i < 4; ++i) {
        if (i < n) {
            for (int a = 0; a < 3; ++a) { lo[a][i] = ch[i].lo[a]; hi[a][i] = ch[i].hi[a]; }
            codes[i] = collapse4(bin, ch[i].code, out4, depth + 1, max_depth);
        } else {
            for (int a = 0; a < 3; ++a) { lo[a][i] = FLT_MAX; hi[a][i] = -FLT_MAX; }
            codes[i] = bvh_detail::leaf_code(0, 0);
        }
    }
    for (int a = 0; a < 3; ++a) {
        out4[8 * id + a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        out4[8 * id + 3 + a] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
    out4[8 * id + 6] = make_float4(u2f((uint32_t)codes[0]), u2f((uint32_t)codes[1]), u2f((uint32_t)codes[2]), u2f((uint32_t)codes[3]));
    out4[8 * id + 7] = make_float4(0.f, 0.f, 0.f, 0.f);
    return (int32_t)id;
}

inline void build_accel(const RtSceneDesc &d, const PackedMeshes &pk, Accel &out) {
    using namespace bvh_detail;
    out = Accel();
    const uint32_t NONE = 0xFFFFFFFFu;
    out.ref_next.assign(pk.total_refs, NONE);
    out.ref_leaf.assign(pk.total_refs, NONE);
    out.node_parent.assign(pk.lo.size(), NONE);
    for (uint32_t mi = 0; mi < d.n_meshes; ++mi) {
        const RtSceneMesh &m = d.meshes[mi];
        // parents and leaf owners from the pre-order array
        std::vector<uint32_t> stack;
        for (uint32_t i = pk.node_begin[mi]; i < pk.node_end[mi]; ++i) {
            const uint32_t hw = f2u(pk.hi[i].w);
            const bool leaf = (hw & 0x80000000u) != 0u;
            while (!stack.empty() && f2u(pk.lo[stack.back()].w) <= i) stack.pop_back();
            out.node_parent[i] = stack.empty() ? NONE : stack.back();
            if (leaf) {
                const uint32_t first = f2u(pk.lo[i].w), cnt = hw & 0x7FFFFFFFu;
                for (uint32_t r = first; r < first + cnt; ++r) out.ref_leaf[r] = i;
            } else {
                stack.push_back(i);
            }
        }
        // distinct triangles: chain their references in ascending order
        const uint32_t rb = pk.ref_begin[mi];
        std::vector<uint32_t> first_ref(m.n_triangles, NONE), last_ref(m.n_triangles, NONE);
        for (uint32_t k = 0; k < m.n_leaf_refs; ++k) {
            const uint32_t T = m.leaf_refs[k].tri_index, r = rb + k;
            if (first_ref[T] == NONE) first_ref[T] = r; else out.ref_next[last_ref[T]] = r;
            last_ref[T] = r;
        }
        std::vector<Prim> prims;
        std::vector<uint32_t> always;
        for (uint32_t T = 0; T < m.n_triangles; ++T) {
            if (first_ref[T] == NONE) continue;   // dropped by the depth-100 cut-off: never hit (KDTree.cpp:101-103)
            const RtTriRef &r = m.leaf_refs[first_ref[T] - rb];
            Prim p; p.box.reset(); p.ref = first_ref[T];
            bool finite = true;
            double c[3][3];
            for (int v = 0; v < 3; ++v) {
                float q[3];
                for (int a = 0; a < 3; ++a) { q[a] = 1.000001f * m.positions[3 * r.v[v] + a]; finite = finite && std::isfinite(q[a]); c[v][a] = q[a]; }
                p.box.grow(q);
            }
            if (!finite) continue;                // a NaN/inf vertex can only produce "no intersection"
            // How far outside the true triangle can a point be and still pass the reference's barycentric
            // test? u1, u2 are quotients by denom = d00*d11 - d01^2 (Triangle.h:62-75), which cancels
            // catastrophically for slivers: the slop is ~ eps * kappa * (longest edge), kappa = d00*d11/denom
            // = 1/sin^2 of the corner angle (measured: 2.8e-4 on a 1.7-long, 4.5-degree sliver of pond.off).
            double e0[3], e1[3], d00 = 0, d01 = 0, d11 = 0;
            for (int a = 0; a < 3; ++a) { e0[a] = c[1][a] - c[0][a]; e1[a] = c[2][a] - c[0][a]; d00 += e0[a] * e0[a]; d01 += e0[a] * e1[a]; d11 += e1[a] * e1[a]; }
            const double denom = d00 * d11 - d01 * d01;
            // zero area in exact arithmetic (the fp32 inputs are exact in fp64; only the last products round):
            // identical or collinear corners. If the fp32 cross product is exactly 0 the normal is 0/0 = NaN and
            // every test fails; if rounding leaves a residue the plane is garbage but finite, so
            // such a triangle goes to the always-tested list rather than being dropped.
            if (!(denom > 0.0)) {
                // the device's own fp32 arithmetic (precompute_triangle), replayed on the host without FMA
                const float fe0[3] = {(float)c[1][0] - (float)c[0][0], (float)c[1][1] - (float)c[0][1], (float)c[1][2] - (float)c[0][2]};
                const float fe1[3] = {(float)c[2][0] - (float)c[0][0], (float)c[2][1] - (float)c[0][1], (float)c[2][2] - (float)c[0][2]};
                const volatile float m0 = fe0[1] * fe1[2], m1 = fe0[2] * fe1[1], m2 = fe0[2] * fe1[0], m3 = fe0[0] * fe1[2],
                                     m4 = fe0[0] * fe1[1], m5 = fe0[1] * fe1[0];
                const float nx = m0 - m1, ny = m2 - m3, nz = m4 - m5;
                if (nx == 0.f && ny == 0.f && nz == 0.f) continue;   // n = 0/0 = NaN: dotRN < 0 is never true
                always.push_back(p.ref);
                continue;
            }
            {
                // When the denominator the device STORES is exactly 0 or not finite, u1 = N / den is never inside [0, 1] (Triangle.h:62-75:
                // +-inf or NaN fail every comparison) and the triangle can never report a hit: dropped. One of the two collinear slivers of
                // the triceratops mesh is such a triangle; as an always-tested one it cost every ray a test.
                const float q0[3] = {(float)c[0][0], (float)c[0][1], (float)c[0][2]}, q1[3] = {(float)c[1][0], (float)c[1][1], (float)c[1][2]},
                            q2[3] = {(float)c[2][0], (float)c[2][1], (float)c[2][2]};
                const float fden = bvh_detail::stored_denominator(q0, q1, q2);
                if (fden == 0.f || !std::isfinite(fden)) continue;
            }
            const double kappa = d00 * d11 / denom;
            const double lmax = std::sqrt(std::max(d00, d11));
            // error analysis of u1 = (d11*d20 - d01*d21)/denom in fp32 gives <= ~36 eps kappa lmax of spatial
            // slop; 128 leaves a 3.5x margin. Beyond kappa = 1e5 the fp32 denominator has lost most of its
            // bits (its sign can flip near 1e7) and the accepted region is not usefully bounded: those few
            // triangles (<= 8 per mesh in the reference's assets) are simply tested for every ray.
            // (Round 2: the limit was 1e5, and five triangles of pond.off sit at kappa = 1.09e5, two of the pool scene at 1.05e5 and
            // 3.1e5: as always-tested triangles they were candidates of EVERY light cone — 27 % of the pond scene's (hit, light) pairs
            // were sampled for them alone — and tested by every ray. The fp32 denominator's relative error is <= ~12 eps kappa = 0.29 at
            // 4e5: its sign holds and the quotients grow by at most 1 / (1 - 0.29), so the first-order bound 36 eps kappa lmax becomes
            // 51: the factor 128 below still leaves 2.5x. profiles/r02_notes.md, r03y.)
            if (!(kappa <= RT_BVH_KAPPA_MAX)) { always.push_back(p.ref); continue; }
            const double slop = 128.0 * 5.96e-8 * kappa * lmax;
            for (int a = 0; a < 3; ++a) {
                // + hundreds of ulps of the coordinates involved, plus an absolute floor
                const float pad = (float)slop + 1e-4f * (std::fabs(p.box.lo[a]) + std::fabs(p.box.hi[a]) + (p.box.hi[a] - p.box.lo[a])) + 1e-5f;
                p.box.lo[a] -= pad; p.box.hi[a] += pad;
                p.c[a] = 0.5f * (p.box.lo[a] + p.box.hi[a]);
            }
            prims.push_back(p);
        }
        out.always_first.push_back((uint32_t)out.tris.size());
        out.always_count.push_back((uint32_t)always.size());
        out.tris.insert(out.tris.end(), always.begin(), always.end());
        if (prims.empty()) { out.mesh_root.push_back(-1); out.mesh_root4.push_back(-1); continue; }
        const size_t nodes_mark = out.nodes.size(), tris_mark = out.tris.size();
        int max_depth = 0;
        int32_t root = build(prims, 0, prims.size(), out, false, 0, max_depth, RT_BVH_LEAF_TRIS);
        if (max_depth > 56) {   // device stack holds 64 entries
            out.nodes.resize(nodes_mark); out.tris.resize(tris_mark);
            max_depth = 0;
            root = build(prims, 0, prims.size(), out, true, 0, max_depth, RT_BVH_LEAF_TRIS);
        }
        if (root >= 0) {
            out.mesh_root.push_back(root);
        } else {
            // a mesh so small that it is a single leaf: wrap it in a node whose second child is empty
            Box b; b.reset();
            for (const Prim &p : prims) b.grow(p.box);
            const size_t id = out.nodes.size() / 4;
            out.nodes.push_back(make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]));
            out.nodes.push_back(make_float4(b.hi[1], b.hi[2], FLT_MAX, FLT_MAX));
            out.nodes.push_back(make_float4(FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX));
            out.nodes.push_back(make_float4(u2f((uint32_t)root), u2f((uint32_t)leaf_code(0, 0)), 0.f, 0.f));
            out.mesh_root.push_back((int32_t)id);
        }
        {   // 4-wide copy of this mesh's hierarchy; the walk pushes up to three children per level on a 64-entry stack
            const size_t mark4 = out.nodes4.size();
            int depth4 = 0;
            int32_t r4 = collapse4(out.nodes, out.mesh_root.back(), out.nodes4, 0, depth4);
            if (3 * depth4 > 60) { out.nodes4.resize(mark4); r4 = -1; }   // deeper than the stack allows: this mesh keeps the binary walk
            out.mesh_root4.push_back(r4);
        }
    }
}

// ---- the same idea for the analytic primitives -------------------------------------------------
// Scene::computeIntersection / computeShadow test every sphere and every square for every ray
// (Scene.h:207-220, 236-247): 82 sphere tests per ray on config 2. Each test is independent of the
// others, so a conservative hierarchy over their (motion-swept) bounds may skip tests that cannot
// produce an acceptable hit; accepted candidates go through the reference's own sphere / square
// arithmetic, and the ORDER semantics are restored explicitly: closest hit = smallest t, the earliest
// primitive in the reference's sequence (spheres by index, then squares by index) winning ties;
// occlusion = candidates collected in a bit mask and replayed in sequence order so that the
// random_float() draws of computeShadow happen exactly as in the reference.
// Conservativeness: a sphere test can report a hit for a ray that misses the true sphere by up to
// ~eps*|o-c|^2/(2r) (the discriminant cancels for grazing rays), so boxes are enlarged PER RAY by
// k_ray*coef with coef = 1/r and k_ray = 32 eps (|o-C|+R)^2, (C, R) a bounding sphere of all
// analytic primitives, on top of a static 1e-4 relative pad; squares get coef 0 and the linear term.
struct AnalyticAccel {
    std::vector<float4> nodes;     // same node format as Accel; n3.z / n3.w = max 1/r below child 0 / 1
    std::vector<uint32_t> tris;    // leaf contents: sequence index (sphere i -> i, square j -> n_spheres + j)
    int32_t root = -1;             // -1: not built (too few or too many primitives)
    float center[3] = {0, 0, 0};
    float radius = 0.f;
    // a ball that holds every sphere CENTRE at every time in [0, 1]: |o - c_i| <= |o - center_s| + radius_s for every sphere i, which is all
    // the quadratic term of the per-ray padding needs (eps |o - c|^2 / r); (center, radius) also holds the squares — one large ground
    // square put R at ~70 in the random-spheres scene and padded a sphere of radius 0.2 by 0.05 (profiles/r02_notes.md, r03r/s)
    float center_s[3] = {0, 0, 0};
    float radius_s = 0.f;
    // <= 32 primitives: their padded, motion-swept boxes in sequence order, {lo.xyz, coef} {hi.xyz, 0} (intersect_lc's flat test)
    std::vector<float4> flat;
};

inline void build_analytic_accel(const RtSceneDesc &d, AnalyticAccel &out) {
    using namespace bvh_detail;
    out = AnalyticAccel();
    const uint32_t n = d.n_spheres + d.n_squares;
    // above 128: the occlusion mask is 128 bits. Below 24 the reference's linear loops are as fast for closest-hit rays,
    // but a scene with lights AND meshes still gets the (tiny) hierarchy: it is what gives it the per-light candidate
    // walk of variants 5/6 (occluder mask + candidate triangle list instead of a mesh walk per shadow sample).
    const bool small_ok = d.n_lights > 0 && d.n_meshes > 0;
    if (n == 0 || n > 128 || (n < 24 && !small_ok)) return;
    std::vector<Prim> prims;
    Box all; all.reset();
    for (uint32_t i = 0; i < d.n_spheres; ++i) {
        const RtSphere &s = d.spheres[i];
        Prim p; p.box.reset(); p.ref = i;
        const float r = std::fabs(s.radius);
        for (int e = 0; e < 2; ++e) {      // time 0 and time 1 (ray.time is in [0,1), motion is linear)
            float lo[3], hi[3];
            for (int a = 0; a < 3; ++a) { const float c = s.center[a] + (float)e * s.material.motion[a]; lo[a] = c - r; hi[a] = c + r; }
            p.box.grow(lo); p.box.grow(hi);
        }
        p.coef = r > 0.f ? 1.f / r : 1e30f;
        prims.push_back(p);
    }
    for (uint32_t j = 0; j < d.n_squares; ++j) {
        const RtSquare &q = d.squares[j];
        Prim p; p.box.reset(); p.ref = d.n_spheres + j;
        for (int e = 0; e < 2; ++e)
            for (int corner = 0; corner < 4; ++corner) {
                float c[3];
                for (int a = 0; a < 3; ++a) {
                    const float right = q.v1[a] - q.v0[a], up = q.v3[a] - q.v0[a];
                    c[a] = q.v0[a] + (float)e * q.material.motion[a] + ((corner & 1) ? right : 0.f) + ((corner & 2) ? up : 0.f);
                }
                p.box.grow(c);
            }
        p.coef = 0.f;
        prims.push_back(p);
    }
    for (Prim &p : prims) {
        bool finite = true;
        for (int a = 0; a < 3; ++a) finite = finite && std::isfinite(p.box.lo[a]) && std::isfinite(p.box.hi[a]);
        if (!finite) { for (int a = 0; a < 3; ++a) { p.box.lo[a] = -1e30f; p.box.hi[a] = 1e30f; } }
        for (int a = 0; a < 3; ++a) {
            const float pad = 1e-4f * (std::fabs(p.box.lo[a]) + std::fabs(p.box.hi[a]) + (p.box.hi[a] - p.box.lo[a])) + 1e-4f;
            p.box.lo[a] -= pad; p.box.hi[a] += pad;
            p.c[a] = 0.5f * (p.box.lo[a] + p.box.hi[a]);
        }
        if (finite) all.grow(p.box);
    }
    float r2 = 0.f;
    for (int a = 0; a < 3; ++a) { out.center[a] = 0.5f * (all.lo[a] + all.hi[a]); const float h = 0.5f * (all.hi[a] - all.lo[a]); r2 += h * h; }
    out.radius = std::sqrt(r2);
    if (prims.size() <= 32) {   // prims are still in sequence order here (build() permutes them)
        for (const Prim &p : prims) {
            out.flat.push_back(make_float4(p.box.lo[0], p.box.lo[1], p.box.lo[2], p.coef));
            out.flat.push_back(make_float4(p.box.hi[0], p.box.hi[1], p.box.hi[2], 0.f));
        }
    }
    {
        Box cs; cs.reset();
        bool any = false;
        for (uint32_t i = 0; i < d.n_spheres; ++i) {
            const RtSphere &s = d.spheres[i];
            float c0[3], c1[3];
            bool finite = true;
            for (int a = 0; a < 3; ++a) { c0[a] = s.center[a]; c1[a] = s.center[a] + s.material.motion[a]; finite = finite && std::isfinite(c0[a]) && std::isfinite(c1[a]); }
            if (!finite) continue;   // its box is everything: always visited, whatever the padding
            cs.grow(c0); cs.grow(c1); any = true;
        }
        float q2 = 0.f;
        for (int a = 0; a < 3; ++a) {
            out.center_s[a] = any ? 0.5f * (cs.lo[a] + cs.hi[a]) : out.center[a];
            const float h = any ? 0.5f * (cs.hi[a] - cs.lo[a]) : 0.f;
            q2 += h * h;
        }
        out.radius_s = std::sqrt(q2) * 1.0001f + 1e-6f;   // the device adds time * motion in float
    }
    int max_depth = 0;
    int32_t root = build(prims, 0, prims.size(), out, false, 0, max_depth, RT_BVH_LEAF_ANALYTIC);
    if (max_depth > 56) { out.nodes.clear(); out.tris.clear(); max_depth = 0; root = build(prims, 0, prims.size(), out, true, 0, max_depth, RT_BVH_LEAF_ANALYTIC); }
    if (root < 0) {
        // one or two primitives: a single leaf. Wrap it in a node whose second child is empty (inverted box)
        Box b; b.reset();
        float k0 = 0.f;
        for (const Prim &p : prims) { b.grow(p.box); k0 = std::max(k0, p.coef); }
        const size_t id = out.nodes.size() / 4;
        out.nodes.push_back(make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]));
        out.nodes.push_back(make_float4(b.hi[1], b.hi[2], FLT_MAX, FLT_MAX));
        out.nodes.push_back(make_float4(FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX));
        out.nodes.push_back(make_float4(u2f((uint32_t)root), u2f((uint32_t)leaf_code(0, 0)), k0, 0.f));
        root = (int32_t)id;
    }
    out.root = root;
}

}  // namespace rt
#endif
