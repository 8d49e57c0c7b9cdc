#include "KDTree.hpp"
#include <algorithm>

namespace {
// box of one triangle on one axis (Triangle::getAABB, Triangle.h:128-140)
inline void tri_extent(const std::vector<MeshVertex> &V, const MeshTriangle &t, unsigned int axis, float &lo, float &hi) {
    const float a = V[t[0]].position[axis], b = V[t[1]].position[axis], c = V[t[2]].position[axis];
    lo = std::min(std::min(std::min(FLT_MAX, a), b), c);
    hi = std::max(std::max(std::max(-FLT_MAX, a), b), c);
}
inline void put_box(RtKdNode &n, const AABB &b) {
    for (unsigned int i = 0; i < 3; ++i) { n.bmin[i] = b.p0[i]; n.bmax[i] = b.p1[i]; }
}
}  // namespace

KDTree::KDTree(const std::vector<MeshTriangle> &triangles, const AABB &aabb, const std::vector<MeshVertex> &vertices)
    : vertices(vertices), aabb(aabb), root(-1) {
    root = buildTree(triangles, aabb, 0);
}

AABBCuttingPlane KDTree::cut(const std::vector<MeshTriangle> &triangles, int depth) const {
    const unsigned int axis = depth % 3;
    std::vector<float> mins;
    mins.reserve(triangles.size());
    for (const MeshTriangle &t : triangles) {
        float lo, hi;
        tri_extent(vertices, t, axis, lo, hi);
        mins.push_back(lo);
    }
    std::sort(mins.begin(), mins.end());
    return AABBCuttingPlane(axis, (float)((double)mins[mins.size() / 2] + EPSILON));
}

int KDTree::emit_leaf(const std::vector<MeshTriangle> &triangles, const AABB &box) {
    RtKdNode n{};
    put_box(n, box);
    n.is_leaf = 1;
    n.first_ref = (uint32_t)leaf_refs.size();
    n.n_refs = (uint32_t)triangles.size();
    for (const MeshTriangle &t : triangles) leaf_refs.push_back(RtTriRef{{t[0], t[1], t[2]}, t[3]});
    const int id = (int)nodes.size();
    n.skip = (uint32_t)id + 1;
    nodes.push_back(n);
    return id;
}

int KDTree::buildTree(const std::vector<MeshTriangle> &triangles, const AABB &box, unsigned int depth) {
    if (triangles.empty() || depth > KDTREE_MAX_DEPTH) return -1;
    max_depth_ = std::max(max_depth_, depth);
    if ((int)triangles.size() <= KDTREE_TRIANGLES_PER_LEAF) return emit_leaf(triangles, box);

    const AABBCuttingPlane plane = cut(triangles, depth);
    const std::pair<AABB, AABB> halves = box.split(plane);

    std::vector<MeshTriangle> left, right;
    for (const MeshTriangle &t : triangles) {
        float lo, hi;
        tri_extent(vertices, t, plane.axis, lo, hi);
        if (hi <= plane.position - EPSILON) left.push_back(t);
        else if (lo >= plane.position + EPSILON) right.push_back(t);
        else { left.push_back(t); right.push_back(t); }
    }
    if (left.size() == right.size()) return emit_leaf(triangles, box);

    const int id = (int)nodes.size();
    RtKdNode n{};
    put_box(n, box);
    nodes.push_back(n);
    const int l = buildTree(left, halves.first, depth + 1);
    const int r = buildTree(right, halves.second, depth + 1);
    nodes[id].skip = (uint32_t)nodes.size();
    // Pre-order cannot tell "only a left child" from "only a right child" (traversal does not
    // care: the surviving subtree simply follows). Keep the fact for Scene::dump(), in a field
    // inner nodes do not otherwise use.
    if (l < 0 && r >= 0) nodes[id].first_ref = 0x80000000u;
    return id;
}

KDTree::Stats KDTree::stats() const {
    Stats s{nodes.size(), 0, 0, leaf_refs.size(), 0, max_depth_};
    for (const RtKdNode &n : nodes)
        if (n.is_leaf) {
            ++s.leaves;
            if (n.n_refs == 0) ++s.empty_leaves;
            s.max_leaf = std::max<size_t>(s.max_leaf, n.n_refs);
        }
    return s;
}
