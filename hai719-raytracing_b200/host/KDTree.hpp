// Host API — KDTree (src/KDTree.hpp, src/KDTree.cpp). The BUILD stays on the host in C++, as
// BASELINE.json's north_star asks, and reproduces the reference's tree exactly:
//   * split axis = depth % 3; split position = median of the triangles' box minima on that axis,
//     plus EPSILON added in fp64 and rounded to float            (KDTree.cpp:87-98)
//   * a triangle goes left if its box max <= pos - EPSILON, right if its box min >= pos + EPSILON
//     (both compared in fp64), otherwise to BOTH sides           (KDTree.cpp:130-140)
//   * <= 40 triangles -> leaf; equal-sized halves -> leaf holding ALL triangles (KDTree.cpp:142-145)
//   * depth > 100 -> no node at all, the triangles are dropped   (KDTree.cpp:101-103)
//   * child boxes = parent box with one coordinate replaced      (AABB.h:67-75)
// What differs is the representation: nodes are emitted straight into a flat PRE-ORDER array
// (RtKdNode, include/hai719_rt.h) with skip links, and leaf triangle lists into one contiguous
// RtTriRef array in build order, so flatten() is a memcpy and the device needs no stack.
// Traversal (KDTree::intersect / Node::intersect, KDTree.cpp:31-85) is device code:
// csrc/rt_intersect.cuh : mesh_closest().
#ifndef HAI719_HOST_KDTREE_HPP
#define HAI719_HOST_KDTREE_HPP
#include <vector>
#include "AABB.h"
#include "Constants.h"
#include "Mesh.h"
#include "hai719_rt.h"

class KDTree {
public:
    const std::vector<MeshVertex> &vertices;
    AABB aabb;
    int root;                        // 0, or -1 for an empty tree (the reference's nullptr)
    std::vector<RtKdNode> nodes;     // pre-order
    std::vector<RtTriRef> leaf_refs; // leaf triangle lists, back to back

    KDTree(const std::vector<MeshTriangle> &triangles, const AABB &aabb, const std::vector<MeshVertex> &vertices);

    struct Stats { size_t nodes, leaves, empty_leaves, refs, max_leaf; unsigned int max_depth; };
    Stats stats() const;

private:
    int buildTree(const std::vector<MeshTriangle> &triangles, const AABB &box, unsigned int depth);
    AABBCuttingPlane cut(const std::vector<MeshTriangle> &triangles, int depth) const;
    int emit_leaf(const std::vector<MeshTriangle> &triangles, const AABB &box);
    unsigned int max_depth_ = 0;
};
#endif
