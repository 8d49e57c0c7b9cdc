// rt_pack.hpp — host-side packing of the per-mesh KD arrays of an RtSceneDesc into the shared node
// array the device walks (layout documented at rt::DMesh). Pure bit moves plus index offsets; used by
// rt_scene_create (rt_capi.cu) and by the test-only CPU harness (tests/hostsim), so both see the
// same layout. Include after rt_core.cuh.
#ifndef HAI719_RT_PACK_HPP
#define HAI719_RT_PACK_HPP
#include <cstdint>
#include <string>
#include <vector>
#include "hai719_rt.h"

namespace rt {

struct PackedMeshes {
    std::vector<float4> lo, hi;                  // all nodes
    std::vector<uint32_t> node_begin, node_end;  // per mesh
    std::vector<uint32_t> ref_begin;             // per mesh: first slot of its leaf refs in the shared triangle arrays
    uint32_t total_refs = 0;
};

// Returns "" on success, else the reason the description is inconsistent.
inline std::string pack_meshes(const RtSceneDesc &d, PackedMeshes &out) {
    out = PackedMeshes();
    for (uint32_t mi = 0; mi < d.n_meshes; ++mi) {
        const RtSceneMesh &m = d.meshes[mi];
        if ((m.n_vertices && !m.positions) || (m.n_triangles && !m.triangles) || (m.n_nodes && !m.nodes) || (m.n_leaf_refs && !m.leaf_refs))
            return "mesh with a null array";
        if (m.color_type == RT_COLOR_VERTEX && !m.vert_colors) return "vertex-coloured mesh without vert_colors";
        if (m.color_type == RT_COLOR_FACE && !m.face_colors) return "face-coloured mesh without face_colors";
        const uint32_t ref_base = out.total_refs;
        out.ref_begin.push_back(ref_base);
        const uint32_t begin = (uint32_t)out.lo.size();
        out.node_begin.push_back(begin);
        if (m.n_nodes) {
            const uint32_t base = begin + 1;  // the tree starts after the synthetic root
            out.lo.push_back(make_float4(m.root_bmin[0], m.root_bmin[1], m.root_bmin[2], u2f(base + m.n_nodes)));
            out.hi.push_back(make_float4(m.root_bmax[0], m.root_bmax[1], m.root_bmax[2], u2f(0u)));
            for (uint32_t k = 0; k < m.n_nodes; ++k) {
                const RtKdNode &n = m.nodes[k];
                if (n.is_leaf) {
                    if ((uint64_t)n.first_ref + n.n_refs > m.n_leaf_refs || n.n_refs >= 0x80000000u) return "KD leaf range out of bounds";
                    out.lo.push_back(make_float4(n.bmin[0], n.bmin[1], n.bmin[2], u2f(ref_base + n.first_ref)));
                    out.hi.push_back(make_float4(n.bmax[0], n.bmax[1], n.bmax[2], u2f(0x80000000u | n.n_refs)));
                } else {
                    if (n.skip <= k || n.skip > m.n_nodes) return "KD skip link out of bounds";
                    out.lo.push_back(make_float4(n.bmin[0], n.bmin[1], n.bmin[2], u2f(base + n.skip)));
                    out.hi.push_back(make_float4(n.bmax[0], n.bmax[1], n.bmax[2], u2f(0u)));
                }
            }
        }
        out.node_end.push_back((uint32_t)out.lo.size());
        for (uint32_t k = 0; k < m.n_leaf_refs; ++k) {
            const RtTriRef &r = m.leaf_refs[k];
            if (r.v[0] >= m.n_vertices || r.v[1] >= m.n_vertices || r.v[2] >= m.n_vertices) return "leaf ref vertex index out of bounds";
            if (r.tri_index >= m.n_triangles) return "leaf ref triangle index out of bounds";
        }
        // the shading code follows triangles[] into the per-vertex arrays (vertex colours)
        for (uint64_t k = 0; k < (uint64_t)3 * m.n_triangles; ++k)
            if (m.triangles[k] >= m.n_vertices) return "triangle vertex index out of bounds";
        if ((uint64_t)out.total_refs + m.n_leaf_refs >= 0x7FFFFFFFull) return "too many leaf refs";
        out.total_refs += m.n_leaf_refs;
    }
    return "";
}

}  // namespace rt
#endif
