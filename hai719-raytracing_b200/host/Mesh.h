// Host API — Mesh, MeshVertex, MeshTriangle (src/Mesh.h:31-285, src/Mesh.cpp) without OpenGL.
//
// Kept: the data members scene code touches (vertices, triangles, vertColors, faceColors,
// colorType, aabb, kdtree, material), the OFF/COFF loader, the in-place transforms (which move
// vertices[].position only, Mesh.h:173-224), normals, AABB, KD-tree construction.
// Not here: draw() and the GL client arrays (preview only), and intersect()/intersectOld() —
// ray/mesh intersection IS the hot path and exists only as CUDA (csrc/rt_intersect.cuh).
#ifndef HAI719_HOST_MESH_H
#define HAI719_HOST_MESH_H
#include <cfloat>
#include <memory>
#include <string>
#include <vector>
#include "AABB.h"
#include "Constants.h"
#include "Material.h"
#include "Vec3.h"

class KDTree;

struct MeshVertex {
    MeshVertex() {}
    MeshVertex(const Vec3 &p, const Vec3 &n) : position(p), normal(n) {}
    Vec3 position;
    Vec3 normal;
    float u = 0.f, v = 0.f;
};

struct MeshTriangle {
    MeshTriangle() : v{0, 0, 0, 0} {}
    MeshTriangle(unsigned int v0, unsigned int v1, unsigned int v2) : v{v0, v1, v2, 0} {}
    unsigned int &operator[](unsigned int i) { return v[i]; }
    unsigned int operator[](unsigned int i) const { return v[i]; }
    unsigned int v[4];  // three vertex indices + the triangle's own index (Mesh.cpp:71-73)
};

enum ColorType { ColorType_Vertex, ColorType_Face, ColorType_None };

class Mesh {
public:
    std::vector<MeshVertex> vertices;
    std::vector<MeshTriangle> triangles;
    std::vector<Vec3> vertColors;
    std::vector<Vec3> faceColors;
    ColorType colorType = ColorType_None;
    AABB aabb;
    std::shared_ptr<KDTree> kdtree;  // null until computeKDTree(): the device then brute-forces
                                     // behind the mesh box, like Mesh::intersectOld (Mesh.h:257-277)
    Material material;

    virtual ~Mesh() {}

    void loadOFF(const std::string &filename);
    void recomputeNormals();
    void centerAndScaleToUnit();
    void computeKDTree();
    void computeAABB();
    virtual void build_arrays() { recomputeNormals(); computeAABB(); }

    void translate(Vec3 const &translation);
    void apply_transformation_matrix(Mat3 transform);
    void scale(Vec3 const &scale);
    void rotate(Vec3 const &angles) { rotate_x(angles[0]); rotate_y(angles[1]); rotate_z(angles[2]); }
    void rotate_x(float angle);
    void rotate_y(float angle);
    void rotate_z(float angle);
};
#endif
