#include "Mesh.h"
#include <cmath>
#include <fstream>
#include <sstream>
#include "KDTree.hpp"
#include "errors.h"

// OFF / COFF reader (behaviour of Mesh.cpp:9-74):
//   header "OFF" or "COFF", then nV nT nE; COFF vertex lines are  x y z r g b a  (colours / 255);
//   the first face line decides whether faces carry colours (anything after "3 i j k");
//   every triangle remembers its own index in v[3].
// Numbers are read with the iostream extractors, as the reference does, so every float parses
// to the same bits.
// Hardened: a file the reference would read out of bounds on, or allocate for blindly, is an error with
// the file name here - wrong magic, counts that cannot fit in the file, a vertex or face list that ends early, a face
// that names a vertex the file does not have. Every well-formed file parses exactly as before.
void Mesh::loadOFF(const std::string &filename) {
    std::ifstream in(filename.c_str());
    if (!in) hai719::fatal("cannot open mesh file: " + filename);  // reference: exit(EXIT_FAILURE)
    in.seekg(0, std::ios::end);
    const long long file_bytes = (long long)in.tellg();
    in.seekg(0, std::ios::beg);

    std::string magic;
    long long nV64 = 0, nT64 = 0, nE64 = 0;
    in >> magic >> nV64 >> nT64 >> nE64;
    if (magic != "OFF" && magic != "COFF") hai719::fatal("not an OFF file (magic '" + magic + "'): " + filename);
    if (!in) hai719::fatal("unreadable OFF header: " + filename);
    // a vertex line takes at least 6 bytes ("0 0 0\n"), a face line at least 8 ("3 0 1 2\n")
    if (nV64 < 0 || nT64 < 0 || nV64 > file_bytes / 6 || nT64 > file_bytes / 8)
        hai719::fatal("OFF counts (" + std::to_string(nV64) + " vertices, " + std::to_string(nT64) + " faces) do not fit in " +
                      std::to_string(file_bytes) + " bytes: " + filename);
    const unsigned int nV = (unsigned int)nV64, nT = (unsigned int)nT64;
    unsigned int tmp = 0;
    vertices.assign(nV, MeshVertex());
    triangles.assign(nT, MeshTriangle());
    vertColors.clear();
    faceColors.clear();
    colorType = (magic == "COFF") ? ColorType_Vertex : ColorType_None;

    if (colorType == ColorType_Vertex) {
        vertColors.resize(nV);
        for (unsigned int i = 0; i < nV; ++i) {
            in >> vertices[i].position >> vertColors[i] >> tmp;
            vertColors[i] /= 255.0;
        }
    } else {
        for (unsigned int i = 0; i < nV; ++i) in >> vertices[i].position;
    }
    if (!in) hai719::fatal("OFF vertex list ends early or is not numeric: " + filename);

    std::string line;
    std::getline(in, line);  // rest of the last vertex line
    for (unsigned int i = 0; i < nT; ++i) {
        std::getline(in, line);
        std::istringstream ls(line);
        int arity;
        ls >> arity;
        long long idx[3] = {0, 0, 0};
        for (unsigned int j = 0; j < 3; ++j) ls >> idx[j];
        if (!ls) hai719::fatal("OFF face " + std::to_string(i) + " of " + std::to_string(nT) + " is missing or incomplete: " + filename);
        for (unsigned int j = 0; j < 3; ++j) {
            if (idx[j] < 0 || idx[j] >= (long long)nV)
                hai719::fatal("OFF face " + std::to_string(i) + " names vertex " + std::to_string(idx[j]) + " of " + std::to_string(nV) + ": " + filename);
            triangles[i].v[j] = (unsigned int)idx[j];
        }
        if (i == 0 && !(ls >> std::ws).eof()) {
            colorType = ColorType_Face;
            faceColors.resize(nT);
        }
        if (colorType == ColorType_Face) {
            ls >> faceColors[i][0] >> faceColors[i][1] >> faceColors[i][2];
            faceColors[i] /= 255.0f;
        }
        triangles[i].v[3] = i;
    }
}

void Mesh::recomputeNormals() {
    for (MeshVertex &v : vertices) v.normal = Vec3(0.f, 0.f, 0.f);
    for (const MeshTriangle &t : triangles) {
        Vec3 n = Vec3::cross(vertices[t.v[1]].position - vertices[t.v[0]].position,
                             vertices[t.v[2]].position - vertices[t.v[0]].position);
        n.normalize();
        for (unsigned int j = 0; j < 3; ++j) vertices[t.v[j]].normal += n;
    }
    for (MeshVertex &v : vertices) v.normal.normalize();
}

void Mesh::centerAndScaleToUnit() {
    Vec3 c(0, 0, 0);
    for (const MeshVertex &v : vertices) c += v.position;
    c /= vertices.size();
    float maxD = (vertices[0].position - c).length();
    for (const MeshVertex &v : vertices) {
        const float m = (v.position - c).length();
        if (m > maxD) maxD = m;
    }
    for (MeshVertex &v : vertices) v.position = (v.position - c) / maxD;
}

// Mesh.h:143-156 — note the upper corner starts at FLT_MIN (the smallest POSITIVE float), so the
// box of a mesh lying entirely at negative coordinates reaches up to ~0.
void Mesh::computeAABB() {
    Vec3 lo(FLT_MAX), hi(FLT_MIN);
    for (const MeshVertex &v : vertices)
        for (unsigned int a = 0; a < 3; ++a) {
            if (v.position[a] < lo[a]) lo[a] = v.position[a];
            if (v.position[a] > hi[a]) hi[a] = v.position[a];
        }
    lo -= Vec3(EPSILON);
    hi += Vec3(EPSILON);
    aabb = AABB(lo, hi);
}

void Mesh::computeKDTree() {
    computeAABB();
    kdtree = std::make_shared<KDTree>(triangles, aabb, vertices);
}

void Mesh::translate(Vec3 const &t) { for (MeshVertex &v : vertices) v.position += t; }
void Mesh::apply_transformation_matrix(Mat3 m) { for (MeshVertex &v : vertices) v.position = m * v.position; }
void Mesh::scale(Vec3 const &s) { apply_transformation_matrix(Mat3(s[0], 0., 0., 0., s[1], 0., 0., 0., s[2])); }

// angle is in degrees; converted in fp64 then stored as float; cos/sin are the fp64 functions
// applied to that float (Mesh.h:198-224, as bound in the reference's translation units).
void Mesh::rotate_x(float angle) {
    const float a = angle * M_PI / 180.;
    const double c = std::cos((double)a), s = std::sin((double)a);
    apply_transformation_matrix(Mat3(1., 0., 0., 0., c, -s, 0., s, c));
}
void Mesh::rotate_y(float angle) {
    const float a = angle * M_PI / 180.;
    const double c = std::cos((double)a), s = std::sin((double)a);
    apply_transformation_matrix(Mat3(c, 0., s, 0., 1., 0., -s, 0., c));
}
void Mesh::rotate_z(float angle) {
    const float a = angle * M_PI / 180.;
    const double c = std::cos((double)a), s = std::sin((double)a);
    apply_transformation_matrix(Mat3(c, -s, 0., s, c, 0., 0., 0., 1.));
}
