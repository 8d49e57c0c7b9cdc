// Host API — Sphere (src/Sphere.h:41-47): a centre, a radius and the Mesh base it inherits its
// Material from. Sphere::intersect (Sphere.h:91-132) lives on the device
// (csrc/rt_intersect.cuh : sphere_test / sphere_finish); build_arrays() only fed the GL preview.
#ifndef HAI719_HOST_SPHERE_H
#define HAI719_HOST_SPHERE_H
#include "Mesh.h"
class Sphere : public Mesh {
public:
    Vec3 m_center;
    float m_radius = 0.f;
    Sphere() : Mesh() {}
    Sphere(Vec3 c, float r) : Mesh(), m_center(c), m_radius(r) {}
    void build_arrays() override {}
};
#endif
