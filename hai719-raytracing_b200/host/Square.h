// Host API — Square (src/Square.h:20-63). setQuad() fills the four vertices and the
// m_normal / m_right_vector / m_up_vector members exactly as the reference does; transforms
// then move the vertices only, so after a rotate the members are stale — and the reference
// keeps using the stale m_right_vector / m_up_vector as the normal-map tangent frame
// (Scene.h:284). flatten() ships both. Square::intersect (Square.h:65-126) lives on
// the device (csrc/rt_intersect.cuh : square_test).
#ifndef HAI719_HOST_SQUARE_H
#define HAI719_HOST_SQUARE_H
#include "Mesh.h"
class Square : public Mesh {
public:
    Vec3 m_normal, m_bottom_left, m_right_vector, m_up_vector;

    Square() : Mesh() {}
    Square(Vec3 const &bottomLeft, Vec3 const &rightVector, Vec3 const &upVector, float width = 1., float height = 1.,
           float uMin = 0.f, float uMax = 1.f, float vMin = 0.f, float vMax = 1.f) : Mesh() {
        setQuad(bottomLeft, rightVector, upVector, width, height, uMin, uMax, vMin, vMax);
    }

    void setQuad(Vec3 const &bottomLeft, Vec3 const &rightVector, Vec3 const &upVector, float width = 1., float height = 1.,
                 float uMin = 0.f, float uMax = 1.f, float vMin = 0.f, float vMax = 1.f) {
        m_bottom_left = bottomLeft;
        m_normal = Vec3::cross(rightVector, upVector);
        m_normal.normalize();
        m_right_vector = rightVector;
        m_right_vector.normalize();
        m_right_vector = m_right_vector * width;
        m_up_vector = upVector;
        m_up_vector.normalize();
        m_up_vector = m_up_vector * height;

        const Vec3 corner[4] = {bottomLeft, bottomLeft + m_right_vector, bottomLeft + m_right_vector + m_up_vector,
                                bottomLeft + m_up_vector};
        const float us[4] = {uMin, uMax, uMax, uMin}, vs[4] = {vMin, vMin, vMax, vMax};
        vertices.assign(4, MeshVertex());
        for (int k = 0; k < 4; ++k) {
            vertices[k].position = corner[k];
            vertices[k].normal = m_normal;
            vertices[k].u = us[k];
            vertices[k].v = vs[k];
        }
        triangles.assign(2, MeshTriangle());
        triangles[0] = MeshTriangle(0, 1, 2);
        triangles[1] = MeshTriangle(0, 2, 3);
    }
};
#endif
