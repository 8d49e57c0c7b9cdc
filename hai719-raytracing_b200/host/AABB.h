// Host API — AABB and cutting plane (src/AABB.h:11-75). The ray/box slab test
// (AABB::intersects, AABB.h:48-65) runs on the device: csrc/rt_intersect.cuh : slab_hit().
#ifndef HAI719_HOST_AABB_H
#define HAI719_HOST_AABB_H
#include <cfloat>
#include <utility>
#include "Vec3.h"

struct AABBCuttingPlane {
    float position = 0.f;
    unsigned int axis = 0;
    AABBCuttingPlane() {}
    AABBCuttingPlane(unsigned int axis, float position) : position(position), axis(axis) {}
};

class AABB {
public:
    Vec3 p0, p1;
    AABB() : p0(FLT_MAX), p1(-FLT_MAX) {}
    // corners are ordered per component (AABB.h:27-37)
    AABB(const Vec3 &a, const Vec3 &b) {
        for (unsigned int i = 0; i < 3; ++i) {
            if (a[i] < b[i]) { p0[i] = a[i]; p1[i] = b[i]; }
            else             { p1[i] = a[i]; p0[i] = b[i]; }
        }
    }
    void extend(const AABB &o) {
        for (unsigned int i = 0; i < 3; ++i) {
            p0[i] = o.p0[i] < p0[i] ? o.p0[i] : p0[i];
            p1[i] = o.p1[i] > p1[i] ? o.p1[i] : p1[i];
        }
    }
    // children are copies of the box with one coordinate replaced by the plane (AABB.h:67-75)
    std::pair<AABB, AABB> split(const AABBCuttingPlane &plane) const {
        std::pair<AABB, AABB> r(*this, *this);
        r.first.p1[plane.axis] = plane.position;
        r.second.p0[plane.axis] = plane.position;
        return r;
    }
};
#endif
