// Host API — Line / Ray (src/Line.h:8-20, src/Ray.h:4-9). The constructor normalises the
// direction (Line.h:15); the device ray record reproduces that second normalisation.
#ifndef HAI719_HOST_RAY_H
#define HAI719_HOST_RAY_H
#include "Vec3.h"
class Line {
    Vec3 m_origin, m_direction;
public:
    Line() {}
    Line(Vec3 const &o, Vec3 const &d) : m_origin(o), m_direction(d) { m_direction.normalize(); }
    Vec3 &origin() { return m_origin; }
    Vec3 const &origin() const { return m_origin; }
    Vec3 &direction() { return m_direction; }
    Vec3 const &direction() const { return m_direction; }
};
class Ray : public Line {
public:
    float time = 0.f;
    Ray() : Line() {}
    Ray(Vec3 const &o, Vec3 const &d, float time) : Line(o, d), time(time) {}
};
#endif
