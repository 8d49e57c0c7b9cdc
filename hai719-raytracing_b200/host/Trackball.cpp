#include "Trackball.h"
#include <cmath>

namespace {
const float kBallSize = 0.8f;      // TRACKBALLSIZE
const int kRenormEvery = 97;       // RENORMCOUNT

float len3(const float *v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
void cross3(const float *a, const float *b, float *out) {
    const float t[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
}
// Project (x, y) onto a sphere of radius r, or onto a hyperbolic sheet away from the centre.
float project_to_sphere(float r, float x, float y) {
    const float d = std::sqrt(x * x + y * y);
    if (d < r * 0.70710678118654752440) return std::sqrt(r * r - d * d);
    const float t = r / 1.41421356237309504880;
    return t * t / d;
}
void normalize_quat(float q[4]) {
    const float mag = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    for (int i = 0; i < 4; ++i) q[i] /= mag;
}
}  // namespace

void axis_to_quat(float a[3], float phi, float q[4]) {
    const float l = len3(a);
    for (int i = 0; i < 3; ++i) q[i] = a[i] * (1.0f / l);
    const float s = std::sin(phi / 2.0);
    for (int i = 0; i < 3; ++i) q[i] *= s;
    q[3] = std::cos(phi / 2.0);
}

void trackball(float q[4], float p1x, float p1y, float p2x, float p2y) {
    if (p1x == p2x && p1y == p2y) {  // no motion: identity rotation
        q[0] = q[1] = q[2] = 0.f;
        q[3] = 1.f;
        return;
    }
    float p1[3] = {p1x, p1y, project_to_sphere(kBallSize, p1x, p1y)};
    float p2[3] = {p2x, p2y, project_to_sphere(kBallSize, p2x, p2y)};
    float axis[3];
    cross3(p2, p1, axis);
    const float d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    float t = len3(d) / (2.0f * kBallSize);
    if (t > 1.0f) t = 1.0f;
    if (t < -1.0f) t = -1.0f;
    const float phi = 2.0 * std::asin(t);
    axis_to_quat(axis, phi, q);
}

void negate_quat(float *q, float *qn) {
    qn[0] = -q[0]; qn[1] = -q[1]; qn[2] = -q[2]; qn[3] = q[3];
}

void add_quats(float *q1, float *q2, float *dest) {
    static int count = 0;
    float t1[3] = {q1[0] * q2[3], q1[1] * q2[3], q1[2] * q2[3]};
    float t2[3] = {q2[0] * q1[3], q2[1] * q1[3], q2[2] * q1[3]};
    float t3[3];
    cross3(q2, q1, t3);
    float tf[4];
    for (int i = 0; i < 3; ++i) tf[i] = t1[i] + t2[i] + t3[i];
    tf[3] = q1[3] * q2[3] - (q1[0] * q2[0] + q1[1] * q2[1] + q1[2] * q2[2]);
    for (int i = 0; i < 4; ++i) dest[i] = tf[i];
    if (++count > kRenormEvery) {
        count = 0;
        normalize_quat(dest);
    }
}

// Trackball.cpp:323-345: elements evaluated in fp64, stored as float.
void build_rotmatrix(float m[4][4], float q[4]) {
    m[0][0] = 1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]);
    m[0][1] = 2.0 * (q[0] * q[1] - q[2] * q[3]);
    m[0][2] = 2.0 * (q[2] * q[0] + q[1] * q[3]);
    m[1][0] = 2.0 * (q[0] * q[1] + q[2] * q[3]);
    m[1][1] = 1.0 - 2.0 * (q[2] * q[2] + q[0] * q[0]);
    m[1][2] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
    m[2][0] = 2.0 * (q[2] * q[0] - q[1] * q[3]);
    m[2][1] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
    m[2][2] = 1.0 - 2.0 * (q[1] * q[1] + q[0] * q[0]);
    m[0][3] = m[1][3] = m[2][3] = 0.f;
    m[3][0] = m[3][1] = m[3][2] = 0.f;
    m[3][3] = 1.f;
}
