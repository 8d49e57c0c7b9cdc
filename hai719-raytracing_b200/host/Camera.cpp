#include "Camera.h"
#include <cmath>
#include <cstring>
#include "Trackball.h"

namespace {
void identity(float *m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
// c = c * m, column-major, float
void post_multiply(float *c, const float *m) {
    float r[16];
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row)
            r[col * 4 + row] = c[row] * m[col * 4] + c[4 + row] * m[col * 4 + 1] + c[8 + row] * m[col * 4 + 2] +
                               c[12 + row] * m[col * 4 + 3];
    std::memcpy(c, r, sizeof r);
}
void translate(float *c, float tx, float ty, float tz) {
    float t[16];
    identity(t);
    t[12] = tx; t[13] = ty; t[14] = tz;
    post_multiply(c, t);
}
}  // namespace

Camera::Camera()
    : fovAngle(45.0f), aspectRatio(1.0f), nearPlane(4.1f), farPlane(10000.0f), spinning(0), moving(0), beginu(0),
      beginv(0), H(1), W(1), x(0.f), y(0.f), z(0.f), _zoom(3.0f) {
    trackball(curquat, 0.0, 0.0, 0.0, 0.0);
    for (int i = 0; i < 4; ++i) lastquat[i] = curquat[i];
    identity(projection);
    identity(modelview);
}

// gluPerspective semantics: f = cot(fovy/2) in fp64, elements rounded to float, multiplied
// into an identity projection matrix.
void Camera::resize(int _W, int _H) {
    H = _H;
    W = _W;
    aspectRatio = static_cast<float>(W) / static_cast<float>(H);
    const double fovy = fovAngle, aspect = aspectRatio, zn = nearPlane, zf = farPlane;
    const double f = 1.0 / std::tan(fovy * M_PI / 360.0);
    float p[16] = {0};
    p[0] = (float)(f / aspect);
    p[5] = (float)f;
    p[10] = (float)((zf + zn) / (zn - zf));
    p[11] = -1.f;
    p[14] = (float)(2.0 * zf * zn / (zn - zf));
    identity(projection);
    post_multiply(projection, p);
}

void Camera::move(float dx, float dy, float dz) { x += dx; y += dy; z += dz; }

void Camera::beginRotate(int u, int v) { beginu = u; beginv = v; moving = 1; spinning = 0; }

void Camera::rotate(int u, int v) {
    if (!moving) return;
    trackball(lastquat, (2.0 * beginu - W) / W, (H - 2.0 * beginv) / H, (2.0 * u - W) / W, (H - 2.0 * v) / H);
    beginu = u;
    beginv = v;
    spinning = 1;
    add_quats(lastquat, curquat, curquat);
}

void Camera::endRotate() { moving = 0; }
void Camera::zoom(float dz) { _zoom += dz; }

void Camera::apply() {
    identity(modelview);
    translate(modelview, x, y, z);
    float rot[4][4];
    build_rotmatrix(rot, curquat);
    translate(modelview, 0.0f, 0.0f, -_zoom);
    post_multiply(modelview, &rot[0][0]);
}

void Camera::getPos(float &X, float &Y, float &Z) {
    float m[4][4];
    build_rotmatrix(m, curquat);
    const float px = -x, py = -y, pz = -z + _zoom;
    X = m[0][0] * px + m[0][1] * py + m[0][2] * pz;
    Y = m[1][0] * px + m[1][1] * py + m[1][2] * pz;
    Z = m[2][0] * px + m[2][1] * py + m[2][2] * pz;
}
