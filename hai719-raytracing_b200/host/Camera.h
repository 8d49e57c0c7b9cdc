// Host API — Camera (src/Camera.{h,cpp}) with the OpenGL matrix stack replaced by two member
// matrices. Same public methods; apply()/resize() compute, in software and in float, what the
// reference pushes through glLoadIdentity/glTranslatef/glMultMatrixf/gluPerspective
// (Camera.cpp:46-56,125-132) and MatrixUtilities reads back with glGetDoublev:
//     projection = perspective(fov 45, W/H, near 4.1, far 10000)         (Camera.cpp:24-38)
//     modelview  = I * T(x,y,z) * T(0,0,-zoom) * R(curquat)
// Each post-multiplication C*M is done in float, element = c0*m0 + c1*m1 + c2*m2 + c3*m3 left to
// right — the arithmetic oracle/gl_stub.cpp defines for the reference build; the two are
// compared bit for bit in tests/test_host_scene.py.
#ifndef HAI719_HOST_CAMERA_H
#define HAI719_HOST_CAMERA_H
#include "Vec3.h"

class Camera {
public:
    Camera();
    virtual ~Camera() {}

    float getFovAngle() const { return fovAngle; }
    void setFovAngle(float v) { fovAngle = v; }
    float getAspectRatio() const { return aspectRatio; }
    float getNearPlane() const { return nearPlane; }
    void setNearPlane(float v) { nearPlane = v; }
    float getFarPlane() const { return farPlane; }
    void setFarPlane(float v) { farPlane = v; }
    unsigned int getScreenWidth() const { return W; }
    unsigned int getScreenHeight() const { return H; }

    void resize(int W, int H);
    void move(float dx, float dy, float dz);
    void beginRotate(int u, int v);
    void rotate(int u, int v);
    void endRotate();
    void zoom(float z);
    void apply();
    void getPos(float &x, float &y, float &z);
    void getPos(Vec3 &p) { getPos(p[0], p[1], p[2]); }

    // column-major, what glGetDoublev(GL_PROJECTION_MATRIX / GL_MODELVIEW_MATRIX) would return
    const float *projectionMatrix() const { return projection; }
    const float *modelviewMatrix() const { return modelview; }

private:
    float fovAngle, aspectRatio, nearPlane, farPlane;
    int spinning, moving, beginu, beginv;
    int H, W;
    float curquat[4], lastquat[4];
    float x, y, z, _zoom;
    float projection[16], modelview[16];
};
#endif
