// Host API — MatrixUtilities (src/matrixUtilities.h) reading the Camera's software matrices
// instead of the GL state. Holds the fp64 inverses that ray generation needs; the per-sample
// unprojection itself (screen_space_to_world_space_ray, matrixUtilities.h:53-74) runs on the
// device (csrc/rt_kernels.cu : camera_ray) with the same fp64 operation order.
#ifndef HAI719_HOST_MATRIXUTILITIES_H
#define HAI719_HOST_MATRIXUTILITIES_H
#include "Camera.h"
#include "hai719_rt.h"

// 4x4 inverse by cofactors, column-major, in T (the classic MESA gluInvertMatrix expansion:
// each cofactor is a sum of six triple products, added in the order written; det from the
// first row of cofactors; result = cofactor * (1/det)). Returns false if singular.
template <class T> bool gluInvertMatrix(const T m[16], T invOut[16]);

class MatrixUtilities {
public:
    double modelview[16];
    double modelviewInverse[16];
    double projection[16];
    double projectionInverse[16];
    double nearAndFarPlanes[2];

    MatrixUtilities() {
        for (int i = 0; i < 16; ++i) modelview[i] = modelviewInverse[i] = projection[i] = projectionInverse[i] = (i % 5 == 0);
        nearAndFarPlanes[0] = 0.0;
        nearAndFarPlanes[1] = 1.0;
    }
    // matrixUtilities.h:34-51 with glGetDoublev replaced by the camera's members
    void updateMatrices(const Camera &camera) {
        for (int i = 0; i < 16; ++i) { modelview[i] = camera.modelviewMatrix()[i]; projection[i] = camera.projectionMatrix()[i]; }
        gluInvertMatrix(modelview, modelviewInverse);
        gluInvertMatrix(projection, projectionInverse);
        nearAndFarPlanes[0] = 0.0;  // GL_DEPTH_RANGE default
        nearAndFarPlanes[1] = 1.0;
    }
    void fill(RtCamera &out) const {
        for (int i = 0; i < 16; ++i) { out.modelview_inverse[i] = modelviewInverse[i]; out.projection_inverse[i] = projectionInverse[i]; }
        out.depth_near = nearAndFarPlanes[0];
    }
};

template <class T> bool gluInvertMatrix(const T m[16], T invOut[16]) {
    T inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    T det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0) return false;
    det = 1.0 / det;
    for (int i = 0; i < 16; ++i) invOut[i] = inv[i] * det;
    return true;
}
#endif
