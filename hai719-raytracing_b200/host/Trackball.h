// Host API — virtual trackball (src/Trackball.{h,cpp}; the classic SGI trackball by Gavin Bell):
// quaternion from two mouse positions, quaternion accumulation, rotation matrix.
#ifndef HAI719_HOST_TRACKBALL_H
#define HAI719_HOST_TRACKBALL_H
void trackball(float q[4], float p1x, float p1y, float p2x, float p2y);
void negate_quat(float *q, float *qn);
void add_quats(float *q1, float *q2, float *dest);
void build_rotmatrix(float m[4][4], float q[4]);
void axis_to_quat(float a[3], float phi, float q[4]);
#endif
