/* TEST INFRASTRUCTURE (oracle) — see gl.h. Only gluPerspective is needed (Camera.cpp:53). */
#ifndef ORACLE_GL_STUB_GLU_H
#define ORACLE_GL_STUB_GLU_H
#include "gl.h"
#ifdef __cplusplus
extern "C" {
#endif
void gluPerspective(GLdouble fovy, GLdouble aspect, GLdouble zNear, GLdouble zFar);
#ifdef __cplusplus
}
#endif
#endif
