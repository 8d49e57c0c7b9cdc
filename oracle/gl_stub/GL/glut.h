/* TEST INFRASTRUCTURE (oracle) — see gl.h. The reference's src/ headers include <GL/glut.h>
 * only to reach the GL types; no glut entry point is used off main.cpp. */
#ifndef ORACLE_GL_STUB_GLUT_H
#define ORACLE_GL_STUB_GLUT_H
#include "gl.h"
#include "glu.h"
#endif
