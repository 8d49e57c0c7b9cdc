// TEST INFRASTRUCTURE (oracle) — software matrix stack behind oracle/gl_stub/GL/*.h.
//
// Semantics: two current matrices (modelview, projection), float, column-major.
// glTranslatef / glMultMatrix{f,d} post-multiply the current matrix (C = C * M), in float,
// accumulating each element left to right as  c0*m0 + c1*m1 + c2*m2 + c3*m3.
// gluPerspective follows the GLU definition: f = cot(fovy/2) in double, matrix elements
// rounded to float when multiplied in. Depth range is the GL default {0, 1}.
// The product's host Camera (hai719-raytracing_b200/host/Camera.cpp) defines the same
// arithmetic; tests/test_host_scene.py checks the two agree bit for bit.
#include <GL/gl.h>
#include <GL/glu.h>
#include <cmath>
#include <cstring>

namespace {
struct Stack {
    float mv[16];
    float pr[16];
    int mode;
    Stack() : mode(GL_MODELVIEW) { ident(mv); ident(pr); }
    static void ident(float *m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float *cur() { return mode == GL_PROJECTION ? pr : mv; }
};
// one stack per thread so concurrent oracle scenes/cameras do not trample each other
thread_local Stack g;

void post_multiply(float *c, const float *m) {
    float r[16];
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row)
            r[col * 4 + row] = c[0 * 4 + row] * m[col * 4 + 0] + c[1 * 4 + row] * m[col * 4 + 1] +
                               c[2 * 4 + row] * m[col * 4 + 2] + c[3 * 4 + row] * m[col * 4 + 3];
    std::memcpy(c, r, sizeof r);
}
}  // namespace

extern "C" {
void glMatrixMode(GLenum mode) { g.mode = (int)mode; }
void glLoadIdentity(void) { Stack::ident(g.cur()); }
void glTranslatef(GLfloat x, GLfloat y, GLfloat z) {
    float t[16];
    Stack::ident(t);
    t[12] = x; t[13] = y; t[14] = z;
    post_multiply(g.cur(), t);
}
void glMultMatrixf(const GLfloat *m) { post_multiply(g.cur(), m); }
void glMultMatrixd(const GLdouble *m) {
    float f[16];
    for (int i = 0; i < 16; ++i) f[i] = (float)m[i];
    post_multiply(g.cur(), f);
}
void glViewport(GLint, GLint, GLsizei, GLsizei) {}
void glGetDoublev(GLenum pname, GLdouble *out) {
    if (pname == GL_MODELVIEW_MATRIX) { for (int i = 0; i < 16; ++i) out[i] = g.mv[i]; }
    else if (pname == GL_PROJECTION_MATRIX) { for (int i = 0; i < 16; ++i) out[i] = g.pr[i]; }
    else if (pname == GL_DEPTH_RANGE) { out[0] = 0.0; out[1] = 1.0; }
}
void gluPerspective(GLdouble fovy, GLdouble aspect, GLdouble zNear, GLdouble zFar) {
    const double f = 1.0 / std::tan(fovy * M_PI / 360.0);
    double m[16] = {0};
    m[0] = f / aspect;
    m[5] = f;
    m[10] = (zFar + zNear) / (zNear - zFar);
    m[11] = -1.0;
    m[14] = 2.0 * zFar * zNear / (zNear - zFar);
    glMultMatrixd(m);
}
void glMaterialfv(GLenum, GLenum, const GLfloat *) {}
void glMaterialf(GLenum, GLenum, GLfloat) {}
void glEnableClientState(GLenum) {}
void glNormalPointer(GLenum, GLsizei, const GLvoid *) {}
void glVertexPointer(GLint, GLenum, GLsizei, const GLvoid *) {}
void glDrawElements(GLenum, GLsizei, GLenum, const GLvoid *) {}
}
