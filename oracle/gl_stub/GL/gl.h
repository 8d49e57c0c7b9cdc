/* TEST INFRASTRUCTURE (oracle) — a software stand-in for the few OpenGL entry points the
 * reference's src/ touches, so its sources compile headless and unmodified.
 *
 * Only the matrix stack is functional (Camera.cpp:46-56,125-132 writes it;
 * matrixUtilities.h:34-51 reads it back with glGetDoublev). Draw calls are no-ops.
 * Matrices are float, column-major, like the fixed-function pipeline stores them;
 * glGetDoublev widens each element to double.
 *
 * Nothing under hai719-raytracing_b200/ includes this file.
 */
#ifndef ORACLE_GL_STUB_GL_H
#define ORACLE_GL_STUB_GL_H

typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef unsigned int GLuint;
typedef float GLfloat;
typedef double GLdouble;
typedef void GLvoid;

#define GL_MODELVIEW 0x1700
#define GL_PROJECTION 0x1701
#define GL_MODELVIEW_MATRIX 0x0BA6
#define GL_PROJECTION_MATRIX 0x0BA7
#define GL_DEPTH_RANGE 0x0B70
#define GL_FLOAT 0x1406
#define GL_UNSIGNED_INT 0x1405
#define GL_LINES 0x0001
#define GL_TRIANGLES 0x0004
#define GL_FRONT_AND_BACK 0x0408
#define GL_AMBIENT 0x1200
#define GL_DIFFUSE 0x1201
#define GL_SPECULAR 0x1202
#define GL_SHININESS 0x1601
#define GL_VERTEX_ARRAY 0x8074
#define GL_NORMAL_ARRAY 0x8075

#ifdef __cplusplus
extern "C" {
#endif
void glMatrixMode(GLenum mode);
void glLoadIdentity(void);
void glTranslatef(GLfloat x, GLfloat y, GLfloat z);
void glMultMatrixf(const GLfloat *m);
void glMultMatrixd(const GLdouble *m);
void glViewport(GLint x, GLint y, GLsizei w, GLsizei h);
void glGetDoublev(GLenum pname, GLdouble *out);

/* preview drawing: accepted and ignored */
void glMaterialfv(GLenum face, GLenum pname, const GLfloat *params);
void glMaterialf(GLenum face, GLenum pname, GLfloat param);
void glEnableClientState(GLenum cap);
void glNormalPointer(GLenum type, GLsizei stride, const GLvoid *ptr);
void glVertexPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *ptr);
void glDrawElements(GLenum mode, GLsizei count, GLenum type, const GLvoid *indices);
#ifdef __cplusplus
}
#endif
#endif
