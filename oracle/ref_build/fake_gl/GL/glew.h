/* TEST INFRASTRUCTURE: just enough of <GL/glew.h> for the reference's renderer.c / controls.c and the headers they
 * include to COMPILE headless (oracle/ref_build/render_stubs.c has the no-op bodies).  Nothing is drawn. */
#ifndef FAKE_GLEW_H
#define FAKE_GLEW_H
#include <limits.h>     /* renderer.c:335 uses SHRT_MAX without including it */
typedef unsigned int GLuint, GLenum, GLbitfield;
typedef int GLint, GLsizei;
typedef float GLfloat, GLclampf;
typedef unsigned char GLubyte, GLboolean;
typedef char GLchar;
typedef void GLvoid;
#define GL_COLOR_BUFFER_BIT 0x4000
static inline void glClearColor(float r, float g, float b, float a) { (void)r; (void)g; (void)b; (void)a; }
static inline void glClear(GLbitfield m) { (void)m; }
#endif
