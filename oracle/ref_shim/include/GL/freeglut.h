/* TEST INFRASTRUCTURE (oracle/_ref build only): Camera.cpp:222-229 draws the tone-mapped frame with glDrawPixels; headless here. */
#pragma once
typedef unsigned int GLenum;
typedef void GLvoid;
#define GL_UNSIGNED_BYTE 0x1401
#define GL_RGBA 0x1908
#define GL_UNPACK_ALIGNMENT 0x0CF5
inline void glPixelStorei(GLenum, int) {}
inline void glDrawPixels(int, int, GLenum, GLenum, const GLvoid*) {}
