/*
 * dsref_host.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only).  The two functions of the NVIDIA SDK sample utility file
 * (reference: Util/sutil.cpp, GLUT/GLEW-bound and therefore not compiled) that the path calls, restated; and the in-memory
 * stand-in for the EXR file Camera::saveToDisk writes.
 */
#include <chrono>

#include "include/OpenEXR/ImfOutputFile.h"
#include "include/Util/sutil.h"

/* Util/sutil.cpp:501-524 */
void sutil::calculateCameraVariables(optix::float3 eye, optix::float3 lookat, optix::float3 up, float fov, float aspect_ratio, optix::float3& U,
                                     optix::float3& V, optix::float3& W, bool fov_is_vertical)
{
    float ulen, vlen, wlen;
    W = lookat - eye; /* not normalized: it implies the focal length */
    wlen = length(W);
    U = normalize(cross(W, up));
    V = normalize(cross(U, W));
    if (fov_is_vertical) {
        vlen = wlen * tanf(0.5f * fov * M_PIf / 180.0f);
        V *= vlen;
        ulen = vlen * aspect_ratio;
        U *= ulen;
    } else {
        ulen = wlen * tanf(0.5f * fov * M_PIf / 180.0f);
        U *= ulen;
        vlen = ulen / aspect_ratio;
        V *= vlen;
    }
}

double sutil::currentTime()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

dsref::ExrImage& dsref::lastExr()
{
    static ExrImage img;
    return img;
}
