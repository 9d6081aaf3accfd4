/* TEST INFRASTRUCTURE (oracle/_ref build only).  The reference's Util/sutil.cpp is the NVIDIA OptiX SDK sample utility file
 * (GLUT/GLEW display, PPM/HDR loaders); its only two functions on the path are restated in ../../dsref_host.cpp:
 * calculateCameraVariables (sutil.cpp:501-524) and currentTime (:559). */
#pragma once
#include <optixu/optixpp_namespace.h>
namespace sutil {
void calculateCameraVariables(optix::float3 eye, optix::float3 lookat, optix::float3 up, float fov, float aspect_ratio, optix::float3& U,
                              optix::float3& V, optix::float3& W, bool fov_is_vertical);
double currentTime();
} // namespace sutil
