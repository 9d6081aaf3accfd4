/*
 * dsref_device.h -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * What <optix.h>, <optix_device.h>, <optix_world.h> and the CUDA device headers give an OptiX 5.1 program, re-expressed for
 * g++ so that one reference .cu file compiles, unmodified, as a host translation unit ("module").  Include order inside a
 * module wrapper (oracle/ref_shim/modules/*.cpp):
 *     #define DSREF_MODULE_NAME "cloud....cu"
 *     #include "dsref_device.h"
 *     #include "CUDA/cloud....cu"            <- straight from /root/reference
 *     DSREF_PROGRAM(...) DSREF_BUFFER(...) DSREF_SAMPLER(...)
 *
 * Substitutions the oracle header (oracle/ds_oracle.cpp) documents, applied here by macro so the reference text stays as is:
 *   clock()            -> dsref::streamId()         (random.cuh:38; explicit RNG stream instead of the SM clock)
 *   expf, log, cos, sin, log2f, powf(2, x) -> include/ds_detmath.h   (the reference builds with --use_fast_math; both the
 *                         oracle and this library evaluate the same deterministic fp32 kernels instead)
 * Everything else (sqrt, floor, fabs, fminf, acos, tanf, powf with another base) is libm / IEEE.
 */
#ifndef DSREF_DEVICE_H
#define DSREF_DEVICE_H

/* every std header a reference file may include, pulled in BEFORE the math macros below */
#include <array>
#include <cassert>
#include <cfloat>
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ds_detmath.h"
#include "dsref_gsl.h"
#include "dsref_runtime.h"

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define RT_HOSTDEVICE
#define RT_PROGRAM static
#define cudaReadModeNormalizedFloat 1
#define cudaReadModeElementType 0

using optix::size_t2;
using optix::size_t3;
using std::max;
using std::min;

#ifndef DSREF_MODULE_NAME
#error "define DSREF_MODULE_NAME before including dsref_device.h"
#endif
static dsref::Module* const dsref_this_module = dsref::registerModule(DSREF_MODULE_NAME);

/* rtDeclareVariable(type, name, semantic, annotation): a file-static object the launcher fills by name */
#define rtDeclareVariable(type, name, semantic, annotation)                                                            \
    static type name;                                                                                                  \
    static const int name##_dsref_registered = dsref::addVar(dsref_this_module, #name, #semantic, (void*)&name, sizeof(type))

namespace dsref {
template <int N> struct LaunchSize;
template <> struct LaunchSize<1> { typedef size_t type; };
template <> struct LaunchSize<2> { typedef optix::size_t2 type; };
template <> struct LaunchSize<3> { typedef optix::size_t3 type; };

template <typename T, int N = 1>
struct DevBuffer : DevBufferBase {
    typename LaunchSize<N>::type size() const;
    T& operator[](size_t i) { return ((T*)data)[i]; }
    T& operator[](const uint2& i) { return ((T*)data)[(size_t)i.y * dim[0] + i.x]; }
    T& operator[](const optix::size_t2& i) { return ((T*)data)[i.y * dim[0] + i.x]; }
    T& operator[](const uint3& i) { return ((T*)data)[((size_t)i.z * dim[1] + i.y) * dim[0] + i.x]; }
};
template <typename T, int N> inline typename LaunchSize<N>::type DevBuffer<T, N>::size() const
{
    if constexpr (N == 1) {
        return dim[0];
    } else if constexpr (N == 2) {
        return optix::size_t2{dim[0], dim[1]};
    } else {
        return optix::size_t3{dim[0], dim[1], dim[2]};
    }
}

template <typename T, int N, int Mode>
struct DevTex : DevTexBase {
};
} // namespace dsref

#define rtBuffer static dsref::DevBuffer
#define rtTextureSampler static dsref::DevTex

template <typename T, int M> inline float tex1D(const dsref::DevTex<T, 1, M>& t, float u) { return dsref::fetch1D(t.state, u); }
template <typename T, int M> inline float tex3D(const dsref::DevTex<T, 3, M>& t, float u, float v, float w)
{
    dsref::counters().boundTaps++;
    return dsref::fetch3D(t.state, u, v, w);
}
template <typename R> inline R rtTex3D(int id, float u, float v, float w)
{
    dsref::counters().bindlessTaps++;
    return dsref::fetch3D(dsref::samplerById(id), u, v, w);
}
template <typename R> inline R rtTex3DLod(int id, float u, float v, float w, float lod)
{
    return dsref::fetch3DLod(dsref::samplerById(id), u, v, w, lod);
}

template <typename PRD> inline void rtTrace(rtObject, const optix::Ray& ray, PRD& prd) { dsref::traceRay(ray, &prd, sizeof(PRD)); }
inline bool rtPotentialIntersection(float t) { return dsref::potentialIntersection(t); }
inline bool rtReportIntersection(unsigned int material) { return dsref::reportIntersection(material); }

/* CUDA device intrinsics the included headers mention */
inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

/* ---- deterministic math substitutions (see header) ---- */
inline float dsref_expf(float x) { return ds_expf(x); }
inline float dsref_log(float x) { return ds_logf(x); }
inline double dsref_log(double x) { return ::log(x); }
inline float dsref_cos(float x)
{
    float s, c;
    ds_sincosf(x, &s, &c);
    return c;
}
inline double dsref_cos(double x) { return ::cos(x); }
inline float dsref_sin(float x)
{
    float s, c;
    ds_sincosf(x, &s, &c);
    return s;
}
inline double dsref_sin(double x) { return ::sin(x); }
inline float dsref_log2f(float x) { return ds_log2f(x); }
inline float dsref_powf(float base, float x) { return base == 2.0f ? ds_exp2f(x) : ::powf(base, x); }

#define clock() dsref::streamId()
#define expf dsref_expf
#define log dsref_log
#define cos dsref_cos
#define sin dsref_sin
#define log2f dsref_log2f
#define powf dsref_powf

/* registration helpers for the wrapper that includes the .cu */
#define DSREF_CAT2(a, b) a##b
#define DSREF_CAT(a, b) DSREF_CAT2(a, b)
#define DSREF_PROGRAM(fn) static const int DSREF_CAT(dsref_prog_, __LINE__) = dsref::addProg(dsref_this_module, #fn, &fn);
#define DSREF_BUFFER(name) static const int DSREF_CAT(dsref_buf_, __LINE__) = dsref::addBuf(dsref_this_module, #name, &name);
#define DSREF_SAMPLER(name) static const int DSREF_CAT(dsref_tex_, __LINE__) = dsref::addTex(dsref_this_module, #name, &name);

#endif /* DSREF_DEVICE_H */
