/*
 * ds_detmath.h -- deterministic fp32 arithmetic contract.
 *
 * The reference (DeepestScatter DataGen) is built with --use_fast_math
 * (DeepestScatter_DataGen.vcxproj:320), so its expf/logf/sincos are the GPU's
 * approximate MUFU forms and are not reproducible on a CPU.  To make the
 * host oracle and the sm_100a kernels comparable bit for bit, both sides
 * evaluate the transcendental functions on the path with the SAME sequence
 * of IEEE-754 binary32 operations (add, mul, fma, div, sqrt, floor -- all
 * correctly rounded on x86-64 SSE/FMA and on sm_100a with -fmad=false /
 * -prec-div=true / -prec-sqrt=true).  Every multiply-add below is an explicit
 * fmaf(); nothing relies on compiler contraction (build with
 * -ffp-contract=off on the host and -fmad=false on the device).
 *
 * Polynomials are the classic Cephes single-precision kernels.  Accuracy
 * (checked in tests/test_detmath.py against numpy float64): expf <= 2 ulp on
 * [-87, 0], logf <= 2 ulp on [2^-20, 2^20], sincos abs err <= 2e-7 on [0, 2pi].
 *
 * Call sites in the reference that these replace:
 *   expf   CUDA/cloud.cuh:93, CUDA/inScatter.cu:58
 *   log    CUDA/cloud.cuh:99
 *   cos/sin CUDA/random.cuh:127-128,143-144,167-168
 *   log2f/powf CUDA/DisneyDescriptor.cuh:83,88
 */
#ifndef DS_DETMATH_H
#define DS_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DS_HD __host__ __device__ __forceinline__
#else
#define DS_HD static inline
#endif

DS_HD float ds_bits_to_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

DS_HD uint32_t ds_float_to_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

/* e^x.  Cody-Waite reduction x = n*ln2 + r, |r| <= ln2/2, degree-5 kernel. */
DS_HD float ds_expf(float x)
{
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float n = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    const float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    p = fmaf(p, z, r);
    p = p + 1.0f;
    /* 2^n with n in [-126, 127] after the clamps above */
    const int ni = (int)n;
    const float scale = ds_bits_to_float((uint32_t)(ni + 127) << 23);
    return p * scale;
}

/* ln(x) for positive normal x. */
DS_HD float ds_logf(float x)
{
    const uint32_t bits = ds_float_to_bits(x);
    int e = (int)((bits >> 23) & 0xffu) - 126;                 /* x = m * 2^e, m in [0.5, 1) */
    float m = ds_bits_to_float((bits & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) {
        e -= 1;
        m = (m + m) - 1.0f;
    } else {
        m = m - 1.0f;
    }
    const float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    const float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

/* sin and cos of phi, phi in [0, 2*pi] (the only range the path uses:
 * phi = rnd * pi * 2).  Quadrant reduction with a 3-part pi/2. */
DS_HD void ds_sincosf(float phi, float* s, float* c)
{
    const float k = floorf(fmaf(phi, 0.636619772367581343f, 0.5f)); /* 0..4 */
    float r = fmaf(k, -1.5703125f, phi);
    r = fmaf(k, -4.83751296997070312e-4f, r);
    r = fmaf(k, -7.54978995489188216e-8f, r);
    const float z = r * r;
    float sp = -1.9515295891e-4f;
    sp = fmaf(sp, z, 8.3321608736e-3f);
    sp = fmaf(sp, z, -1.6666654611e-1f);
    const float sr = fmaf(sp * z, r, r);
    float cp = 2.443315711809948e-5f;
    cp = fmaf(cp, z, -1.388731625493765e-3f);
    cp = fmaf(cp, z, 4.166664568298827e-2f);
    float cr = (cp * z) * z;
    cr = fmaf(-0.5f, z, cr);
    cr = cr + 1.0f;
    const int q = ((int)k) & 3;
    const float ss = (q & 1) ? cr : sr;
    const float cc = (q & 1) ? sr : cr;
    *s = (q & 2) ? -ss : ss;
    *c = ((q + 1) & 2) ? -cc : cc;
}

DS_HD float ds_log2f(float x)
{
    return ds_logf(x) * 1.44269504088896341f;
}

DS_HD float ds_exp2f(float x)
{
    return ds_expf(x * 0.693147180559945309f);
}

#endif /* DS_DETMATH_H */
