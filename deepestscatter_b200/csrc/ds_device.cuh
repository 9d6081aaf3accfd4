/*
 * ds_device.cuh -- device-side building blocks of the radiance estimator (sm_100a).
 *
 * Two arithmetic flavours, selected by the template parameter FAST (include/ds_abi.h DsPrecision):
 *   FAST = false  software trilinear from the linear u8 grid, ds_detmath transcendental kernels,
 *                 translation unit compiled with -fmad=false  -> bit-identical to the host oracle.
 *   FAST = true   hardware 3-D texture filtering (cudaTextureObject over a block-linear cudaArray),
 *                 MUFU intrinsics, FMA contraction allowed.
 *
 * Reference lines each helper restates are cited inline (CU/ = DataGen src/CUDA/).
 */
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ds_detmath.h"

namespace dsk {

constexpr int MIE_N = 4096;
constexpr int MAX_DEPTH = 2000;                    /* CU/cloudRadianceMaterials.cu:4 */
constexpr float PI_F = 3.14159265358979323846f;
constexpr float SUN_TO_SPHERE = 5.334615707397461e-06f; /* CU/cloud.cuh:148-151 in fp32, bits 0x36b30000 */
constexpr int MAX_LEVELS = 16;

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

template <bool FAST>
__device__ __forceinline__ V3 normalize(V3 v)
{
    if (FAST) {
        return v * rsqrtf(dot(v, v));
    } else {
        const float invLen = 1.0f / sqrtf(dot(v, v)); /* optixu normalize */
        return v * invLen;
    }
}

template <bool FAST>
__device__ __forceinline__ float expNeg(float x) /* e^x */
{
    return FAST ? __expf(x) : ds_expf(x);
}
template <bool FAST>
__device__ __forceinline__ float logPos(float x)
{
    return FAST ? __logf(x) : ds_logf(x);
}
template <bool FAST>
__device__ __forceinline__ void sinCos(float phi, float* s, float* c)
{
    if (FAST) {
        __sincosf(phi, s, c);
    } else {
        ds_sincosf(phi, s, c);
    }
}

/* Everything the estimator kernels read; passed by value as a kernel argument. */
struct DevScene {
    /* volume (DG/Util/Resources.cpp:127-141): u8 [nz][ny][nx], x fastest */
    const uint8_t* density;
    const uint8_t* inscatter;
    cudaTextureObject_t densityTex;   /* FAST only */
    cudaTextureObject_t inscatterTex; /* FAST only */
    cudaTextureObject_t fusedTex;     /* k_trace_fast only: RG8 texels {density, sun transmittance}; 0 = use the two R8 arrays */
    int nx, ny, nz;
    /* DG/Scene/VDBCloud.cpp:99-110 */
    V3 bbox;
    V3 texScale;
    float mult;     /* densityMultiplier */
    float step;     /* sampleStep */
    float minRay;   /* minimalRayDistance */
    /* DG/Scene/Sun.cpp:15-17 (direction normalised) */
    V3 light;
    V3 lightColor;
    float lightIntensity;
    /* DG/Mie.cpp:8206-8297 samplers, 4096 floats each, global memory */
    const float* mie;
    const float* chopped;
    const float* cdf;
    /* empty-space occupancy bit mask over cells of (1 << occShift)^3 voxels */
    const uint32_t* occ;
    int occShift;
    int ocx, ocy, ocz;
    int occWords;
    /* Chebyshev distance (in cells, saturated at 255) from each cell to the nearest occupied cell; 0 = occupied */
    const uint8_t* cellDist;
    /* escape octants (NULL = none): bit (dx > 0) | (dy > 0) << 1 | (dz > 0) << 2 of a cell is set when the cell and every cell beyond it in
     * those three directions are empty and the grid's faces are zero: a ray there reads 0 until it leaves the box */
    const uint8_t* cellEscape;
    /* two-level guide of the chopped-Mie CDF (k_trace_fast): bucket k of guideA covers val in [k, k+1) / GUIDE_A_N, bucket k of guideB covers
     * val in [k, k+1) * GUIDE_B_LIMIT / GUIDE_B_N (val < GUIDE_B_LIMIT, where the CDF is flat and its knots are dense).  Entry = first index
     * i with cdf[i] >= bucket start; at most GUIDE_MAX_KNOTS knots lie inside a bucket (checked when the tables are built), so the inversion
     * is that many fixed probes (4 + 2 + 1 + 1), no loop.  Sized so that all tables of k_trace_fast take < 31 KiB of shared memory: the
     * SM then keeps 224 instead of 192 KiB of L1 for the texture path (every 32 KiB of L1 is worth ~5 %) */
    const uint16_t* guideA;
    const uint16_t* guideB;
    const uint16_t* choppedHalf; /* the chopped phase sampler as IEEE halves (k_trace_fast's shared-memory copy) */
    /* 1 when every voxel on the six faces of the grid is zero (VDB imports are padded by one voxel,
     * Resources.cpp:97-101): clamped taps outside the grid then read 0 */
    int borderEmpty;
};

constexpr int GUIDE_A_N = 1024;
constexpr int GUIDE_B_N = 2048;
constexpr int GUIDE_MAX_KNOTS = 8;
constexpr float GUIDE_B_LIMIT = 0.125f;
constexpr int CDF_PAD_N = MIE_N + 12; /* k_trace_fast keeps the CDF in shared memory as {0, cdf[0..4095], +inf x 11}: the probes of the inversion may read
                                         up to GUIDE_MAX_KNOTS entries past the last knot */

/* ---- CU/random.cuh ---- */

/* random.cuh:35-49, v1 = explicit stream id instead of clock() */
__device__ __forceinline__ uint32_t tea4(uint32_t val0, uint32_t stream)
{
    uint32_t v0 = val0, v1 = stream, s0 = 0;
#pragma unroll
    for (int n = 0; n < 4; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}

/* random.cuh:52-58, 67-70 */
__device__ __forceinline__ float rnd(uint32_t& prev)
{
    prev = 1664525u * prev + 1013904223u;
    return (float)(prev & 0x00FFFFFFu) * (1.0f / 16777216.0f); /* exact: division by 2^24 */
}

/* ---- texture semantics (restated; see oracle/ds_oracle.cpp header) ---- */

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

/* trilinear, clamp-to-edge, normalised coordinates, normalised-float read of a u8 volume */
__device__ __forceinline__ float tex3dSoft(const uint8_t* __restrict__ data, int nx, int ny, int nz, float u, float v, float w)
{
    const float x = u * (float)nx - 0.5f;
    const float y = v * (float)ny - 0.5f;
    const float z = w * (float)nz - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    const float tx = x - fx0, ty = y - fy0, tz = z - fz0;
    const int ix = (int)fminf(fmaxf(fx0, -2.0f), (float)nx + 1.0f);
    const int iy = (int)fminf(fmaxf(fy0, -2.0f), (float)ny + 1.0f);
    const int iz = (int)fminf(fmaxf(fz0, -2.0f), (float)nz + 1.0f);
    const int x0 = clampi(ix, 0, nx - 1), x1 = clampi(ix + 1, 0, nx - 1);
    const int y0 = clampi(iy, 0, ny - 1), y1 = clampi(iy + 1, 0, ny - 1);
    const int z0 = clampi(iz, 0, nz - 1), z1 = clampi(iz + 1, 0, nz - 1);
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    const uint8_t* r00 = data + y0 * sy + z0 * sz;
    const uint8_t* r10 = data + y1 * sy + z0 * sz;
    const uint8_t* r01 = data + y0 * sy + z1 * sz;
    const uint8_t* r11 = data + y1 * sy + z1 * sz;
    const float v000 = (float)__ldg(r00 + x0), v100 = (float)__ldg(r00 + x1);
    const float v010 = (float)__ldg(r10 + x0), v110 = (float)__ldg(r10 + x1);
    const float v001 = (float)__ldg(r01 + x0), v101 = (float)__ldg(r01 + x1);
    const float v011 = (float)__ldg(r11 + x0), v111 = (float)__ldg(r11 + x1);
    const float c00 = fmaf(tx, v100 - v000, v000);
    const float c10 = fmaf(tx, v110 - v010, v010);
    const float c01 = fmaf(tx, v101 - v001, v001);
    const float c11 = fmaf(tx, v111 - v011, v011);
    const float c0 = fmaf(ty, c10 - c00, c00);
    const float c1 = fmaf(ty, c11 - c01, c01);
    const float c = fmaf(tz, c1 - c0, c0);
    return c * (1.0f / 255.0f);
}

/* 1-D linear, clamp, normalised lookup into a 4096-entry float table (global or shared) */
__device__ __forceinline__ float tex1dSoft(const float* table, float u)
{
    const float x = u * (float)MIE_N - 0.5f;
    const float f0 = floorf(x);
    const float t = x - f0;
    const int i = (int)fminf(fmaxf(f0, -2.0f), (float)MIE_N + 1.0f);
    const int i0 = clampi(i, 0, MIE_N - 1), i1 = clampi(i + 1, 0, MIE_N - 1);
    const float a = table[i0], b = table[i1];
    return fmaf(t, b - a, a);
}

/* CU/cloud.cuh:58-62 */
template <bool FAST>
__device__ __forceinline__ float sampleCloud(const DevScene& sc, V3 pos)
{
    pos = pos * sc.texScale;
    if (FAST) {
        return tex3D<float>(sc.densityTex, pos.x, pos.y, pos.z);
    } else {
        return tex3dSoft(sc.density, sc.nx, sc.ny, sc.nz, pos.x, pos.y, pos.z);
    }
}

/* CU/cloud.cuh:64-68 */
template <bool FAST>
__device__ __forceinline__ float sampleInScatter(const DevScene& sc, V3 pos)
{
    pos = pos * sc.texScale;
    if (FAST) {
        return tex3D<float>(sc.inscatterTex, pos.x, pos.y, pos.z);
    } else {
        return tex3dSoft(sc.inscatter, sc.nx, sc.ny, sc.nz, pos.x, pos.y, pos.z);
    }
}

/* CU/cloud.cuh:40-44 */
__device__ __forceinline__ bool isInBox(const DevScene& sc, V3 pos)
{
    return pos.x >= -0.01f && pos.y >= -0.01f && pos.z >= -0.01f && pos.x <= sc.bbox.x + 0.01f && pos.y <= sc.bbox.y + 0.01f &&
           pos.z <= sc.bbox.z + 0.01f;
}

/* Is the density tap at box-local position `pos` guaranteed to read only zero voxels?
 * The tap reads voxels i0, i0+1 per axis with i0 = clamp(floor(u*N - 0.5)); cell = i0 >> occShift.
 * The occupancy bit of a cell covers voxels [c*2^s, c*2^s + 2^s] per axis (one voxel of dilation). */
__device__ __forceinline__ bool tapIsEmpty(const DevScene& sc, const uint32_t* occ, V3 pos)
{
    const float x = pos.x * sc.texScale.x * (float)sc.nx - 0.5f;
    const float y = pos.y * sc.texScale.y * (float)sc.ny - 0.5f;
    const float z = pos.z * sc.texScale.z * (float)sc.nz - 0.5f;
    const int ix = clampi((int)fminf(fmaxf(floorf(x), -2.0f), (float)sc.nx + 1.0f), 0, sc.nx - 1) >> sc.occShift;
    const int iy = clampi((int)fminf(fmaxf(floorf(y), -2.0f), (float)sc.ny + 1.0f), 0, sc.ny - 1) >> sc.occShift;
    const int iz = clampi((int)fminf(fmaxf(floorf(z), -2.0f), (float)sc.nz + 1.0f), 0, sc.nz - 1) >> sc.occShift;
    const int cell = (iz * sc.ocy + iy) * sc.ocx + ix;
    return ((occ[cell >> 5] >> (cell & 31)) & 1u) == 0u;
}

/* optix::Onb + inverse_transform (optixu_math_namespace.h; call sites CU/cloud.cuh:184-185, CU/random.cuh:170-172) */
template <bool FAST>
__device__ __forceinline__ V3 onbInverseTransform(V3 n, V3 p)
{
    V3 binormal;
    if (fabsf(n.x) > fabsf(n.z)) {
        binormal = mk(-n.y, n.x, 0.0f);
    } else {
        binormal = mk(0.0f, -n.z, n.y);
    }
    binormal = normalize<FAST>(binormal);
    const V3 tangent = cross(binormal, n);
    return p.x * tangent + p.y * binormal + p.z * n;
}

/* CU/random.cuh:122-131 */
template <bool FAST>
__device__ __forceinline__ V3 uniformOnSphereCircle(uint32_t& seed, float cosTheta)
{
    const float phi = rnd(seed) * PI_F * 2;
    const float sinTheta = sqrtf(1 - cosTheta * cosTheta);
    float s, c;
    sinCos<FAST>(phi, &s, &c);
    return mk(sinTheta * c, sinTheta * s, cosTheta);
}

/* CU/random.cuh:133-149 */
template <bool FAST>
__device__ __forceinline__ V3 uniformOnSphere(uint32_t& seed)
{
    const float u = rnd(seed);
    const float v = rnd(seed);
    const float phi = u * PI_F * 2;
    const float cosTheta = 2 * v - 1;
    const float sinTheta = sqrtf(1 - cosTheta * cosTheta);
    float s, c;
    sinCos<FAST>(phi, &s, &c);
    return mk(c * sinTheta, s * sinTheta, cosTheta);
}

/* CU/random.cuh:162-174: (x, 0, y) in the Onb of `normal` */
template <bool FAST>
__device__ __forceinline__ V3 uniformOnDisc(uint32_t& seed, V3 normal)
{
    const float theta = rnd(seed) * PI_F * 2;
    const float sqrtR = sqrtf(rnd(seed));
    float s, c;
    sinCos<FAST>(theta, &s, &c);
    return onbInverseTransform<FAST>(normal, mk(sqrtR * c, 0.0f, sqrtR * s));
}

/* CU/cloud.cuh:160-188: 16-step bisection of the chopped-Mie CDF (table in shared memory), then
 * rotate (sinT cos phi, sinT sin phi, cosT) into the frame of the previous direction */
template <bool FAST>
__device__ __forceinline__ V3 getNewDirection(const float* cdf, uint32_t& seed, V3 previousDirection)
{
    float l = 0.f, r = 1.f;
    const float val = rnd(seed);
#pragma unroll 4
    for (int i = 0; i < 16; i++) {
        const float m = (l + r) / 2.f;
        if (val > tex1dSoft(cdf, m)) {
            l = m;
        } else {
            r = m;
        }
    }
    const float cosTheta = (l + r) - 1;
    V3 d = uniformOnSphereCircle<FAST>(seed, cosTheta);
    d = onbInverseTransform<FAST>(previousDirection, d);
    return normalize<FAST>(d);
}

/* CU/cloudBBox.cu:7-37: tHit, or < 0 for no intersection */
__device__ __forceinline__ float intersectBox(const DevScene& sc, V3 o, V3 d)
{
    const V3 boxmax = mk(sc.bbox.x * 0.5f, sc.bbox.y * 0.5f, sc.bbox.z * 0.5f); /* x/2 == x*0.5 exactly */
    const V3 boxmin = -boxmax;
    const float t0x = (boxmin.x - o.x) / d.x, t1x = (boxmax.x - o.x) / d.x;
    const float t0y = (boxmin.y - o.y) / d.y, t1y = (boxmax.y - o.y) / d.y;
    const float t0z = (boxmin.z - o.z) / d.z, t1z = (boxmax.z - o.z) / d.z;
    const float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    if (tmin <= tmax) {
        return tmin > 0.0f ? tmin : sc.minRay;
    }
    return -1.0f;
}

} // namespace dsk
