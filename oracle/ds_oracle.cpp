/*
 * ds_oracle.cpp -- HOST ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain, single-threaded-per-path CPU restatement of the reference's radiance
 * estimation path (marsermd/DeepestScatter, DataGen).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (deepestscatter_b200/csrc) never links or
 * calls it and has no CPU fallback.
 *
 * PARITY STATUS: **pinned to the reference's own source.**  The reference has
 * no tests, golden images or fixtures, but its DataGen code --
 * the CUDA/ *.cu device programs AND the host classes that drive them -- is
 * compiled UNMODIFIED from /root/reference by g++ against a small OptiX 5.1
 * emulation (oracle/ref_shim/, output oracle/_ref/libds_ref.so).
 * tests/test_oracle_vs_ref.py holds every entry point of this file to that
 * library BIT FOR BIT on seeded inputs (estimators in all three modes with
 * equal step/event counts, bake bytes, mip chain, Mie samplers, scene
 * variables, camera frames, Welford buffers, Reinhard bytes, generated points,
 * descriptor bytes, network-input pass, the RadianceCollector schedule, the
 * importer); tests/golden/ref_path.npz (tools/make_golden_ref.py) keeps vectors
 * of that library for machines without /root/reference.  Also pinned from
 * reference artefacts: the protobuf record bytes (tests/golden/records.json,
 * from the reference's own PythonProtocols _pb2 modules).  What the pin cannot
 * cover is listed next: arithmetic of the absent SDK, defined identically in
 * this file and in the emulation.
 *
 * Third-party arithmetic not present under /root/reference and therefore
 * DEFINED here (NVIDIA OptiX SDK 5.1.0 / CUDA 9.2, Dependencies.md:3-4):
 *   - tex3D / rtTex3D / rtTex3DLod on u8, linear, clamp-to-edge, normalized
 *     coordinates, normalized-float read:  x = u*N - 0.5, i = floor(x),
 *     weights frac(x) in exact fp32 (the hardware uses 8 fractional bits),
 *     indices clamped; LOD clamps to [0, L-1] and lerps the two levels.
 *   - tex1D on float, linear, clamp, normalized.
 *   - optix::Onb, normalize, cross, dot, lerp, float3/float (a * (1/s)).
 *   - rtPotentialIntersection(t): tmin < t < tmax, with tmin = sceneEPS = 0
 *     (declared at cameraCommon.cuh:13, never set by the host).
 *   - expf/logf/sin/cos: the deterministic kernels of include/ds_detmath.h
 *     (the reference uses fast-math MUFU approximations).
 *   - clock() in the RNG seed (random.cuh:38) is replaced by an explicit
 *     stream id: subframeId for renders, frameId for point radiance, 0 for
 *     point generation.
 *
 * Abbreviations in citations: CU/ = DeepestScatter_DataGen/DeepestScatter_DataGen/src/CUDA/,
 * DG/ = DeepestScatter_DataGen/DeepestScatter_DataGen/src/.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ds_detmath.h"
#include "../include/ds_synth.h"

namespace {

struct f3 {
    float x, y, z;
};
inline f3 mk(float x, float y, float z) { return f3{x, y, z}; }
inline f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline f3 operator-(f3 a) { return mk(-a.x, -a.y, -a.z); }
inline f3 operator*(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline f3 operator*(f3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline f3 operator*(float s, f3 a) { return mk(a.x * s, a.y * s, a.z * s); }
inline f3 operator/(f3 a, f3 b) { return mk(a.x / b.x, a.y / b.y, a.z / b.z); }
/* optixu: float3 / float multiplies by the reciprocal */
inline f3 operator/(f3 a, float s)
{
    const float inv = 1.0f / s;
    return a * inv;
}
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline f3 cross(f3 a, f3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline f3 normalize(f3 v)
{
    const float invLen = 1.0f / sqrtf(dot(v, v));
    return v * invLen;
}
inline float length(f3 v) { return sqrtf(dot(v, v)); }
inline float lerpf(float a, float b, float t) { return a + t * (b - a); }
inline float saturatef(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

constexpr float PI_F = 3.14159265358979323846f;
/* sunArea / sphereArea of CU/cloud.cuh:148-151 evaluated once in fp32:
 * 2*pi*(1 - cosf(0.265*pi/180)) / (4*pi), bits 0x36b30000 */
constexpr float SUN_TO_SPHERE = 5.334615707397461e-06f;
constexpr int MAX_DEPTH = 2000; /* CU/cloudRadianceMaterials.cu:4 */
constexpr int MIE_N = 4096;

struct Level {
    int nx, ny, nz;
    std::vector<uint8_t> v;
};

struct Scene {
    /* volume: DG/Util/Resources.cpp:68-155 */
    std::vector<Level> levels;
    std::vector<uint8_t> inScatter;
    /* DG/Scene/VDBCloud.cpp:99-110 */
    f3 bboxSize{1, 1, 1};
    f3 textureScale{1, 1, 1};
    float densityMultiplier = 700.0f;
    float cloudSizeInMeters = 7000.0f;
    float voxelSizeInMeters = 0;
    float voxelSizeInTermsOfFreePath = 0;
    /* DG/installers.cpp:86, DG/Scene/CloudMaterial.cpp:23 */
    float sampleStep = 1.0f / 512.0f;
    float minimalRayDistance = 0.000001f;
    /* DG/Scene/Sun.cpp:15-17 */
    f3 lightDirection{0, -1, 0};
    f3 lightColor{1, 1, 1};
    float lightIntensity = 1e6f;
    /* DG/Mie.cpp:8206-8297 */
    float mie[MIE_N], choppedMie[MIE_N], choppedMieIntegral[MIE_N];
};

/* work counters: per-thread, folded into the totals by the API entry points */
thread_local unsigned long long tlPaths = 0, tlEvents = 0, tlSteps = 0;
unsigned long long gPaths = 0, gEvents = 0, gSteps = 0;
void foldCounters()
{
#pragma omp critical(orc_counters)
    {
        gPaths += tlPaths;
        gEvents += tlEvents;
        gSteps += tlSteps;
    }
    tlPaths = tlEvents = tlSteps = 0;
}

/* ---- texture fetch definitions (OptiX/CUDA semantics restated) ---- */

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

float tex3d(const Level& L, const uint8_t* data, float u, float v, float w)
{
    const float x = u * (float)L.nx - 0.5f;
    const float y = v * (float)L.ny - 0.5f;
    const float z = w * (float)L.nz - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    const float tx = x - fx0, ty = y - fy0, tz = z - fz0;
    /* clamp in float first so huge coordinates cannot overflow the int cast */
    const float cx = fminf(fmaxf(fx0, -2.0f), (float)L.nx + 1.0f);
    const float cy = fminf(fmaxf(fy0, -2.0f), (float)L.ny + 1.0f);
    const float cz = fminf(fmaxf(fz0, -2.0f), (float)L.nz + 1.0f);
    const int ix = (int)cx, iy = (int)cy, iz = (int)cz;
    const int x0 = clampi(ix, 0, L.nx - 1), x1 = clampi(ix + 1, 0, L.nx - 1);
    const int y0 = clampi(iy, 0, L.ny - 1), y1 = clampi(iy + 1, 0, L.ny - 1);
    const int z0 = clampi(iz, 0, L.nz - 1), z1 = clampi(iz + 1, 0, L.nz - 1);
    const size_t sx = 1, sy = (size_t)L.nx, sz = (size_t)L.nx * L.ny;
    auto at = [&](int xx, int yy, int zz) { return (float)data[xx * sx + yy * sy + zz * sz]; };
    const float c00 = fmaf(tx, at(x1, y0, z0) - at(x0, y0, z0), at(x0, y0, z0));
    const float c10 = fmaf(tx, at(x1, y1, z0) - at(x0, y1, z0), at(x0, y1, z0));
    const float c01 = fmaf(tx, at(x1, y0, z1) - at(x0, y0, z1), at(x0, y0, z1));
    const float c11 = fmaf(tx, at(x1, y1, z1) - at(x0, y1, z1), at(x0, y1, z1));
    const float c0 = fmaf(ty, c10 - c00, c00);
    const float c1 = fmaf(ty, c11 - c01, c01);
    const float c = fmaf(tz, c1 - c0, c0);
    return c * (1.0f / 255.0f);
}

float tex3dLod(const Scene& s, float u, float v, float w, float lod)
{
    const int last = (int)s.levels.size() - 1;
    float l = fminf(fmaxf(lod, 0.0f), (float)last);
    const float lf = floorf(l);
    const int l0 = (int)lf;
    const float t = l - lf;
    const float a = tex3d(s.levels[l0], s.levels[l0].v.data(), u, v, w);
    if (l0 >= last || t == 0.0f) return a;
    const float b = tex3d(s.levels[l0 + 1], s.levels[l0 + 1].v.data(), u, v, w);
    return fmaf(t, b - a, a);
}

float tex1d(const float* table, float u)
{
    const float x = u * (float)MIE_N - 0.5f;
    const float f0 = floorf(x);
    const float t = x - f0;
    const float c = fminf(fmaxf(f0, -2.0f), (float)MIE_N + 1.0f);
    const int i = (int)c;
    const int i0 = clampi(i, 0, MIE_N - 1), i1 = clampi(i + 1, 0, MIE_N - 1);
    return fmaf(t, table[i1] - table[i0], table[i0]);
}

/* ---- CU/random.cuh ---- */

/* random.cuh:35-49 with v1 = stream instead of clock() */
uint32_t tea4(uint32_t val0, uint32_t stream)
{
    uint32_t v0 = val0, v1 = stream, s0 = 0;
    for (int n = 0; n < 4; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}

/* random.cuh:52-58 */
uint32_t lcg(uint32_t& prev)
{
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00FFFFFFu;
}

/* random.cuh:67-70 */
float rnd(uint32_t& prev) { return (float)lcg(prev) / (float)0x01000000; }

struct Onb {
    f3 tangent, binormal, normal;
    explicit Onb(f3 n)
    {
        normal = n;
        if (fabsf(n.x) > fabsf(n.z)) {
            binormal = mk(-n.y, n.x, 0.0f);
        } else {
            binormal = mk(0.0f, -n.z, n.y);
        }
        binormal = normalize(binormal);
        tangent = cross(binormal, normal);
    }
    f3 inverse_transform(f3 p) const { return p.x * tangent + p.y * binormal + p.z * normal; }
};

/* random.cuh:122-131 */
f3 uniformOnSphereCircle(uint32_t& prev, float cosTheta)
{
    const float phi = rnd(prev) * PI_F * 2;
    const float sinTheta = sqrtf(1 - cosTheta * cosTheta);
    float s, c;
    ds_sincosf(phi, &s, &c);
    return mk(sinTheta * c, sinTheta * s, cosTheta);
}

/* random.cuh:133-149 */
f3 uniformOnSphere(uint32_t& prev)
{
    const float u = rnd(prev);
    const float v = rnd(prev);
    const float phi = u * PI_F * 2;
    const float cosTheta = 2 * v - 1;
    const float sinTheta = sqrtf(1 - cosTheta * cosTheta);
    float s, c;
    ds_sincosf(phi, &s, &c);
    return mk(c * sinTheta, s * sinTheta, cosTheta);
}

/* random.cuh:162-174 -- note (x, 0, y): the disc spans tangent x NORMAL */
f3 uniformOnDisc(uint32_t& prev, f3 normal)
{
    const float theta = rnd(prev) * PI_F * 2;
    const float sqrtR = sqrtf(rnd(prev));
    float s, c;
    ds_sincosf(theta, &s, &c);
    const float x = sqrtR * c;
    const float y = sqrtR * s;
    Onb onb(normal);
    return onb.inverse_transform(mk(x, 0, y));
}

/* ---- CU/cloud.cuh ---- */

/* cloud.cuh:40-44 */
bool isInBox(const Scene& s, f3 pos)
{
    return pos.x >= -0.01f && pos.y >= -0.01f && pos.z >= -0.01f && pos.x <= s.bboxSize.x + 0.01f &&
           pos.y <= s.bboxSize.y + 0.01f && pos.z <= s.bboxSize.z + 0.01f;
}

/* cloud.cuh:46-54 */
float getMiePhase(const Scene& s, float cosTheta) { return tex1d(s.mie, (cosTheta + 1) / 2); }
float getChoppedMiePhase(const Scene& s, float cosTheta) { return tex1d(s.choppedMie, (cosTheta + 1) / 2); }

/* cloud.cuh:58-62 */
float sampleCloud(const Scene& s, f3 pos)
{
    pos = pos * s.textureScale;
    return tex3d(s.levels[0], s.levels[0].v.data(), pos.x, pos.y, pos.z);
}

/* cloud.cuh:64-68 */
float sampleInScatter(const Scene& s, f3 pos)
{
    pos = pos * s.textureScale;
    return tex3d(s.levels[0], s.inScatter.data(), pos.x, pos.y, pos.z);
}

struct ScatteringEvent {
    bool hasScattered;
    f3 scatterPos;
    float transmittance;
};

/* cloud.cuh:77-114 */
ScatteringEvent getNextScatteringEvent(const Scene& s, float opticalDistance, f3 pos, f3 direction, bool stopAtScatterPos = true)
{
    const f3 stepAlongRay = direction * s.sampleStep;
    float transmittance = 1;
    bool hasScattered = false;
    f3 scatterPos = mk(0, 0, 0);
    while (isInBox(s, pos)) {
        pos = pos + stepAlongRay;
        tlSteps++;
        const float density = sampleCloud(s, pos) * s.densityMultiplier;
        const float extinction = density * s.sampleStep;
        const float currentTransmit = ds_expf(-extinction);
        transmittance *= currentTransmit;
        if (!hasScattered && opticalDistance > transmittance) {
            hasScattered = true;
            scatterPos = pos - direction * ds_logf(opticalDistance / transmittance) / density;
            if (stopAtScatterPos) break;
        }
    }
    if (!hasScattered && !isInBox(s, pos)) scatterPos = pos;
    return ScatteringEvent{hasScattered, scatterPos, transmittance};
}

/* cloud.cuh:116-122 */
ScatteringEvent getNextScatteringEvent(const Scene& s, uint32_t& seed, f3 pos, f3 direction)
{
    const float opticalDistance = rnd(seed);
    return getNextScatteringEvent(s, opticalDistance, pos, direction);
}

/* cloud.cuh:146-158 */
f3 getInScattering(const Scene& s, const ScatteringEvent& scatter, f3 direction, bool choppedMiePhase)
{
    const float cosLightAngle = dot(-s.lightDirection, direction);
    const float phase = choppedMiePhase ? getChoppedMiePhase(s, cosLightAngle) : getMiePhase(s, cosLightAngle);
    tlEvents++;
    return s.lightColor * s.lightIntensity * sampleInScatter(s, scatter.scatterPos) * phase * SUN_TO_SPHERE;
}

/* cloud.cuh:160-188 */
f3 getNewDirection(const Scene& s, uint32_t& seed, f3 previousDirection)
{
    float l = 0.f, r = 1.f, m = 0.5f;
    const float val = rnd(seed);
    for (int i = 0; i < 16; i++) {
        m = (l + r) / 2.f;
        if (val > tex1d(s.choppedMieIntegral, m)) {
            l = m;
        } else {
            r = m;
        }
    }
    const float cosTheta = (l + r) - 1;
    f3 newDirection = uniformOnSphereCircle(seed, cosTheta);
    Onb onb(previousDirection);
    newDirection = onb.inverse_transform(newDirection);
    return normalize(newDirection);
}

/* ---- CU/cloudBBox.cu:7-37: returns tHit or a negative value for "miss" ---- */
float intersectBox(const Scene& s, f3 origin, f3 direction)
{
    const f3 boxmin = -s.bboxSize / 2, boxmax = s.bboxSize / 2;
    const f3 t0 = (boxmin - origin) / direction;
    const f3 t1 = (boxmax - origin) / direction;
    const f3 tnear = mk(fminf(t0.x, t1.x), fminf(t0.y, t1.y), fminf(t0.z, t1.z));
    const f3 tfar = mk(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z));
    const float tmin = fmaxf(fmaxf(tnear.x, tnear.y), tnear.z);
    const float tmax = fminf(fminf(tfar.x, tfar.y), tfar.z);
    if (tmin <= tmax) {
        if (tmin > 0.0f) return tmin;                 /* rtPotentialIntersection(tmin), ray tmin = sceneEPS = 0 */
        return s.minimalRayDistance;                  /* checkBack: 0 < 1e-6 < RT_DEFAULT_MAX */
    }
    return -1.0f;
}

enum Mode { ALL_SCATTER = 0, MULTIPLE_SCATTER = 1, SINGLE_SCATTER = 2 };

/* CU/cloudRadianceMaterials.cu:9-66 (sky branch is compiled out: shouldSampleSky = false, :25) */
f3 totalRadiance(const Scene& s, f3 hitPoint, f3 rayDirection, uint32_t seed)
{
    f3 radiance = mk(0, 0, 0);
    f3 pos = hitPoint;
    f3 direction = normalize(rayDirection);
    int depth = 0;
    while (isInBox(s, pos)) {
        depth++;
        if (depth == MAX_DEPTH) break;
        ScatteringEvent scatter = getNextScatteringEvent(s, seed, pos, direction);
        if (!scatter.hasScattered || !isInBox(s, scatter.scatterPos)) {
            break;
        } else {
            radiance = radiance + getInScattering(s, scatter, direction, depth != 1);
            pos = scatter.scatterPos;
            direction = getNewDirection(s, seed, direction);
        }
    }
    return radiance;
}

/* CU/cloudRadianceMaterials.cu:72-115 */
f3 multipleScatterSunRadiance(const Scene& s, f3 hitPoint, f3 rayDirection, uint32_t seed)
{
    f3 radiance = mk(0, 0, 0);
    f3 pos = hitPoint;
    f3 direction = normalize(rayDirection);
    direction = getNewDirection(s, seed, direction);
    int depth = 0;
    while (isInBox(s, pos)) {
        depth++;
        if (depth == MAX_DEPTH) break;
        ScatteringEvent scatter = getNextScatteringEvent(s, seed, pos, direction);
        if (!scatter.hasScattered || !isInBox(s, scatter.scatterPos)) {
            break;
        } else {
            radiance = radiance + getInScattering(s, scatter, direction, true);
            pos = scatter.scatterPos;
            direction = getNewDirection(s, seed, direction);
        }
    }
    return radiance;
}

/* CU/cloudRadianceMaterials.cu:120-148 */
f3 singleScatterSunRadiance(const Scene& s, f3 hitPoint, f3 rayDirection, uint32_t seed)
{
    f3 radiance = mk(0, 0, 0);
    f3 pos = hitPoint;
    f3 direction = normalize(rayDirection);
    ScatteringEvent scatter = getNextScatteringEvent(s, seed, pos, direction);
    if (!scatter.hasScattered || !isInBox(s, scatter.scatterPos)) {
    } else {
        radiance = radiance + getInScattering(s, scatter, direction, false);
    }
    return radiance;
}

/* rtTrace from (origin, direction) against the single box + closest hit by mode.
 * seedVal0 is launchID.x * 4096 + launchID.y (cloudRadianceMaterials.cu:21). */
f3 traceRadiance(const Scene& s, int mode, f3 origin, f3 direction, uint32_t seedVal0, uint32_t stream)
{
    tlPaths++;
    const float tHit = intersectBox(s, origin, direction);
    if (tHit < 0.0f) return mk(0, 0, 0); /* miss program is a no-op: progressive.cu:44-46 */
    f3 hitPoint = origin + tHit * direction;
    hitPoint = hitPoint + 0.5f * s.bboxSize;
    const uint32_t seed = tea4(seedVal0, stream);
    switch (mode) {
    case ALL_SCATTER:
        return totalRadiance(s, hitPoint, direction, seed);
    case MULTIPLE_SCATTER:
        return multipleScatterSunRadiance(s, hitPoint, direction, seed);
    default:
        return singleScatterSunRadiance(s, hitPoint, direction, seed);
    }
}

/* CU/cameraCommon.cuh:19-29 + CU/pathTracingCamera.cu:12-21 */
f3 cameraDirection(const float* cam /* eye,U,V,W */, uint32_t px, uint32_t py, uint32_t w, uint32_t h)
{
    const float dx = (float)px / (float)w * 2.f - 1.f;
    const float dy = (float)py / (float)h * 2.f - 1.f;
    const f3 U = mk(cam[3], cam[4], cam[5]), V = mk(cam[6], cam[7], cam[8]), W = mk(cam[9], cam[10], cam[11]);
    return normalize(dx * U + dy * V + W);
}

void setupDerived(Scene& s)
{
    const Level& L = s.levels[0];
    /* VDBCloud.cpp:101-106: bboxSize is the voxel-count size, normalised by its max */
    const float fx = (float)L.nx, fy = (float)L.ny, fz = (float)L.nz;
    const float maxSize = std::max({fx, fy, fz});
    s.bboxSize = mk(fx / maxSize, fy / maxSize, fz / maxSize);
    s.textureScale = mk(maxSize / fx, maxSize / fy, maxSize / fz);
    /* VDBCloud.cpp:35-46 */
    const size_t mx = (size_t)std::max({L.nx, L.ny, L.nz});
    s.voxelSizeInMeters = s.cloudSizeInMeters / mx;
}

} // namespace

/* ====================================================================== */
/*                               C API                                    */
/* ====================================================================== */

extern "C" {

void* orc_create() { return new Scene(); }
void orc_destroy(void* h) { delete (Scene*)h; }

/* DG/Mie.cpp:8206-8282: phase sampler = table / mean, integral = running sum of table / sum */
void orc_set_mie(void* h, const float* mie, const float* chopped)
{
    Scene& s = *(Scene*)h;
    auto phase = [](const float* src, float* dst) {
        float average = 0;
        for (int i = 0; i < MIE_N; i++) average += src[i];
        average /= MIE_N;
        for (int i = 0; i < MIE_N; i++) dst[i] = src[i] / average;
    };
    phase(mie, s.mie);
    phase(chopped, s.choppedMie);
    float sum = 0;
    for (int i = 0; i < MIE_N; i++) sum += chopped[i];
    float integral = 0;
    for (int i = 0; i < MIE_N; i++) {
        integral += chopped[i] / sum;
        s.choppedMieIntegral[i] = integral;
    }
}

void orc_get_mie(void* h, float* mie, float* chopped, float* integral)
{
    Scene& s = *(Scene*)h;
    memcpy(mie, s.mie, sizeof(s.mie));
    memcpy(chopped, s.choppedMie, sizeof(s.choppedMie));
    memcpy(integral, s.choppedMieIntegral, sizeof(s.choppedMieIntegral));
}

/* DG/Util/Resources.cpp:169-209: box-filter mip chain, out-of-range children read 0, integer /8 */
static void buildMips(Scene& s)
{
    int maxSize = std::max({s.levels[0].nx, s.levels[0].ny, s.levels[0].nz});
    int levelCount = 1;
    while (maxSize /= 2) levelCount++; /* Resources.cpp:110-115 */
    for (int level = 1; level < levelCount; level++) {
        const Level& p = s.levels[level - 1];
        Level c;
        c.nx = std::max(1, s.levels[0].nx >> level); /* optix getMipLevelSize */
        c.ny = std::max(1, s.levels[0].ny >> level);
        c.nz = std::max(1, s.levels[0].nz >> level);
        c.v.resize((size_t)c.nx * c.ny * c.nz);
        auto get = [&](int x, int y, int z) -> uint16_t {
            if (x < 0 || x >= p.nx || y < 0 || y >= p.ny || z < 0 || z >= p.nz) return 0;
            return p.v[(size_t)z * p.nx * p.ny + (size_t)y * p.nx + x];
        };
        for (int z = 0; z < c.nz; z++)
            for (int y = 0; y < c.ny; y++)
                for (int x = 0; x < c.nx; x++) {
                    uint16_t cur = get(x * 2, y * 2, z * 2) + get(x * 2, y * 2, z * 2 + 1) + get(x * 2, y * 2 + 1, z * 2) +
                                   get(x * 2, y * 2 + 1, z * 2 + 1) + get(x * 2 + 1, y * 2, z * 2) +
                                   get(x * 2 + 1, y * 2, z * 2 + 1) + get(x * 2 + 1, y * 2 + 1, z * 2) +
                                   get(x * 2 + 1, y * 2 + 1, z * 2 + 1);
                    cur /= 8;
                    c.v[(size_t)z * c.nx * c.ny + (size_t)y * c.nx + x] = (uint8_t)cur;
                }
        s.levels.push_back(std::move(c));
    }
}

void orc_volume_set(void* h, const uint8_t* data, int nx, int ny, int nz, int buildMipmaps)
{
    Scene& s = *(Scene*)h;
    s.levels.clear();
    Level L;
    L.nx = nx;
    L.ny = ny;
    L.nz = nz;
    L.v.assign(data, data + (size_t)nx * ny * nz);
    s.levels.push_back(std::move(L));
    if (buildMipmaps) buildMips(s);
    s.inScatter.clear();
    setupDerived(s);
}

/* DG/Util/Resources.cpp:127-141: dense float grid -> u8 = narrow_cast<uint8_t>(value / maxDensity * 255),
 * maxDensity a double (openvdb Extrema::max, :95) */
void orc_quantize_float_grid(const float* values, size_t count, double maxDensity, uint8_t* out)
{
    for (size_t i = 0; i < count; i++) out[i] = (uint8_t)(values[i] / maxDensity * 255);
}

void orc_volume_synth(void* h, int n, int kind, uint32_t seed, int buildMipmaps)
{
    std::vector<uint8_t> v((size_t)n * n * n);
#pragma omp parallel for schedule(static)
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) v[((size_t)z * n + y) * n + x] = ds_synth_voxel(kind, seed, n, x, y, z);
    orc_volume_set(h, v.data(), n, n, n, buildMipmaps);
}

int orc_volume_level_count(void* h) { return (int)((Scene*)h)->levels.size(); }

void orc_volume_level_dims(void* h, int level, int* dims)
{
    const Level& L = ((Scene*)h)->levels[level];
    dims[0] = L.nx;
    dims[1] = L.ny;
    dims[2] = L.nz;
}

void orc_volume_level_get(void* h, int level, uint8_t* out)
{
    const Level& L = ((Scene*)h)->levels[level];
    memcpy(out, L.v.data(), L.v.size());
}

/* scene parameters: cloud size (m), mean free path (m), sample step, light (normalised as
 * DirectionalLight's ctor does, SceneDescription.h:17-19), colour, intensity */
void orc_scene_set(void* h, float cloudSizeM, float meanFreePathM, float sampleStep, const float* lightDir,
                   const float* lightColor, float lightIntensity)
{
    Scene& s = *(Scene*)h;
    s.cloudSizeInMeters = cloudSizeM;
    s.densityMultiplier = cloudSizeM / meanFreePathM; /* VDBCloud.cpp:109 */
    s.sampleStep = sampleStep;
    s.lightDirection = normalize(mk(lightDir[0], lightDir[1], lightDir[2]));
    s.lightColor = mk(lightColor[0], lightColor[1], lightColor[2]);
    s.lightIntensity = lightIntensity;
    setupDerived(s);
    s.voxelSizeInTermsOfFreePath = s.voxelSizeInMeters / meanFreePathM; /* VDBCloud.cpp:43-46 */
}

void orc_scene_get_derived(void* h, float* out /* bbox[3], texScale[3], mult, voxelM, voxelFP, light[3] */)
{
    Scene& s = *(Scene*)h;
    const float v[] = {s.bboxSize.x, s.bboxSize.y, s.bboxSize.z, s.textureScale.x, s.textureScale.y, s.textureScale.z,
                       s.densityMultiplier, s.voxelSizeInMeters, s.voxelSizeInTermsOfFreePath, s.lightDirection.x,
                       s.lightDirection.y, s.lightDirection.z};
    memcpy(out, v, sizeof(v));
}

/* CU/inScatter.cu:40-66 */
void orc_bake_inscatter(void* h)
{
    Scene& s = *(Scene*)h;
    const Level& L = s.levels[0];
    s.inScatter.assign(L.v.size(), 0);
    const size_t maxSize = (size_t)std::max({L.nx, L.ny, L.nz});
    const float minScale = fminf(fminf(s.textureScale.x, s.textureScale.y), s.textureScale.z);
    const f3 stepToLight = (-normalize(s.lightDirection)) * s.sampleStep;
    const int stepCount = (int)(1 / s.sampleStep);
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z < L.nz; z++)
        for (int y = 0; y < L.ny; y++)
            for (int x = 0; x < L.nx; x++) {
                f3 samplePos = (mk((float)x, (float)y, (float)z) / mk((float)maxSize, (float)maxSize, (float)maxSize)) / minScale;
                float transmittance = 1;
                for (int i = 0; i < stepCount; i++) {
                    const float density = sampleCloud(s, samplePos) * s.densityMultiplier;
                    const float extinction = density * s.sampleStep;
                    transmittance *= ds_expf(-extinction);
                    samplePos = samplePos + stepToLight;
                    if (transmittance * 255.f < 1.f) break;
                }
                s.inScatter[((size_t)z * L.ny + y) * L.nx + x] = (uint8_t)(transmittance * 255.f);
            }
}

/* Same bake, for the large benchmark grids: taps whose 2x2x2 footprint lies in an all-zero 8^3 cell (plus one
 * voxel of dilation) are skipped, which cannot change the result (density 0 -> exp(-0) == 1, T unchanged).
 * tests/test_oracle.py asserts byte equality with orc_bake_inscatter. */
void orc_bake_inscatter_skip(void* h)
{
    Scene& s = *(Scene*)h;
    const Level& L = s.levels[0];
    const int shift = 3, c = 1 << shift;
    const int ocx = (L.nx + c - 1) / c, ocy = (L.ny + c - 1) / c, ocz = (L.nz + c - 1) / c;
    std::vector<uint8_t> occ((size_t)ocx * ocy * ocz, 0);
#pragma omp parallel for schedule(static)
    for (int cz = 0; cz < ocz; cz++)
        for (int cy = 0; cy < ocy; cy++)
            for (int cx = 0; cx < ocx; cx++) {
                bool any = false;
                for (int z = cz * c; z <= std::min(cz * c + c, L.nz - 1) && !any; z++)
                    for (int y = cy * c; y <= std::min(cy * c + c, L.ny - 1) && !any; y++)
                        for (int x = cx * c; x <= std::min(cx * c + c, L.nx - 1); x++)
                            if (L.v[((size_t)z * L.ny + y) * L.nx + x]) {
                                any = true;
                                break;
                            }
                occ[((size_t)cz * ocy + cy) * ocx + cx] = any;
            }
    auto cellOf = [&](float u, int n) {
        const float x = u * (float)n - 0.5f;
        const int i = clampi((int)fminf(fmaxf(floorf(x), -2.0f), (float)n + 1.0f), 0, n - 1);
        return i >> shift;
    };
    const size_t maxSize = (size_t)std::max({L.nx, L.ny, L.nz});
    const float minScale = fminf(fminf(s.textureScale.x, s.textureScale.y), s.textureScale.z);
    const f3 stepToLight = (-normalize(s.lightDirection)) * s.sampleStep;
    const int stepCount = (int)(1 / s.sampleStep);
    /* Cells whose voxels all see the sun unoccluded: walk the sun ray of the cell centre in quarter-cell
     * strides and require the 5x5x5 cell neighbourhood of every sample to be empty.  Every tap of every voxel
     * ray of the cell stays within one cell (Chebyshev) of some centre sample, so T stays exactly 1 -> 255. */
    std::vector<uint8_t> lit((size_t)ocx * ocy * ocz, 0);
    {
        const float voxel = 1.0f / (float)maxSize; /* box-local length of one voxel */
        const float stride = 0.25f * c * voxel;
        const float reach = (float)stepCount * s.sampleStep + 2.0f * c * voxel;
        const f3 toLight = -normalize(s.lightDirection);
#pragma omp parallel for schedule(dynamic, 1)
        for (int cz = 0; cz < ocz; cz++)
            for (int cy = 0; cy < ocy; cy++)
                for (int cx = 0; cx < ocx; cx++) {
                    const f3 centre = mk((cx + 0.5f) * c * voxel, (cy + 0.5f) * c * voxel, (cz + 0.5f) * c * voxel);
                    bool clear = true;
                    for (float t = -2.0f * c * voxel; t <= reach && clear; t += stride) {
                        const f3 q = centre + toLight * t;
                        const int qx = (int)floorf(q.x / (c * voxel)), qy = (int)floorf(q.y / (c * voxel)), qz = (int)floorf(q.z / (c * voxel));
                        for (int dz = -2; dz <= 2 && clear; dz++)
                            for (int dy = -2; dy <= 2 && clear; dy++)
                                for (int dx = -2; dx <= 2; dx++) {
                                    const int ax = qx + dx, ay = qy + dy, az = qz + dz;
                                    if (ax < 0 || ay < 0 || az < 0 || ax >= ocx || ay >= ocy || az >= ocz) continue;
                                    if (occ[((size_t)az * ocy + ay) * ocx + ax]) {
                                        clear = false;
                                        break;
                                    }
                                }
                    }
                    lit[((size_t)cz * ocy + cy) * ocx + cx] = clear;
                }
    }
    s.inScatter.assign(L.v.size(), 0);
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int z = 0; z < L.nz; z++)
        for (int y = 0; y < L.ny; y++)
            for (int x = 0; x < L.nx; x++) {
                const size_t o = ((size_t)z * L.ny + y) * L.nx + x;
                if (lit[((size_t)(z >> shift) * ocy + (y >> shift)) * ocx + (x >> shift)]) {
                    s.inScatter[o] = 255;
                    continue;
                }
                f3 samplePos = (mk((float)x, (float)y, (float)z) / mk((float)maxSize, (float)maxSize, (float)maxSize)) / minScale;
                float transmittance = 1;
                for (int i = 0; i < stepCount; i++) {
                    const f3 uvw = samplePos * s.textureScale;
                    const size_t cell = ((size_t)cellOf(uvw.z, L.nz) * ocy + cellOf(uvw.y, L.ny)) * ocx + cellOf(uvw.x, L.nx);
                    if (occ[cell]) {
                        const float density = sampleCloud(s, samplePos) * s.densityMultiplier;
                        const float extinction = density * s.sampleStep;
                        transmittance *= ds_expf(-extinction);
                    }
                    samplePos = samplePos + stepToLight;
                    if (transmittance * 255.f < 1.f) break;
                }
                s.inScatter[o] = (uint8_t)(transmittance * 255.f);
            }
}

void orc_inscatter_get(void* h, uint8_t* out)
{
    Scene& s = *(Scene*)h;
    memcpy(out, s.inScatter.data(), s.inScatter.size());
}

void orc_inscatter_set(void* h, const uint8_t* in)
{
    Scene& s = *(Scene*)h;
    s.inScatter.assign(in, in + s.levels[0].v.size());
}

/* sampler probes for the parity tests: which = 0 density, 1 inScatter; lod < 0 means plain tex3D */
void orc_sample_volume(void* h, int which, const float* pos, int n, float lod, float* out)
{
    Scene& s = *(Scene*)h;
    for (int i = 0; i < n; i++) {
        f3 p = mk(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        if (which == 1) {
            out[i] = sampleInScatter(s, p);
        } else if (lod < 0) {
            out[i] = sampleCloud(s, p);
        } else {
            p = p * s.textureScale;
            out[i] = tex3dLod(s, p.x, p.y, p.z, lod);
        }
    }
}

void orc_sample_table(void* h, int which, const float* u, int n, float* out)
{
    Scene& s = *(Scene*)h;
    const float* t = which == 0 ? s.mie : (which == 1 ? s.choppedMie : s.choppedMieIntegral);
    for (int i = 0; i < n; i++) out[i] = tex1d(t, u[i]);
}

/* deterministic math probes */
void orc_math_probe(int fn, const float* x, int n, float* out, float* out2)
{
    for (int i = 0; i < n; i++) {
        switch (fn) {
        case 0: out[i] = ds_expf(x[i]); break;
        case 1: out[i] = ds_logf(x[i]); break;
        case 2: ds_sincosf(x[i], &out[i], &out2[i]); break;
        case 3: out[i] = ds_log2f(x[i]); break;
        default: out[i] = ds_exp2f(x[i]); break;
        }
    }
}

/* RNG probes: seeds[i] = tea4(val0[i], stream[i]); then `draws` rnd() values each */
void orc_rng_probe(const uint32_t* val0, const uint32_t* stream, int n, int draws, uint32_t* seeds, float* out)
{
    for (int i = 0; i < n; i++) {
        uint32_t s = tea4(val0[i], stream[i]);
        seeds[i] = s;
        for (int d = 0; d < draws; d++) out[(size_t)i * draws + d] = rnd(s);
    }
}

/* getNewDirection probe: out[i] = direction sampled around prev[i] with seed tea4(val0[i], stream[i]) */
void orc_new_directions(void* h, const uint32_t* val0, const uint32_t* stream, const float* prev, int n, float* out)
{
    const Scene& s = *(Scene*)h;
    for (int i = 0; i < n; i++) {
        uint32_t seed = tea4(val0[i], stream[i]);
        const f3 d = getNewDirection(s, seed, mk(prev[3 * i], prev[3 * i + 1], prev[3 * i + 2]));
        out[3 * i] = d.x;
        out[3 * i + 1] = d.y;
        out[3 * i + 2] = d.z;
    }
}

/* DG/Util/sutil.cpp:501-524 with fov_is_vertical = false (Camera.cpp:109-111) */
void orc_camera_look_at(const float* eye, const float* lookat, const float* up, float hfovDeg, float aspect, float* cam)
{
    const f3 e = mk(eye[0], eye[1], eye[2]);
    f3 W = mk(lookat[0], lookat[1], lookat[2]) - e;
    const float wlen = length(W);
    f3 U = normalize(cross(W, mk(up[0], up[1], up[2])));
    f3 V = normalize(cross(U, W));
    const float ulen = wlen * tanf(0.5f * hfovDeg * PI_F / 180.0f);
    U = U * ulen;
    const float vlen = ulen / aspect;
    V = V * vlen;
    const float o[12] = {e.x, e.y, e.z, U.x, U.y, U.z, V.x, V.y, V.z, W.x, W.y, W.z};
    memcpy(cam, o, sizeof(o));
}

/* one path per (origin, direction): out[i] = radiance rgb.  seedVal0/stream per path. */
void orc_trace_paths(void* h, int mode, int n, const float* origins, const float* dirs, const uint32_t* seedVal0,
                     const uint32_t* stream, float* out)
{
    const Scene& s = *(Scene*)h;
    for (int i = 0; i < n; i++) {
        const f3 r = traceRadiance(s, mode, mk(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                                   mk(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), seedVal0[i], stream[i]);
        out[3 * i] = r.x;
        out[3 * i + 1] = r.y;
        out[3 * i + 2] = r.z;
    }
    foldCounters();
}

/* PathTracingRenderer::render (PathTracingRenderer.cpp:21-31): one sample per pixel into
 * frameResultBuffer (float4, w = 1), stream = subframeId */
void orc_render_frame_result(void* h, const float* cam, int w, int hgt, int mode, uint32_t subframeId, float* frameResult)
{
    const Scene& s = *(Scene*)h;
#pragma omp parallel
    {
#pragma omp for schedule(dynamic, 16)
        for (int idx = 0; idx < w * hgt; idx++) {
            const uint32_t px = idx % w, py = idx / w;
            const f3 dir = cameraDirection(cam, px, py, w, hgt);
            const f3 r = traceRadiance(s, mode, mk(cam[0], cam[1], cam[2]), dir, px * 4096u + py, subframeId);
            frameResult[4 * (size_t)idx + 0] = r.x;
            frameResult[4 * (size_t)idx + 1] = r.y;
            frameResult[4 * (size_t)idx + 2] = r.z;
            frameResult[4 * (size_t)idx + 3] = 1.0f;
        }
        foldCounters();
    }
}

/* CU/progressive.cu:17-27, applied for subframeId (1-based, Camera.cpp:191) */
void orc_update_frame_result(const float* frameResult, float* progressive, float* variance, size_t nFloats, uint32_t subframeId)
{
    const float newWeight = 1.0f / (float)subframeId;
    for (size_t i = 0; i < nFloats; i++) {
        const float newResult = frameResult[i];
        const float previousMean = progressive[i];
        const float newMean = progressive[i] + (newResult - previousMean) * newWeight;
        progressive[i] = newMean;
        variance[i] = variance[i] + (newResult - previousMean) * (newResult - newMean);
    }
}

/* Camera::render loop (Camera.cpp:189-199) for subframes first..first+n-1 */
void orc_render_accumulate(void* h, const float* cam, int w, int hgt, int mode, uint32_t firstSubframe, uint32_t n,
                           float* progressive, float* variance)
{
    std::vector<float> frame((size_t)w * hgt * 4);
    for (uint32_t k = 0; k < n; k++) {
        orc_render_frame_result(h, cam, w, hgt, mode, firstSubframe + k, frame.data());
        orc_update_frame_result(frame.data(), progressive, variance, frame.size(), firstSubframe + k);
    }
}

/* Camera::isConverged (Camera.cpp:232-268): returns the number of unconverged pixels */
uint32_t orc_unconverged_pixels(const float* progressive, const float* variance, size_t nPixels, uint32_t subframeId)
{
    uint32_t bad = 0;
    for (size_t id = 0; id < nPixels; id++) {
        const float pixelRunningVariance = variance[4 * id];
        const float N = (float)subframeId;
        const float sigma = sqrtf(pixelRunningVariance / N);
        const float absoluteConfidence = 1.96f * sigma / sqrtf(N);
        const float relativeConfidence = absoluteConfidence / (progressive[4 * id] + FLT_EPSILON);
        const bool ok = relativeConfidence < 0.02f || absoluteConfidence < 1e-2f;
        if (!ok) bad++;
    }
    return bad;
}

/* CU/reinhard.cu:26-83.  Returns the average luminance.  lw == 0 gives ld / lw = 0/0 = NaN in the
 * reference (:69); optix::clamp(f, a, b) = fmaxf(a, fminf(f, b)) turns the NaN into 1, so the empty
 * background is written WHITE (255), exactly as oracle/_ref (the reference's own reinhard.cu) does. */
float orc_tonemap(const float* progressive, int w, int hgt, float exposure, uint8_t* screen)
{
    std::vector<float> columns(w);
    for (int x = 0; x < w; x++) {
        columns[x] = 0;
        for (int y = 0; y < hgt; y++) {
            const float* c = progressive + 4 * ((size_t)y * w + x);
            const float luminance = c[0] * 0.265068f + c[1] * 0.67023428f + c[2] * 0.06409157f + c[3] * 0.0f;
            columns[x] += luminance + 0.00001f;
        }
    }
    float result = 0;
    for (int i = 0; i < w; i++) result += columns[i];
    result = result / (float)((uint32_t)w * (uint32_t)hgt);
    for (size_t i = 0; i < (size_t)w * hgt; i++) {
        const float* c = progressive + 4 * i;
        const float lw = c[0] * 0.265068f + c[1] * 0.67023428f + c[2] * 0.06409157f + c[3] * 0.0f;
        float ld = lw * exposure / result;
        ld = ld / (1.f + ld);
        const float k = ld / lw;
        for (int ch = 0; ch < 3; ch++) {
            float v = fmaxf(0.f, fminf(c[ch] * k, 1.f));
            v = powf(v, 1.f / 2.2f);
            screen[4 * i + ch] = (uint8_t)(v * 255);
        }
        screen[4 * i + 3] = 255;
    }
    return result;
}

/* CU/pointGeneratorCamera.cu:20-42 + CU/cloudFirstScatterMaterial.cu:8-28.
 * Thread launchID = firstIndex + i.  generatePoints seeds tea<4>(launchID); the closest-hit
 * re-seeds tea<4>(launchID.x*4096 + launchID.y) with launchID.y = 0 on every attempt.  With
 * clock() gone the closest-hit seed would repeat for every attempt of a thread, so the
 * attempt number (1-based) is the closest-hit stream; the raygen stream is `stream`. */
void orc_generate_points(void* h, uint32_t firstIndex, uint32_t n, uint32_t stream, float* positions, float* directions)
{
    Scene& s = *(Scene*)h;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t launchID = firstIndex + i;
        uint32_t seed = tea4(launchID, stream);
        uint32_t attempt = 0;
        while (true) {
            attempt++;
            const f3 discNormal = uniformOnSphere(seed);
            const float discRadius = sqrtf(3.0f) / 2;
            const f3 position = uniformOnDisc(seed, discNormal) * discRadius;
            const f3 origin = position + discNormal * 2;
            const f3 direction = -discNormal;
            const float tHit = intersectBox(s, origin, direction);
            if (tHit < 0.0f) continue; /* miss program asserts in the reference (:56-59); the ray always hits in practice */
            f3 pos = origin + tHit * direction;
            pos = pos + 0.5f * s.bboxSize;
            const f3 dir = normalize(direction);
            uint32_t hitSeed = tea4(launchID * 4096u, attempt);
            ScatteringEvent scatter = getNextScatteringEvent(s, hitSeed, pos, dir);
            if (isInBox(s, scatter.scatterPos) && scatter.hasScattered) {
                const f3 p = scatter.scatterPos - 0.5f * s.bboxSize;
                positions[3 * i] = p.x;
                positions[3 * i + 1] = p.y;
                positions[3 * i + 2] = p.z;
                directions[3 * i] = -discNormal.x;
                directions[3 * i + 1] = -discNormal.y;
                directions[3 * i + 2] = -discNormal.z;
                break;
            }
        }
    }
}

/* CU/DisneyDescriptor.cuh:48-55 */
static float distanceToBox(const Scene& s, f3 pos, float voxelSize)
{
    f3 dist = pos - s.bboxSize * 0.5f;
    dist = mk(fabsf(dist.x), fabsf(dist.y), fabsf(dist.z));
    const f3 c = s.bboxSize * 0.5f - mk(voxelSize, voxelSize, voxelSize) * 0.5f;
    const f3 boxCorner = mk(fmaxf(c.x, 0.0f), fmaxf(c.y, 0.0f), fmaxf(c.z, 0.0f));
    dist = dist - boxCorner;
    dist = mk(fmaxf(dist.x, 0.0f), fmaxf(dist.y, 0.0f), fmaxf(dist.z, 0.0f));
    return length(dist);
}

/* CU/DisneyDescriptor.cuh:72-112 for TElement = uint8_t (asFloat = 0) or float (asFloat = 1,
 * the `density` members of DisneyNetworkInput; `angle` is not written by this function).
 * Also emits, per tap, the integer voxel address of the floor corner at the floor LOD
 * (x, y, z, level) so that index parity can be asserted bit-exactly. */
void orc_descriptors(void* h, const float* positions, const float* directions, int n, int asFloat, void* out, int32_t* tapIndex)
{
    Scene& s = *(Scene*)h;
    for (int i = 0; i < n; i++) {
        const f3 worldPos = mk(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
        const f3 viewDirection = mk(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
        const f3 eZ = normalize(-s.lightDirection);
        const f3 eX = normalize(cross(eZ, viewDirection));
        const f3 eY = cross(eX, eZ);
        const f3 origin = worldPos + 0.5f * s.bboxSize;
        float scale = 0.5f / s.densityMultiplier;
        float mipmapLevel = -ds_log2f(s.voxelSizeInTermsOfFreePath) - 1;
        for (int layerId = 0; layerId < 10; layerId++) {
            uint32_t sampleId = 0;
            const float mipVoxelSize = ds_exp2f(mipmapLevel) * s.voxelSizeInMeters / s.cloudSizeInMeters;
            for (int z = -2; z <= 6; z++)
                for (int y = -2; y <= 2; y++)
                    for (int x = -2; x <= 2; x++) {
                        const f3 offset = (eX * (float)x + eY * (float)y + eZ * (float)z) * scale;
                        const f3 pos = origin + offset;
                        const f3 uvw = pos * s.textureScale;
                        const float lod = fmaxf(0.0f, mipmapLevel);
                        float density = tex3dLod(s, uvw.x, uvw.y, uvw.z, lod);
                        const float distance = distanceToBox(s, pos, mipVoxelSize);
                        const float t = saturatef(distance / mipVoxelSize);
                        density = lerpf(density, 0, t);
                        const size_t o = (size_t)i * 2250 + (size_t)layerId * 225 + sampleId;
                        if (asFloat)
                            ((float*)out)[o] = density;
                        else
                            ((uint8_t*)out)[o] = (uint8_t)(density * 255.0f);
                        if (tapIndex) {
                            const int last = (int)s.levels.size() - 1;
                            const int l0 = (int)floorf(fminf(lod, (float)last));
                            const Level& L = s.levels[l0];
                            tapIndex[4 * o + 0] = (int)fminf(fmaxf(floorf(uvw.x * (float)L.nx - 0.5f), -2.0f), (float)L.nx + 1.0f);
                            tapIndex[4 * o + 1] = (int)fminf(fmaxf(floorf(uvw.y * (float)L.ny - 0.5f), -2.0f), (float)L.ny + 1.0f);
                            tapIndex[4 * o + 2] = (int)fminf(fmaxf(floorf(uvw.z * (float)L.nz - 0.5f), -2.0f), (float)L.nz + 1.0f);
                            tapIndex[4 * o + 3] = l0;
                        }
                        sampleId++;
                    }
            scale *= 2;
            mipmapLevel++;
        }
    }
}

/*
 * First half of the neural renderers' frame (DG/Scene/Cameras/DisneyRenderer.cpp:84-88, launch 0): for every pixel of a
 * rectangle of the frame, CU/disneyCamera.cu:20-36 (pinholeCamera: ray through pixel launchID + rectOrigin, angle between
 * light and view direction) and CU/disneyDescriptorMaterial.cu:14-46 (sampleDisneyDescriptor: transmittance of the whole
 * ray, a collision forced inside the cloud with xi = 1 - rnd * (1 - T), direct sun radiance with the full Mie phase, float
 * descriptor at the collision).  The seed is tea<4>(launchID.x * 4096 + launchID.y, stream) with the RECT-LOCAL launch
 * index, as in the reference; clock() -> stream.
 *   input: [rectH][rectW][10][226] floats (DisneyNetworkInput: 225 densities + angle per layer), zero where nothing scattered
 *   info:  [rectH][rectW][5] floats: radiance r, g, b; transmittance (1 where the ray misses the box); hasScattered (0 / 1)
 */
void orc_network_input(void* h, const float* cam, int frameW, int frameH, int rectX, int rectY, int rectW, int rectH, uint32_t stream, float* input,
                       float* info)
{
    Scene& s = *(Scene*)h;
    std::vector<float> layer(2250);
    for (int ly = 0; ly < rectH; ly++)
        for (int lx = 0; lx < rectW; lx++) {
            const size_t pix = (size_t)ly * rectW + lx;
            float* in = input + pix * 2260;
            float* out = info + pix * 5;
            for (int k = 0; k < 2260; k++) in[k] = 0.0f;
            out[0] = out[1] = out[2] = 0.0f;
            out[3] = 1.0f;
            out[4] = 0.0f;
            const f3 origin = mk(cam[0], cam[1], cam[2]);
            const f3 rayDirection = cameraDirection(cam, (uint32_t)(lx + rectX), (uint32_t)(ly + rectY), (uint32_t)frameW, (uint32_t)frameH);
            const float angle = acosf(dot(s.lightDirection, rayDirection)); /* disneyCamera.cu:31 */
            const float tHit = intersectBox(s, origin, rayDirection);
            if (tHit >= 0.0f) {
                f3 hitPoint = origin + tHit * rayDirection;
                hitPoint = hitPoint + 0.5f * s.bboxSize;
                const f3 direction = normalize(rayDirection);
                uint32_t seed = tea4((uint32_t)lx * 4096u + (uint32_t)ly, stream);
                const float transmittance = getNextScatteringEvent(s, rnd(seed), hitPoint, direction, false).transmittance;
                const ScatteringEvent scatter = getNextScatteringEvent(s, 1 - rnd(seed) * (1 - transmittance), hitPoint, direction);
                out[3] = transmittance;
                if (scatter.hasScattered && isInBox(s, scatter.scatterPos)) {
                    const f3 radiance = getInScattering(s, scatter, direction, false);
                    out[0] = radiance.x;
                    out[1] = radiance.y;
                    out[2] = radiance.z;
                    out[4] = 1.0f;
                    const f3 world = scatter.scatterPos - 0.5f * s.bboxSize;
                    const float p[3] = {world.x, world.y, world.z}, d[3] = {direction.x, direction.y, direction.z};
                    orc_descriptors(h, p, d, 1, 1, layer.data(), nullptr);
                    for (int l = 0; l < 10; l++)
                        for (int t = 0; t < 225; t++) in[l * 226 + t] = layer[l * 225 + t];
                }
            }
            for (int l = 0; l < 10; l++) in[l * 226 + 225] = angle; /* disneyCamera.cu:32-35: written for every pixel */
        }
}

/* CU/PointRadianceTask.h:70-77 layout, 40 bytes */
struct OrcTask {
    int32_t id;
    uint32_t experimentCount;
    float radiance;
    float runningVariance;
    float position[3];
    float direction[3];
};

/* PointRadianceTask.h:38-49 */
static void addExperimentResult(OrcTask& t, float newRadiance)
{
    t.experimentCount++;
    const float N = (float)t.experimentCount;
    const float newWeight = (float)(1.0 / N);
    const float previousMean = t.radiance;
    const float newMean = t.radiance + (newRadiance - previousMean) * newWeight;
    t.radiance = newMean;
    t.runningVariance += (newRadiance - previousMean) * (newRadiance - newMean);
}

/* PointRadianceTask.h:54-68 (the between-group variance term is ignored, as in the reference) */
static void mergeTask(OrcTask& a, const OrcTask& other)
{
    const float newWeight = other.experimentCount * 1.0f / (a.experimentCount + other.experimentCount);
    a.radiance += (other.radiance - a.radiance) * newWeight;
    a.runningVariance += other.runningVariance;
    a.experimentCount += other.experimentCount;
}

static float absoluteCI(const OrcTask& t)
{
    const float N = (float)t.experimentCount;
    const float sigma = sqrtf(t.runningVariance / N);
    return 1.96f * sigma / sqrtf(N);
}
static float relativeCI(const OrcTask& t) { return absoluteCI(t) / (t.radiance + FLT_EPSILON); }

/* RadianceCollector (DG/Scene/RadianceCollector.cpp:19-54, 73-141, 176-192) with
 * estimateEmission (CU/pointEmissionCamera.cu:20-33).  Thread t of launch frameId draws
 * seed tea<4>(t*4096 + 0, frameId).  Runs until every sample converged or maxUpdates
 * update() calls were made; returns the number of converged samples.  tasksOut[i] is the
 * merged representative of sample i (id = i); converged[i] flags it. */
int orc_point_radiance(void* h, const float* positions, const float* directions, int n, uint32_t maxThreadCount,
                       uint32_t launchesPerUpdate, uint32_t maxUpdates, OrcTask* tasksOut, uint8_t* converged,
                       uint32_t* updatesDone)
{
    const Scene& base = *(Scene*)h;
    std::vector<OrcTask> todo(n);
    for (int i = 0; i < n; i++) {
        OrcTask t{};
        t.id = i;
        memcpy(t.position, positions + 3 * i, 12);
        memcpy(t.direction, directions + 3 * i, 12);
        todo[i] = t;
    }
    for (int i = 0; i < n; i++) converged[i] = 0;
    uint32_t frameId = 0;
    int nConverged = 0;
    uint32_t updates = 0;
    std::vector<OrcTask> threads;
    while (!todo.empty() && updates < maxUpdates) {
        /* scheduleTasks (:176-192) */
        const uint32_t taskRepeatCount = maxThreadCount / (uint32_t)todo.size();
        const uint32_t threadsCount = (uint32_t)todo.size() * taskRepeatCount;
        threads.assign(threadsCount, OrcTask{});
        for (uint32_t i = 0; i < todo.size(); i++) {
            threads[i * taskRepeatCount] = todo[i];
            for (uint32_t j = 1; j < taskRepeatCount; j++) {
                OrcTask f{};
                f.id = todo[i].id;
                memcpy(f.position, todo[i].position, 12);
                memcpy(f.direction, todo[i].direction, 12);
                threads[i * taskRepeatCount + j] = f;
            }
        }
        /* update (:88-93): launchesPerUpdate launches, frameId++ before each */
        const uint32_t frame0 = frameId;
#pragma omp parallel
        {
            const Scene& s = base;
#pragma omp for schedule(dynamic, 8)
            for (int t = 0; t < (int)threadsCount; t++) {
                OrcTask& task = threads[t];
                for (uint32_t l = 1; l <= launchesPerUpdate; l++) {
                    const f3 r = traceRadiance(s, MULTIPLE_SCATTER, mk(task.position[0], task.position[1], task.position[2]),
                                               mk(task.direction[0], task.direction[1], task.direction[2]), (uint32_t)t * 4096u,
                                               frame0 + l);
                    addExperimentResult(task, r.x);
                }
            }
            foldCounters();
        }
        frameId += launchesPerUpdate;
        updates++;
        /* merge + convergence (:100-130) */
        std::vector<OrcTask> next;
        for (uint32_t i = 0; i < todo.size(); i++) {
            OrcTask& representative = threads[i * taskRepeatCount];
            for (uint32_t j = 1; j < taskRepeatCount; j++) mergeTask(representative, threads[i * taskRepeatCount + j]);
            bool isConverged = relativeCI(representative) < 2e-2f || absoluteCI(representative) < 1e-4f;
            if (representative.radiance < FLT_EPSILON) isConverged = representative.experimentCount > 100000;
            tasksOut[representative.id] = representative;
            if (isConverged) {
                converged[representative.id] = 1;
                nConverged++;
            } else {
                next.push_back(representative);
            }
        }
        todo.swap(next);
    }
    if (updatesDone) *updatesDone = updates;
    return nConverged;
}

void orc_counters_get(unsigned long long* out /* paths, events, steps */)
{
    foldCounters();
    out[0] = gPaths;
    out[1] = gEvents;
    out[2] = gSteps;
}

/* benchmarks: torch.distributed.run exports OMP_NUM_THREADS=1; the reference arm asks for all host threads explicitly */
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void orc_counters_reset()
{
    foldCounters();
    gPaths = gEvents = gSteps = 0;
}

} /* extern "C" */
