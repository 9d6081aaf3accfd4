/*
 * dsref_vec.h -- TEST INFRASTRUCTURE (part of the oracle/_ref build, never linked into the product).
 *
 * CUDA vector types and the optixu math namespace as the reference's sources use them.  The OptiX SDK 5.1.0
 * (Dependencies.md:3-4) is not vendored under /root/reference, so the arithmetic of its math header
 * (optixu/optixu_math_namespace.h) is RESTATED here from the SDK's published definitions; every function below is
 * the plain component-wise form, with the three that are not obvious called out:
 *   - float3 / float multiplies by the reciprocal   (operator/: `float inv = 1.0f / s; return a * inv;`)
 *   - normalize(v) = v * (1.0f / sqrtf(dot(v, v)))
 *   - lerp(a, b, t) = a + t * (b - a);  dot is the left-to-right sum x*x + y*y + z*z
 * The library is compiled with -ffp-contract=off, so no multiply-add is fused behind the source's back.
 *
 * One host-only device: the reference returns `const float3&` to temporaries from four inlined device functions
 * (CUDA/cloud.cuh:124,134,146,160 -- legal only because nvcc inlines them away).  g++ turns such a return into a null
 * reference, so operator*(float3, float), make_float3(float) and normalize() hand back a reference into a small ring of
 * static slots instead of a prvalue.  The values are the same; only their storage differs.
 */
#ifndef DSREF_VEC_H
#define DSREF_VEC_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct uint2 { unsigned int x, y; };
struct uint3 { unsigned int x, y, z; };
struct uchar1 { unsigned char x; };
struct uchar3 { unsigned char x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };

typedef unsigned int uint;
typedef unsigned char uchar;

namespace dsref {
inline float3& ringSlot()
{
    static float3 ring[256];
    static unsigned next = 0;
    return ring[next++ & 255u];
}
} // namespace dsref

/* ---- constructors (CUDA vector_functions.h + optixu overloads) ---- */
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float2 make_float2(float s) { return float2{s, s}; }
inline float2 make_float2(const uint2& v) { return float2{(float)v.x, (float)v.y}; }
inline float2 make_float2(const int2& v) { return float2{(float)v.x, (float)v.y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline const float3& make_float3(float s)
{
    float3& r = dsref::ringSlot();
    r.x = s; r.y = s; r.z = s;
    return r;
}
inline float3 make_float3(const uint3& v) { return float3{(float)v.x, (float)v.y, (float)v.z}; }
inline float3 make_float3(const int3& v) { return float3{(float)v.x, (float)v.y, (float)v.z}; }
inline float3 make_float3(const float4& v) { return float3{v.x, v.y, v.z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float4 make_float4(float s) { return float4{s, s, s, s}; }
inline float4 make_float4(const float3& v, float w) { return float4{v.x, v.y, v.z, w}; }
inline float4 make_float4(const float3& v) { return float4{v.x, v.y, v.z, 0.0f}; }
inline int3 make_int3(int x, int y, int z) { return int3{x, y, z}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint3 make_uint3(unsigned x, unsigned y, unsigned z) { return uint3{x, y, z}; }
inline uchar1 make_uchar1(unsigned char x) { return uchar1{x}; }
inline uchar3 make_uchar3(unsigned char x, unsigned char y, unsigned char z) { return uchar3{x, y, z}; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }

/* ---- float2 ---- */
inline float2 operator+(const float2& a, const float2& b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 operator-(const float2& a, const float2& b) { return float2{a.x - b.x, a.y - b.y}; }
inline float2 operator-(const float2& a, float b) { return float2{a.x - b, a.y - b}; }
inline float2 operator-(const float2& a) { return float2{-a.x, -a.y}; }
inline float2 operator*(const float2& a, const float2& b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 operator*(const float2& a, float s) { return float2{a.x * s, a.y * s}; }
inline float2 operator*(float s, const float2& a) { return float2{a.x * s, a.y * s}; }
inline float2 operator/(const float2& a, const float2& b) { return float2{a.x / b.x, a.y / b.y}; }
inline float2 operator/(const float2& a, float s)
{
    const float inv = 1.0f / s;
    return a * inv;
}
inline float dot(const float2& a, const float2& b) { return a.x * b.x + a.y * b.y; }
inline float length(const float2& v) { return sqrtf(dot(v, v)); }

/* ---- float3 ---- */
inline float3 operator+(const float3& a, const float3& b) { return float3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator+(const float3& a, float b) { return float3{a.x + b, a.y + b, a.z + b}; }
inline void operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline float3 operator-(const float3& a, const float3& b) { return float3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(const float3& a, float b) { return float3{a.x - b, a.y - b, a.z - b}; }
inline void operator-=(float3& a, const float3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
inline float3 operator-(const float3& a) { return float3{-a.x, -a.y, -a.z}; }
inline float3 operator*(const float3& a, const float3& b) { return float3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline const float3& operator*(const float3& a, float s)
{
    const float x = a.x * s, y = a.y * s, z = a.z * s; /* `a` may itself live in the ring */
    float3& r = dsref::ringSlot();
    r.x = x; r.y = y; r.z = z;
    return r;
}
inline const float3& operator*(float s, const float3& a) { return a * s; }
inline void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
inline void operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
inline float3 operator/(const float3& a, const float3& b) { return float3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator/(const float3& a, float s)
{
    const float inv = 1.0f / s;
    return a * inv;
}
inline float3 operator/(float s, const float3& a) { return float3{s / a.x, s / a.y, s / a.z}; }
inline void operator/=(float3& a, float s)
{
    const float inv = 1.0f / s;
    a *= inv;
}
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b)
{
    return float3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline const float3& normalize(const float3& v)
{
    const float invLen = 1.0f / sqrtf(dot(v, v));
    return v * invLen;
}
inline float3 fminf(const float3& a, const float3& b) { return float3{::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z)}; }
inline float3 fmaxf(const float3& a, const float3& b) { return float3{::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z)}; }
inline float fminf(const float3& a) { return ::fminf(::fminf(a.x, a.y), a.z); }
inline float fmaxf(const float3& a) { return ::fmaxf(::fmaxf(a.x, a.y), a.z); }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float3 lerp(const float3& a, const float3& b, float t) { return a + t * (b - a); }
inline float clamp(float f, float a, float b) { return ::fmaxf(a, ::fminf(f, b)); }
inline float3 clamp(const float3& v, float a, float b) { return float3{clamp(v.x, a, b), clamp(v.y, a, b), clamp(v.z, a, b)}; }

/* ---- float4 ---- */
inline float4 operator+(const float4& a, const float4& b) { return float4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline float4 operator-(const float4& a, const float4& b) { return float4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline float4 operator*(const float4& a, const float4& b) { return float4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline float4 operator*(const float4& a, float s) { return float4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float4 operator*(float s, const float4& a) { return a * s; }
inline float4 operator/(const float4& a, float s)
{
    const float inv = 1.0f / s;
    return a * inv;
}
inline float dot(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float4 clamp(const float4& v, float a, float b)
{
    return float4{clamp(v.x, a, b), clamp(v.y, a, b), clamp(v.z, a, b), clamp(v.w, a, b)};
}

/* ---- integer vectors ---- */
inline uint2 operator+(const uint2& a, const uint2& b) { return uint2{a.x + b.x, a.y + b.y}; }

namespace optix {
using ::float2; using ::float3; using ::float4; using ::int2; using ::int3; using ::uint2; using ::uint3;
using ::uchar1; using ::uchar3; using ::uchar4; using ::uint;
using ::make_float2; using ::make_float3; using ::make_float4; using ::make_int3; using ::make_uint2; using ::make_uint3;
using ::make_uchar1; using ::make_uchar3; using ::make_uchar4;
using ::dot; using ::cross; using ::length; using ::normalize; using ::lerp; using ::clamp; using ::fminf; using ::fmaxf;

struct size_t2 { size_t x, y; };
struct size_t3 { size_t x, y, z; };
inline size_t3 make_size_t3(size_t x, size_t y, size_t z) { return size_t3{x, y, z}; }
inline float2 make_float2(const size_t2& v) { return float2{(float)v.x, (float)v.y}; }

/* optixu_math_namespace.h: orthonormal basis around a normal */
struct Onb {
    explicit Onb(const float3& normal)
    {
        m_normal = normal;
        if (fabsf(m_normal.x) > fabsf(m_normal.z)) {
            m_binormal.x = -m_normal.y;
            m_binormal.y = m_normal.x;
            m_binormal.z = 0;
        } else {
            m_binormal.x = 0;
            m_binormal.y = -m_normal.z;
            m_binormal.z = m_normal.y;
        }
        m_binormal = normalize(m_binormal);
        m_tangent = cross(m_binormal, m_normal);
    }
    void inverse_transform(float3& p) const { p = p.x * m_tangent + p.y * m_binormal + p.z * m_normal; }
    float3 m_tangent, m_binormal, m_normal;
};

/* optixu_aabb_namespace.h (only what cloudBBox.cu:39-45 touches) */
struct Aabb {
    float3 m_min, m_max;
    void set(const float3& mn, const float3& mx) { m_min = mn; m_max = mx; }
};

#define RT_DEFAULT_MAX 1.e27f
struct Ray {
    Ray() : origin{0, 0, 0}, direction{0, 0, 0}, ray_type(0), tmin(0), tmax(RT_DEFAULT_MAX) {}
    Ray(float3 o, float3 d, unsigned int type, float tmin_, float tmax_ = RT_DEFAULT_MAX)
        : origin(o), direction(d), ray_type(type), tmin(tmin_), tmax(tmax_)
    {
    }
    float3 origin;
    float3 direction;
    unsigned int ray_type;
    float tmin;
    float tmax;
};

/* optixu_matrix_namespace.h: row-major 4x4, the operations Camera.cpp:111-118 and Arcball.cpp use */
struct Matrix4x4 {
    float m[16];
    Matrix4x4() { for (int i = 0; i < 16; i++) m[i] = 0.0f; }
    explicit Matrix4x4(const float* data) { for (int i = 0; i < 16; i++) m[i] = data[i]; }
    float& operator[](int i) { return m[i]; }
    const float& operator[](int i) const { return m[i]; }
    static Matrix4x4 identity()
    {
        Matrix4x4 r;
        r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
        return r;
    }
    /* columns are the basis vectors and the origin */
    static Matrix4x4 fromBasis(const float3& u, const float3& v, const float3& w, const float3& c)
    {
        Matrix4x4 r;
        r.m[0] = u.x; r.m[1] = v.x; r.m[2] = w.x; r.m[3] = c.x;
        r.m[4] = u.y; r.m[5] = v.y; r.m[6] = w.y; r.m[7] = c.y;
        r.m[8] = u.z; r.m[9] = v.z; r.m[10] = w.z; r.m[11] = c.z;
        r.m[12] = 0; r.m[13] = 0; r.m[14] = 0; r.m[15] = 1;
        return r;
    }
    static Matrix4x4 rotate(float radians, const float3& axis)
    {
        Matrix4x4 Mat = identity();
        float* mm = Mat.m;
        const float3 a = normalize(axis);
        const float c = cosf(radians), s = sinf(radians), t = 1.0f - c;
        const float x = a.x, y = a.y, z = a.z;
        mm[0] = t * x * x + c;      mm[1] = t * x * y - s * z;  mm[2] = t * x * z + s * y;
        mm[4] = t * x * y + s * z;  mm[5] = t * y * y + c;      mm[6] = t * y * z - s * x;
        mm[8] = t * x * z - s * y;  mm[9] = t * y * z + s * x;  mm[10] = t * z * z + c;
        return Mat;
    }
    Matrix4x4 inverse() const
    {
        /* general 4x4 inverse by cofactors (the SDK's Matrix<4,4>::inverse) */
        const float* a = m;
        float inv[16];
        inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
        inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
        inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
        inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
        inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
        inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
        inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
        inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
        inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
        inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
        inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
        inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
        inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
        inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
        inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
        inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
        const float det = 1.0f / (a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12]);
        Matrix4x4 r;
        for (int i = 0; i < 16; i++) r.m[i] = inv[i] * det;
        return r;
    }
};
inline Matrix4x4 operator*(const Matrix4x4& a, const Matrix4x4& b)
{
    Matrix4x4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a.m[i * 4 + k] * b.m[k * 4 + j];
            r.m[i * 4 + j] = s;
        }
    return r;
}
inline float4 operator*(const Matrix4x4& a, const float4& v)
{
    float4 r;
    r.x = a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z + a.m[3] * v.w;
    r.y = a.m[4] * v.x + a.m[5] * v.y + a.m[6] * v.z + a.m[7] * v.w;
    r.z = a.m[8] * v.x + a.m[9] * v.y + a.m[10] * v.z + a.m[11] * v.w;
    r.w = a.m[12] * v.x + a.m[13] * v.y + a.m[14] * v.z + a.m[15] * v.w;
    return r;
}
} // namespace optix

#ifndef M_PIf
#define M_PIf 3.14159265358979323846f
#endif
#ifndef CUDART_PI_F
#define CUDART_PI_F 3.141592654f
#endif

#endif /* DSREF_VEC_H */
