/*
 * ds_synth.h -- integer-only procedural cloud grids (synthetic bench/test input).
 *
 * The reference ships no cloud (.vdb files are git-ignored), so
 * the benchmark configs run on procedural u8 density grids.  The generator is
 * pure 32/64-bit integer arithmetic (hash lattice + fixed-point trilinear
 * value noise), so the host build and the sm_100a kernel produce the same
 * bytes without any floating point.  Grids keep a one-voxel zero border, the
 * way the reference importer pads the active bounding box by one voxel
 * (Util/Resources.cpp:97-101).
 *
 * kind 0 "cumulus": ellipsoid (semi-axes 0.42, 0.30, 0.36 of N) with fbm edge
 * kind 1 "cube":    fbm noise masked to a centred cube of side 0.8 N
 * kind 2 "slab":    homogeneous slab, density 255 for 0.25N <= z < 0.75N
 *                   (analytic Beer-Lambert checks)
 */
#ifndef DS_SYNTH_H
#define DS_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define DS_SYNTH_HD __host__ __device__ __forceinline__
#else
#define DS_SYNTH_HD static inline
#endif

DS_SYNTH_HD uint32_t ds_synth_hash32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

/* 16-bit lattice value */
DS_SYNTH_HD uint32_t ds_synth_lattice(uint32_t ix, uint32_t iy, uint32_t iz, uint32_t seed)
{
    return ds_synth_hash32((ix * 0x8da6b343u) ^ (iy * 0xd8163841u) ^ (iz * 0xcb1ab31fu) ^ seed) >> 16;
}

/* trilinear value noise, coordinates in 16.16 fixed point, result 16 bit */
DS_SYNTH_HD uint32_t ds_synth_noise(uint32_t X, uint32_t Y, uint32_t Z, uint32_t seed)
{
    const uint32_t ix = X >> 16, iy = Y >> 16, iz = Z >> 16;
    const uint64_t fx = (X >> 8) & 0xffu, fy = (Y >> 8) & 0xffu, fz = (Z >> 8) & 0xffu;
    uint64_t acc = 0;
    for (uint32_t c = 0; c < 8; ++c) {
        const uint32_t dx = c & 1u, dy = (c >> 1) & 1u, dz = (c >> 2) & 1u;
        const uint64_t w = (dx ? fx : 256u - fx) * (dy ? fy : 256u - fy) * (dz ? fz : 256u - fz);
        acc += w * ds_synth_lattice(ix + dx, iy + dy, iz + dz, seed);
    }
    return (uint32_t)(acc >> 24);
}

/* 4-octave fbm, result in [0, 65535] */
DS_SYNTH_HD uint32_t ds_synth_fbm4(uint32_t X, uint32_t Y, uint32_t Z, uint32_t seed)
{
    uint32_t sum = 0;
    for (uint32_t o = 0; o < 4; ++o) {
        sum += ds_synth_noise(X << o, Y << o, Z << o, seed + o * 0x9e3779b9u) >> (o + 1);
    }
    /* octave weights sum to 15/16 */
    return (sum * 16u) / 15u;
}

DS_SYNTH_HD uint8_t ds_synth_voxel(int kind, uint32_t seed, int n, int x, int y, int z)
{
    if (x <= 0 || y <= 0 || z <= 0 || x >= n - 1 || y >= n - 1 || z >= n - 1) return 0;
    if (kind == 2) {
        return (4 * z >= n && 4 * z < 3 * n) ? 255 : 0;
    }
    const uint32_t X = (uint32_t)(((uint64_t)x * 6u << 16) / (uint32_t)n);
    const uint32_t Y = (uint32_t)(((uint64_t)y * 6u << 16) / (uint32_t)n);
    const uint32_t Z = (uint32_t)(((uint64_t)z * 6u << 16) / (uint32_t)n);
    const int64_t fbm = (int64_t)ds_synth_fbm4(X, Y, Z, seed);
    const int64_t dx = 2 * x + 1 - n, dy = 2 * y + 1 - n, dz = 2 * z + 1 - n; /* doubled offsets from centre */
    if (kind == 1) {
        const int64_t half = (int64_t)n * 8 / 10; /* doubled half side = 0.8 N */
        if (dx < -half || dx > half || dy < -half || dy > half || dz < -half || dz > half) return 0;
        return (uint8_t)(fbm >> 8);
    }
    const int64_t ax = (int64_t)n * 84 / 100, ay = (int64_t)n * 60 / 100, az = (int64_t)n * 72 / 100;
    const int64_t q16 = ((dx * dx) << 16) / (ax * ax) + ((dy * dy) << 16) / (ay * ay) + ((dz * dz) << 16) / (az * az);
    int64_t d16 = ((65536 - q16) * 8) / 5 + ((fbm - 32768) * 9) / 10;
    if (d16 < 0) d16 = 0;
    if (d16 > 65535) d16 = 65535;
    return (uint8_t)(d16 >> 8);
}

#endif /* DS_SYNTH_H */
