/*
 * ds_kernels_fast.cu -- translation unit compiled with FMA contraction: the throughput flavour of the
 * estimator (DsPrecision FAST).  Same algorithm and the same per-path RNG streams as the reference
 * estimator (CU/cloudRadianceMaterials.cu, CU/cloud.cuh), evaluated the way the reference itself runs on a
 * GPU -- hardware trilinear texture filtering (rtTex3D, cloud.cuh:61) and approximate transcendentals
 * (--use_fast_math, vcxproj:320) -- and validated against the oracle statistically.
 *
 * k_trace_fast differs from the generic kernel (ds_kernels.cuh) only in how it spends instructions:
 *   - the path marches in TEXTURE space (q = pos * textureScale), so a tap is one TEX instruction;
 *   - transmittance is carried as optical depth: tau += sigma*step and the collision test xi > exp(-tau)
 *     becomes tau > -ln(xi); one logarithm per free flight instead of one exponential per march step;
 *   - empty space: march steps whose trilinear footprint lies in all-zero occupancy cells read density 0 and
 *     change nothing (tau, the collision test); they are skipped in bulk.  A lane in an empty cell reads the
 *     cell's Chebyshev distance d to the nearest occupied cell, leaves the cube of (2d-1)^3 empty cells through
 *     the face its ray exits ("leap DDA") and repeats until the next cell is occupied or the grid ends; it then
 *     advances the whole number of march steps that fit.  The step counter advances by the same amount: the
 *     reference algorithm performs those steps;
 *   - primary rays do not depend on the subframe (no pixel jitter, CU/cameraCommon.cuh:22), so the empty-space
 *     leg from the camera to the first occupied cell is walked ONCE per pixel by k_primary_prepass and cached:
 *     pixels that never reach an occupied cell are not traced at all (their sample is exactly 0), the others
 *     start at the cloud surface;
 *   - the 16-step bisection of the chopped-Mie CDF (cloud.cuh:167-178) is replaced by the closed-form
 *     inverse of the same piecewise-linear CDF: guide table -> short binary search for the table cell ->
 *     linear solve.  The bisection converges to that root within 2^-16;
 *   - radiance is accumulated as the scalar sum of Tsun*phase and scaled by lightColor*lightIntensity*ratio
 *     once per path.
 */
#include "ds_kernels.cuh"

namespace dsk {

/* the EXACT instantiations live in ds_kernels_exact.cu (see there) */
extern template struct KernelSet<false>;

struct FastState {
    V3 q0;    /* origin of the current free flight, texture coordinates */
    V3 sv;    /* march step in texture space: dir * textureScale * sampleStep */
    float nf; /* march steps taken in this flight; the marching position is q0 + nf * sv */
    float rad;
    float pendT, pendP; /* sun transmittance and phase of the latest event: rad += pendT * pendP is applied at the NEXT event (or
                           when the sample is written), so the texture fetch behind pendT has a whole march phase to land */
    float tau;     /* optical depth accumulated in the current free flight */
    float tauStar; /* -ln(xi) */
    uint32_t seed;
    int depth;
    uint32_t out;
};

constexpr int FAST_MAX_THREADS = 1024; /* 64 registers: one block of 1024 threads or two of 512 per SM */

struct FastConsts {
    V3 stepTs;    /* sampleStep * textureScale */
    V3 invStepTs; /* 1 / stepTs: dir = sv * invStepTs */
    V3 half;      /* 0.5 + 0.01 * textureScale: half extent of the in-box slab in texture space */
    V3 lightTs;   /* lightDirection * invStepTs: dot(light, dir) = dot(lightTs, sv) */
    float c1;     /* densityMultiplier * sampleStep */
    float invStep;
    float nxf, nyf, nzf;
    float backScale; /* 1 / (densityMultiplier * sampleStep): march steps per unit of optical depth at density 1 */
};

/* computed once on the host and passed as a kernel parameter: the values sit in the constant bank and are used as
 * instruction operands, instead of being re-derived (and kept in registers) by every thread */
static FastConsts makeConsts(const DevScene& sc)
{
    FastConsts k;
    k.stepTs = V3{sc.texScale.x * sc.step, sc.texScale.y * sc.step, sc.texScale.z * sc.step};
    k.invStepTs = V3{1.0f / k.stepTs.x, 1.0f / k.stepTs.y, 1.0f / k.stepTs.z};
    k.half = V3{0.5f + 0.01f * sc.texScale.x, 0.5f + 0.01f * sc.texScale.y, 0.5f + 0.01f * sc.texScale.z};
    k.lightTs = V3{sc.light.x * k.invStepTs.x, sc.light.y * k.invStepTs.y, sc.light.z * k.invStepTs.z};
    k.c1 = sc.mult * sc.step;
    k.invStep = 1.0f / sc.step;
    k.nxf = (float)sc.nx;
    k.nyf = (float)sc.ny;
    k.nzf = (float)sc.nz;
    k.backScale = 1.0f / (sc.mult * sc.step);
    return k;
}

__device__ __forceinline__ bool inBoxTs(const FastConsts& k, V3 q)
{
    return fabsf(q.x - 0.5f) <= k.half.x && fabsf(q.y - 0.5f) <= k.half.y && fabsf(q.z - 0.5f) <= k.half.z;
}

/* marching position after n steps of the current flight.  Positions are a function of the step INDEX, never of
 * how the steps were grouped into march iterations and leaps, so a path's arithmetic does not depend on the
 * warp schedule. */
__device__ __forceinline__ V3 posAt(const FastState& s, float n)
{
    return mk(fmaf(n, s.sv.x, s.q0.x), fmaf(n, s.sv.y, s.q0.y), fmaf(n, s.sv.z, s.q0.z));
}

/*
 * Closed-form inverse of the piecewise-linear CDF that cloud.cuh:167-178 bisects.  sCdfPad[i + 1] = cdf[i], sCdfPad[0] = 0 (the value
 * "left of" knot 0), +inf entries at the end.  The two-level guide (DevScene) gives the first knot `lo` that can be the answer; at most
 * GUIDE_MAX_KNOTS = 8 knots of the bucket lie below val, and the CDF is non-decreasing, so four fixed probes (4, 2, 1, 1) count them: no
 * loop, no divergence, and no second guide entry to bound the search (a probe past the bucket reads a knot >= val and fails).
 */
__device__ __forceinline__ float invertCdf(const float* sCdfPad, const uint16_t* sGuideA, const uint16_t* sGuideB, float val)
{
    const bool low = val < GUIDE_B_LIMIT;
    const int kb = (int)(val * (low ? (float)GUIDE_B_N / GUIDE_B_LIMIT : (float)GUIDE_A_N));
    int i = (int)(low ? sGuideB[kb] : sGuideA[kb]); /* first index with cdf[i] >= val is in [i, i + 8] */
    i += sCdfPad[i + 4] < val ? 4 : 0;              /* sCdfPad[i + s] = cdf[i + s - 1]: the s knots from i on are all below val */
    i += sCdfPad[i + 2] < val ? 2 : 0;
    i += sCdfPad[i + 1] < val ? 1 : 0;
    i += sCdfPad[i + 1] < val ? 1 : 0;
    const float a = sCdfPad[i], b = sCdfPad[i + 1]; /* cdf[i - 1] < val <= cdf[i] */
    const float t = __fdividef(val - a, b - a);
    /* tex1D clamps below the first texel centre: every val <= cdf[0] bisects to u = 0 */
    const float u = i == 0 ? 0.0f : ((float)i - 0.5f + t) * (1.0f / (float)MIE_N);
    return 2.0f * fminf(u, 1.0f) - 1.0f;
}

/* CU/cloud.cuh:160-188 */
__device__ __forceinline__ V3 newDirectionFast(const float* sCdfPad, const uint16_t* sGuideA, const uint16_t* sGuideB, uint32_t& seed, V3 prev)
{
    const float val = rnd(seed);
    const float cosTheta = invertCdf(sCdfPad, sGuideA, sGuideB, val);
    const float phi = rnd(seed) * (PI_F * 2.0f);
    const float s2 = fmaxf(fmaf(-cosTheta, cosTheta, 1.0f), 0.0f);
    const float sinTheta = s2 * rsqrtf(fmaxf(s2, 1.0e-30f));
    float s, c;
    __sincosf(phi, &s, &c);
    /* The reference normalises the result (cloud.cuh:185); here it is left as it comes out of the frame: a unit vector rotated by a frame
     * built around a unit axis is unit up to the rounding of rsqrt / sincos (a few 1e-7 per event, a random walk of ~2e-5 over the 2000
     * events a path may have), far inside the 9-bit weights of the texture filter.  Eight instructions per event (+1.2 %); the paths keep
     * following the oracle's event by event (test_c2_grid_fast_flavour_against_the_oracle needs that correlation).  A branch-free frame
     * (Duff et al. 2017) saves eight more (+1.3 %) but decorrelates the paths from the oracle's: not taken. */
    return onbInverseTransform<true>(prev, mk(sinTheta * c, sinTheta * s, cosTheta));
}

/* 4096-entry table, linear, clamp-to-edge, u in [0, 1] (tex1D semantics of the Mie samplers) */
__device__ __forceinline__ float tableLerp(const float* table, float u)
{
    const float x = fminf(fmaxf(fmaf(u, (float)MIE_N, -0.5f), 0.0f), (float)(MIE_N - 1));
    const int i = (int)x;
    const float f = x - (float)i;
    const float a = table[i], b = table[min(i + 1, MIE_N - 1)];
    return fmaf(f, b - a, a);
}
/* the same on a table of IEEE halves (the chopped phase sampler in shared memory: 8 instead of 16 KiB; 2^-11 relative rounding per entry) */
__device__ __forceinline__ float tableLerp(const __half* table, float u)
{
    const float x = fminf(fmaxf(fmaf(u, (float)MIE_N, -0.5f), 0.0f), (float)(MIE_N - 1));
    const int i = (int)x;
    const float f = x - (float)i;
    const float a = __half2float(table[i]), b = __half2float(table[min(i + 1, MIE_N - 1)]);
    return fmaf(f, b - a, a);
}

/* one trilinear fetch of a u8 volume at texture coordinate q: hardware 3-D filtering of the block-linear R8 array,
 * what rtTex3D does (cloud.cuh:61) */
__device__ __forceinline__ float tapVolume(cudaTextureObject_t tex3, V3 q) { return tex3D<float>(tex3, q.x, q.y, q.z); }

/* FUSED: density and sun transmittance interleaved in ONE block-linear RG8 array (texel = {density, transmittance}; DevScene::fusedTex).
 * The filter weights depend on the coordinate alone, so each channel filters to exactly the value the separate R8 array gives; what
 * changes is where the bytes live: the sun tap of an event (at the collision point, within one step of the march tap that collided)
 * reads the sectors that march tap has just brought in, instead of four cold sectors of a second volume. */
template <bool FUSED>
__device__ __forceinline__ float tapDensity(const DevScene& sc, V3 q)
{
    if (FUSED) return tex3D<float2>(sc.fusedTex, q.x, q.y, q.z).x;
    return tex3D<float>(sc.densityTex, q.x, q.y, q.z);
}
template <bool FUSED>
__device__ __forceinline__ float tapSun(const DevScene& sc, V3 q)
{
    if (FUSED) return tex3D<float2>(sc.fusedTex, q.x, q.y, q.z).y;
    return tex3D<float>(sc.inscatterTex, q.x, q.y, q.z);
}

/* PIPE: fetch the densities of the next two march steps (base + 1, base + 2) of lane state `s`.  The second tap is speculative -- it
 * is wasted when the first step already collides -- so it is only issued when a collision at the first step is unlikely: the
 * optical depth one step adds at the density last seen (the field is trilinear, one step is about one voxel) is compared with what
 * is left to the collision threshold, both known BEFORE the fetch.  A skipped second tap is marked by pd2 < 0 and fetched by the
 * march loop if the first step did not collide after all.  Which taps are fetched never changes a path's arithmetic: positions are
 * functions of the step index alone. */
#define DS_ISSUE_TAPS(base, lastD)                                                                    \
    do {                                                                                              \
        pd1 = tapDensity<FUSED>(sc, posAt(s, (base) + 1.0f));                                      \
        if ((lastD) * specC1 > s.tauStar - s.tau) {                                                   \
            pd2 = -1.0f;                                                                              \
            nTaps += 1u;                                                                              \
        } else {                                                                                      \
            pd2 = tapDensity<FUSED>(sc, posAt(s, (base) + 2.0f));                                  \
            nTaps += 2u;                                                                              \
        }                                                                                             \
    } while (0)

/* start of a free flight (cloudRadianceMaterials.cu:28-35, cloud.cuh:120); the flight origin is s.q0 */
template <bool CHECK_BOX>
__device__ __forceinline__ bool beginFlight(const FastConsts& k, FastState& s)
{
    if (CHECK_BOX && !inBoxTs(k, s.q0)) return false;
    s.depth++;
    if (s.depth == MAX_DEPTH) return false;
    const float xi = rnd(s.seed);
    s.tauStar = -__logf(xi); /* xi == 0 -> +inf: never collides, as `0 > T` in the reference */
    s.tau = 0.0f;
    s.nf = 0.0f;
    return true;
}

/* occupancy cell of the trilinear footprint at texture coordinate q; false if the cell is occupied */
__device__ __forceinline__ bool tapCellEmpty(const DevScene& sc, const FastConsts& k, const uint32_t* occ, V3 q)
{
    const int fx = __float2int_rd(fmaf(q.x, k.nxf, -0.5f)), fy = __float2int_rd(fmaf(q.y, k.nyf, -0.5f)), fz = __float2int_rd(fmaf(q.z, k.nzf, -0.5f));
    const int cx = min(max(fx, 0), sc.nx - 1) >> sc.occShift;
    const int cy = min(max(fy, 0), sc.ny - 1) >> sc.occShift;
    const int cz = min(max(fz, 0), sc.nz - 1) >> sc.occShift;
    const int cell = (cz * sc.ocy + cy) * sc.ocx + cx;
    return ((occ[cell >> 5] >> (cell & 31)) & 1u) == 0u;
}

/* Chebyshev distance (cells) from the tap cell of q to the nearest occupied cell; 0 = the cell is occupied */
__device__ __forceinline__ int tapCellDistance(const DevScene& sc, const FastConsts& k, V3 q)
{
    const int fx = __float2int_rd(fmaf(q.x, k.nxf, -0.5f)), fy = __float2int_rd(fmaf(q.y, k.nyf, -0.5f)), fz = __float2int_rd(fmaf(q.z, k.nzf, -0.5f));
    const int cx = min(max(fx, 0), sc.nx - 1) >> sc.occShift;
    const int cy = min(max(fy, 0), sc.ny - 1) >> sc.occShift;
    const int cz = min(max(fz, 0), sc.nz - 1) >> sc.occShift;
    return (int)__ldg(sc.cellDist + (cz * sc.ocy + cy) * sc.ocx + cx);
}

/* whole march steps until the position leaves the in-box slab (the reference's `while (isInBox(pos))`) */
__device__ __forceinline__ float stepsToLeaveBox(const FastConsts& k, V3 q, V3 sv)
{
    const float dx = sv.x, dy = sv.y, dz = sv.z;
    const float big = 1.0e30f;
    const float tx = fabsf(dx) > 1e-12f ? __fdividef((dx > 0.0f ? 0.5f + k.half.x : 0.5f - k.half.x) - q.x, dx) : big;
    const float ty = fabsf(dy) > 1e-12f ? __fdividef((dy > 0.0f ? 0.5f + k.half.y : 0.5f - k.half.y) - q.y, dy) : big;
    const float tz = fabsf(dz) > 1e-12f ? __fdividef((dz > 0.0f ? 0.5f + k.half.z : 0.5f - k.half.z) - q.z, dz) : big;
    return fmaxf(floorf(fminf(fminf(tx, ty), fminf(tz, 65534.0f))) + 1.0f, 0.0f);
}

/* The footprint at q is outside the grid and all face voxels are zero: every tap reads 0 until the ray enters the
 * region floor(x) in [0, N-2] (slab test), or until it leaves the box if it never does. */
__device__ __forceinline__ float outsideGridSteps(const FastConsts& k, V3 q, V3 sv, float x, float y, float z, bool& leaves)
{
    const float vx = sv.x * k.nxf, vy = sv.y * k.nyf, vz = sv.z * k.nzf;
    const float big = 1.0e30f;
    float tn = -big, tf = big, margin = 0.0f;
    const float v[3] = {vx, vy, vz}, p[3] = {x, y, z}, hi[3] = {k.nxf - 1.0f, k.nyf - 1.0f, k.nzf - 1.0f};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (fabsf(v[a]) > 1e-9f) {
            const float inv = __fdividef(1.0f, v[a]);
            const float t0 = (0.0f - p[a]) * inv, t1 = (hi[a] - p[a]) * inv;
            const float lo_ = fminf(t0, t1), hi_ = fmaxf(t0, t1);
            if (lo_ > tn) {
                tn = lo_;
                margin = fabsf(inv);
            }
            tf = fminf(tf, hi_);
        } else if (p[a] < 0.0f || p[a] >= hi[a]) {
            tn = big; /* parallel to the slab and outside it */
        }
    }
    if (tn < tf && tf > 0.0f) return fmaxf(floorf(fminf(tn - 0.01f * margin, 65535.0f)), 0.0f); /* enters the grid */
    leaves = true; /* never enters the grid: the caller adds the steps to the box exit (stepsToLeaveBox) */
    return 0.0f;
}

/*
 * Leap DDA.  Walks the ray from the tap at q through EMPTY cells and returns the whole number of further march
 * steps whose taps are all guaranteed to read 0 (0 when the tap cell is occupied); `more` is set when the walk was
 * cut short after `maxLeaps` leaps and the landing position is still in an empty cell.
 * Coordinates: x = u*N - 0.5 is the voxel coordinate whose floor is the low corner of the trilinear footprint;
 * the occupancy of cell c covers voxels [c*2^s, c*2^s + 2^s], i.e. every footprint with floor(x) in c.
 * A cell at Chebyshev distance d >= 1 from the nearest occupied cell is the centre of a cube of (2d-1)^3 empty
 * cells; the ray leaves that cube through one face, lands in the adjacent cell and repeats.  When the ray leaves
 * the grid and all face voxels are zero (borderEmpty), the rest of its way out of the box reads 0 as well.
 */
__device__ __forceinline__ float emptySteps(const DevScene& sc, const FastConsts& k, V3 q, V3 sv, int maxLeaps, bool& more, bool& leaves)
{
    more = false;
    leaves = false; /* set when every tap from here to the box exit reads 0: the caller then takes stepsToLeaveBox(k, q, sv) steps (kept out of
                       this function so that its code exists once per call site, not once per way of leaving) */
    const float x = fmaf(q.x, k.nxf, -0.5f), y = fmaf(q.y, k.nyf, -0.5f), z = fmaf(q.z, k.nzf, -0.5f);
    const int fx = __float2int_rd(x), fy = __float2int_rd(y), fz = __float2int_rd(z);
    /* outside the grid the footprint is clamped to edge voxels */
    if ((unsigned)fx >= (unsigned)(sc.nx - 1) || (unsigned)fy >= (unsigned)(sc.ny - 1) || (unsigned)fz >= (unsigned)(sc.nz - 1))
        return sc.borderEmpty ? outsideGridSteps(k, q, sv, x, y, z, leaves) : 0.0f;
    const int sh = sc.occShift;
    int cx = fx >> sh, cy = fy >> sh, cz = fz >> sh;
    int dist = (int)__ldg(sc.cellDist + (cz * sc.ocy + cy) * sc.ocx + cx);
    if (dist == 0) return 0.0f; /* the tap cell is occupied */
    if (sc.cellEscape) {
        const int oct = (sv.x > 0.0f ? 1 : 0) | (sv.y > 0.0f ? 2 : 0) | (sv.z > 0.0f ? 4 : 0);
        if ((__ldg(sc.cellEscape + (cz * sc.ocy + cy) * sc.ocx + cx) >> oct) & 1) {
            leaves = true; /* nothing but empty cells ahead */
            return 0.0f;
        }
    }
    const float cs = (float)(1 << sh);
    const float vx = sv.x * k.nxf, vy = sv.y * k.nyf, vz = sv.z * k.nzf; /* voxels per step */
    const float big = 1.0e30f;
    const bool mx = fabsf(vx) > 1e-9f, my = fabsf(vy) > 1e-9f, mz = fabsf(vz) > 1e-9f;
    const float ix = mx ? __fdividef(1.0f, vx) : 0.0f, iy = my ? __fdividef(1.0f, vy) : 0.0f, iz = mz ? __fdividef(1.0f, vz) : 0.0f;
    const bool px = vx > 0.0f, py = vy > 0.0f, pz = vz > 0.0f;
    /* A cube that reaches beyond the grid is empty there as well: taps outside the grid read the face voxels of cells the cube covers
     * (clamp addressing).  So the exit planes need no clamping to the grid; the walk has left the grid when the cell it would enter does
     * not exist. */
    float t = 0.0f, margin = 0.0f;
    bool leftGrid = false, blocked = false;
#pragma unroll 1
    for (int it = 0; it < maxLeaps; ++it) {
        const int r = dist - 1; /* cells [c-r, c+r]^3 are empty */
        /* exit planes of the cube along the direction of travel */
        const float ex = (float)(px ? cx + r + 1 : cx - r) * cs, ey = (float)(py ? cy + r + 1 : cy - r) * cs, ez = (float)(pz ? cz + r + 1 : cz - r) * cs;
        const float tx = mx ? (ex - x) * ix : big, ty = my ? (ey - y) * iy : big, tz = mz ? (ez - z) * iz : big;
        t = fminf(tx, fminf(ty, tz));
        /* cell the ray enters: the exit axis moves one cell past the cube face, the others follow the ray */
        const bool ax = tx <= ty && tx <= tz, ay = !ax && ty <= tz;
        int nx_ = __float2int_rd(fmaf(t, vx, x)) >> sh, ny_ = __float2int_rd(fmaf(t, vy, y)) >> sh, nz_ = __float2int_rd(fmaf(t, vz, z)) >> sh;
        nx_ = min(max(nx_, cx - r), cx + r);
        ny_ = min(max(ny_, cy - r), cy + r);
        nz_ = min(max(nz_, cz - r), cz + r);
        const bool az = !ax && !ay;
        nx_ = ax ? (px ? cx + r + 1 : cx - r - 1) : nx_;
        ny_ = ay ? (py ? cy + r + 1 : cy - r - 1) : ny_;
        nz_ = az ? (pz ? cz + r + 1 : cz - r - 1) : nz_;
        margin = fabsf(ax ? ix : (ay ? iy : iz));
        leftGrid = (unsigned)nx_ >= (unsigned)sc.ocx || (unsigned)ny_ >= (unsigned)sc.ocy || (unsigned)nz_ >= (unsigned)sc.ocz;
        if (leftGrid) break;
        cx = nx_;
        cy = ny_;
        cz = nz_;
        dist = (int)__ldg(sc.cellDist + (cz * sc.ocy + cy) * sc.ocx + cx);
        blocked = dist == 0;
        if (blocked) break;
    }
    if (leftGrid && sc.borderEmpty) {
        leaves = true;
        return 0.0f;
    }
    more = !leftGrid && !blocked;
    /* stay 0.01 voxel short of the plane that stopped the walk */
    return fmaxf(floorf(fminf(t - 0.01f * margin, 65535.0f)), 0.0f);
}

/* ray of work item `idx` (CU/pathTracingCamera.cu:12-21, CU/cameraCommon.cuh:19-29, CU/pointEmissionCamera.cu:20-33) */
__device__ __forceinline__ bool itemRay(const TraceJob& job, unsigned long long idx, V3& o, V3& d, uint32_t& val0, uint32_t& stream,
                                        uint32_t& out, uint32_t& pixel)
{
    pixel = 0;
    if (job.kind == JOB_RENDER) {
        unsigned long long sub;
        uint32_t px, py;
        if (job.hitList) {
            if (job.regionSize) {
                const unsigned long long perRegion = (unsigned long long)job.regionSize * job.nSub;
                const uint32_t region = (uint32_t)(idx / perRegion);
                const uint32_t rem = (uint32_t)(idx - (unsigned long long)region * perRegion);
                const uint32_t first = region * job.regionSize;
                const uint32_t size = min(job.regionSize, job.nHit - first); /* the last run may be short */
                const uint32_t s32 = rem / size;
                sub = s32;
                pixel = job.hitList[first + (rem - s32 * size)];
            } else {
                sub = idx / job.nHit;
                pixel = job.hitList[(uint32_t)(idx - sub * job.nHit)];
            }
            px = pixel % (uint32_t)job.width;
            py = pixel / (uint32_t)job.width;
        } else {
            sub = idx / job.itemsPerSubframe;
            const uint32_t rem = (uint32_t)(idx - sub * job.itemsPerSubframe);
            const uint32_t tile = rem >> 5, within = rem & 31u;
            px = (tile % (uint32_t)job.tilesX) * 8u + (within & 7u);
            py = (tile / (uint32_t)job.tilesX) * 4u + (within >> 3);
            if (px >= (uint32_t)job.width || py >= (uint32_t)job.height) return false;
            pixel = py * (uint32_t)job.width + px;
        }
        const float dx = (float)px / (float)job.width * 2.f - 1.f;
        const float dy = (float)py / (float)job.height * 2.f - 1.f;
        const V3 U = mk(job.U[0], job.U[1], job.U[2]), V = mk(job.V[0], job.V[1], job.V[2]), W = mk(job.W[0], job.W[1], job.W[2]);
        o = mk(job.eye[0], job.eye[1], job.eye[2]);
        d = normalize<true>(dx * U + dy * V + W);
        val0 = px * 4096u + py;
        stream = job.firstSubframe + (uint32_t)sub;
        out = (uint32_t)sub * (uint32_t)(job.width * job.height) + pixel; /* the host keeps the staging buffer below 2^32 slots */
    } else if (job.kind == JOB_POINT) {
        const uint32_t t = (uint32_t)(idx / job.launches);
        const uint32_t l = (uint32_t)(idx - (unsigned long long)t * job.launches);
        const DsPointRadianceTask* task = job.tasks + t;
        o = mk(task->position[0], task->position[1], task->position[2]);
        d = mk(task->direction[0], task->direction[1], task->direction[2]);
        val0 = t * 4096u;
        stream = job.frame0 + l + 1u;
        out = (uint32_t)idx;
    } else {
        o = mk(job.origins[3 * idx], job.origins[3 * idx + 1], job.origins[3 * idx + 2]);
        d = mk(job.dirs[3 * idx], job.dirs[3 * idx + 1], job.dirs[3 * idx + 2]);
        val0 = job.seedVal0[idx];
        stream = job.stream[idx];
        out = (uint32_t)idx;
    }
    return true;
}

/* shared memory of k_trace_fast: padded CDF (float), the chopped phase sampler (half), the two guides: 30.1 KiB, under the 32 KiB carve-out */
constexpr size_t FAST_TABLE_BYTES = (size_t)CDF_PAD_N * 4 + (size_t)MIE_N * 2 + (size_t)(GUIDE_A_N + GUIDE_B_N) * 2;
static_assert(FAST_TABLE_BYTES + 1024 <= 32768, "the tables must fit the 32 KiB carve-out");
static_assert((CDF_PAD_N * 4) % 16 == 0, "alignment of the tables behind the CDF");
__device__ __forceinline__ void stageFastTables(const DevScene& sc, unsigned char* smemRaw, float*& sCdfPad, __half*& sChopped, uint16_t*& sGuideA,
                                                uint16_t*& sGuideB)
{
    sCdfPad = reinterpret_cast<float*>(smemRaw);
    sChopped = reinterpret_cast<__half*>(sCdfPad + CDF_PAD_N);
    sGuideA = reinterpret_cast<uint16_t*>(sChopped + MIE_N);
    sGuideB = sGuideA + GUIDE_A_N;
    for (int i = threadIdx.x; i < MIE_N; i += blockDim.x) {
        sChopped[i] = __ushort_as_half(sc.choppedHalf[i]);
        sCdfPad[i + 1] = sc.cdf[i];
    }
    if (threadIdx.x < CDF_PAD_N - MIE_N) sCdfPad[threadIdx.x == 0 ? 0 : MIE_N + threadIdx.x] = threadIdx.x == 0 ? 0.0f : 3.0e38f;
    for (int i = threadIdx.x; i < GUIDE_A_N; i += blockDim.x) sGuideA[i] = sc.guideA[i];
    for (int i = threadIdx.x; i < GUIDE_B_N; i += blockDim.x) sGuideB[i] = sc.guideB[i];
    __syncthreads();
}

/* lane states of k_trace_fast, one bit each so that a warp vote over a set of states is one LOP + VOTE */
enum FastLaneState : unsigned {
    /* one identity bit each (bits 24-29), plus a 1 in the byte that counts the lane's class: byte 0 = wants work (idle / done), byte 1 = waits
     * for the empty-space phase, byte 2 = busy (marching / event pending).  One warp-wide integer add over st & 0xffffff then gives the three
     * lane counts every phase decision of a round needs -- one REDUX instead of three votes and three population counts */
    F_IDLE = 0x01000001u,  /* wants a work item */
    F_SKIP = 0x02000100u,  /* the tap at the current position fell into an empty cell */
    F_MARCH = 0x04010000u, /* marching */
    F_EVENT = 0x08010000u, /* collided: next-event estimate + new direction pending */
    F_DONE = 0x10000001u,  /* path ended, sample not yet written */
    F_OFF = 0x20000000u    /* idle and the queue is exhausted */
};
/* lane counts of a warp: {free, skip, busy} in bytes 0, 1, 2 */
__device__ __forceinline__ unsigned laneCounts(unsigned st) { return __reduce_add_sync(0xffffffffu, st & 0x00ffffffu); }
#define DS_N_FREE(c) ((c) & 0xffu)
#define DS_N_SKIP(c) (((c) >> 8) & 0xffu)
#define DS_N_BUSY(c) ((c) >> 16)

/* start of the path of one radiance ray (closest-hit entry, cloudRadianceMaterials.cu:9-27 / 72-90) */
template <int MODE>
__device__ __forceinline__ unsigned beginPathFast(const DevScene& sc, const FastConsts& k, const TraceJob& job, const float* sCdfPad, const uint16_t* sGuideA,
                                             const uint16_t* sGuideB, V3 o, V3 d, uint32_t val0, uint32_t stream, FastState& s)
{
    s.rad = s.pendT = s.pendP = 0.0f;
    const float tHit = intersectBox(sc, o, d);
    if (tHit < 0.0f) return F_DONE;
    V3 hit = o + tHit * d;
    hit = hit + 0.5f * sc.bbox;
    s.q0 = hit * sc.texScale;
    V3 dir = normalize<true>(d);
    s.seed = tea4(val0, stream);
    s.depth = 0;
    const int mode = MODE >= 0 ? MODE : job.mode;
    if (mode == DS_MODE_SUN_MULTIPLE_SCATTER) dir = newDirectionFast(sCdfPad, sGuideA, sGuideB, s.seed, dir);
    s.sv = dir * k.stepTs;
    if (!beginFlight<true>(k, s)) return F_DONE;
    return F_MARCH;
}

template <int MODE>
__device__ __forceinline__ unsigned beginItemFast(const DevScene& sc, const FastConsts& k, const TraceJob& job, const float* sCdfPad,
                                             const uint16_t* sGuideA, const uint16_t* sGuideB, unsigned long long idx, FastState& s, bool& valid)
{
    V3 o, d;
    uint32_t val0, stream, pixel;
    valid = itemRay(job, idx, o, d, val0, stream, s.out, pixel);
    if (!valid) return F_IDLE;
    const unsigned st = beginPathFast<MODE>(sc, k, job, sCdfPad, sGuideA, sGuideB, o, d, val0, stream, s);
    /* cached empty-space leg of the primary ray: the first taps that can be non-zero follow step entrySteps[pixel] */
    if (st == F_MARCH && job.kind == JOB_RENDER && job.entrySteps) s.nf = (float)job.entrySteps[pixel];
    return st;
}

/* ---- JOB_ADAPTIVE: device-resident radiance collector ---- */
constexpr uint32_t AD_NONE = 0xffffffffu, AD_ALLDONE = 0xfffffffeu;

/* add c experiments (sum dx, sum of squares dxx) to sample sid and apply the convergence rule of
 * RadianceCollector.cpp:108-118 with the confidence intervals of PointRadianceTask.h:23-36 */
__device__ __forceinline__ void adaptiveCommit(const AdaptiveCollector& ad, uint32_t sid, double dx, double dxx, unsigned c)
{
    const unsigned long long c0 = atomicAdd(ad.count + sid, (unsigned long long)c);
    const double s0 = atomicAdd(ad.sum + sid, dx);
    const double q0 = atomicAdd(ad.sumSq + sid, dxx);
    const double N = (double)(c0 + c);
    /* test at the reference's cadence -- once per `minExperiments` new experiments of the sample (its update adds
     * repeat x 100 per sample between two tests) -- not after every commit: testing a running confidence interval more
     * often stops earlier on an under-estimated variance */
    if ((c0 + c) / ad.minExperiments == c0 / ad.minExperiments) return;
    const double mean = (s0 + dx) / N;
    const double m2 = fmax((q0 + dxx) - N * mean * mean, 0.0) * (double)ad.m2Scale;
    const float Nf = (float)N;
    const float sigma = sqrtf((float)m2 / Nf);
    const float absoluteCI = 1.96f * sigma / sqrtf(Nf);
    const float relativeCI = absoluteCI / ((float)mean + 1.1920929e-07f);
    bool converged = relativeCI < ad.relCI || absoluteCI < ad.absCI;
    if ((float)mean < 1.1920929e-07f) converged = N > (double)ad.zeroMin;
    if (converged) {
        if (atomicCAS(ad.flag + sid, 0u, 1u) == 0u) atomicAdd(ad.closed, 1u);
    } else if (ad.maxExperiments && N >= (double)ad.maxExperiments) {
        if (atomicCAS(ad.flag + sid, 0u, 2u) == 0u) atomicAdd(ad.closed, 1u); /* experiment cap (max_updates) */
    }
}

/* lane 0: next open sample and the first of `quota` fresh experiment ids for it */
__device__ __forceinline__ uint32_t adaptiveTicket(const AdaptiveCollector& ad, uint32_t& base)
{
    for (int tries = 0; tries < 32; ++tries) {
        if (*(volatile uint32_t*)ad.closed >= ad.nSamples) return AD_ALLDONE;
        const uint32_t sid = (uint32_t)(atomicAdd(ad.ticket, 1ull) % ad.nSamples);
        if (*(volatile uint32_t*)(ad.flag + sid) != 0u) continue;
        base = atomicAdd(ad.issued + sid, ad.quota);
        return sid;
    }
    return AD_NONE;
}

__device__ __forceinline__ void writeResultFast(const DevScene& sc, const TraceJob& job, const FastState& s)
{
    const float scale = sc.lightIntensity * SUN_TO_SPHERE * fmaf(s.pendT, s.pendP, s.rad);
    const float r = sc.lightColor.x * scale, g = sc.lightColor.y * scale, b = sc.lightColor.z * scale;
    if (!(fabsf(r + g + b) <= 3.0e38f)) atomicAdd(job.stats + CNT_NONFINITE, 1ull); /* never in a healthy run: not worth a register */
    if (job.kind == JOB_RENDER) {
        job.staging[s.out] = make_float4(r, g, b, 1.0f);
    } else if (job.kind == JOB_POINT) {
        job.xOut[s.out] = r;
    } else {
        job.radianceOut[3 * (size_t)s.out] = r;
        job.radianceOut[3 * (size_t)s.out + 1] = g;
        job.radianceOut[3 * (size_t)s.out + 2] = b;
    }
}

/*
 * Warp-level schedule.  Each lane owns one path; a warp cycles through four phases and runs a phase only when
 * enough of its lanes want it, so the (long, divergent) code of each phase executes with many lanes active:
 *   A  retire + regenerate   lanes whose path ended write their sample and pull the next work item from the
 *                            device queue (one warp-aggregated atomic); runs when >= regenMin lanes are free
 *   B  empty-space phase     lanes whose tap fell into an empty cell (paths leaving the cloud, paths started in
 *                            empty space): leap DDA, then continue marching; runs when >= skipMin lanes wait
 *   C  march phase           step + density tap + collision test.  With UNROLL = 2 the taps of the next TWO steps are
 *                            issued back to back and then consumed in order (the second is speculative: it is wasted
 *                            when the first step collides), which halves the exposed texture latency per step.
 *                            The loop body is the bare minimum: a zero tap changes nothing, so lanes in a hole or
 *                            past the cloud just keep stepping; whether they left the box (all face voxels are zero,
 *                            so taps outside the box read 0 as well) or may leap through empty space is looked at
 *                            ONCE per round, after the loop, and the number of steps the reference would have taken
 *                            is recovered exactly from the step index.  Grids with non-zero faces (BOXTEST) test the
 *                            box per step.
 *   D  event phase           lanes that collided: next-event estimate + new direction.  The sun-transmittance tap is
 *                            issued first and consumed last.
 * A waiting lane costs nothing but its slot; a phase entered with two lanes costs the whole warp its full
 * instruction stream, which is what the thresholds avoid.  Thresholds are ignored when nothing else can run.
 * MODE >= 0 fixes the estimator at compile time (DsMode); MODE = -1 reads job.mode.
 */
template <bool SKIP, bool BOXTEST, int UNROLL, int MODE, bool ADAPT, bool FUSED>
__global__ void __launch_bounds__(FAST_MAX_THREADS, 1)
    k_trace_fast(const __grid_constant__ DevScene sc, const __grid_constant__ TraceJob job, const __grid_constant__ FastConsts k)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float* sCdfPad;
    __half* sChopped;
    uint16_t *sGuideA, *sGuideB;
    stageFastTables(sc, smemRaw, sCdfPad, sChopped, sGuideA, sGuideB);

    const int mode = MODE >= 0 ? MODE : job.mode;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneLt = (1u << lane) - 1u;

    FastState s;
    s.q0 = s.sv = mk(0.f, 0.f, 0.f);
    s.nf = s.rad = s.tau = s.tauStar = s.pendT = s.pendP = 0.f;
    s.seed = 0;
    s.depth = 0;
    s.out = 0;
    unsigned st = F_IDLE;
    uint32_t nPaths = 0, nEvents = 0, nSteps = 0, nTaps = 0;
    float lastDensity = 0.0f;
    /* PIPE: every lane in F_MARCH holds the densities of its next two steps (nf + 1, nf + 2), fetched when the lane
     * ENTERED that state -- at the end of the event phase, of the empty-space phase or of regeneration -- so the
     * texture latency of the first march iteration of a round hides behind the rest of the previous round */
    constexpr bool PIPE = !BOXTEST && UNROLL == 2;
    float pd1 = 0.0f, pd2 = 0.0f;
    /* speculation threshold (DS_ISSUE_TAPS): specPercent = 0 always fetches both taps */
    const float specC1 = k.c1 * (float)job.specPercent * 0.01f;
    unsigned roundIdx = 0;
    uint32_t wSample = 0, wNext = 0, wQuota = 0; /* ADAPT: the warp's current ticket (uniform across lanes) */

    for (;;) {
        unsigned cnt = laneCounts(st);

        /* ---- A: retire + regenerate ---- */
        if (ADAPT && DS_N_FREE(cnt) && ((int)DS_N_FREE(cnt) >= job.regenMin || (DS_N_BUSY(cnt) == 0u && (int)DS_N_SKIP(cnt) < max(job.skipMin, 1)))) {
            /* retire: finished paths are summed per sample (lanes of a warp mostly share one) and committed by one lane */
            const bool isDone = st == F_DONE;
            unsigned rem = __ballot_sync(FULL, isDone);
            float x = 0.0f;
            if (isDone) {
                x = sc.lightColor.x * (sc.lightIntensity * SUN_TO_SPHERE * fmaf(s.pendT, s.pendP, s.rad)); /* pointEmissionCamera.cu:32: result.x */
                if (!(fabsf(x) <= 3.0e38f)) {
                    atomicAdd(job.stats + CNT_NONFINITE, 1ull);
                    x = 0.0f;
                }
                st = F_IDLE;
            }
#pragma unroll 1
            while (rem) {
                const int l0 = __ffs(rem) - 1;
                const uint32_t sid0 = __shfl_sync(FULL, s.out, l0);
                const bool mine = isDone && s.out == sid0;
                const unsigned same = __ballot_sync(FULL, mine);
                double dx = mine ? (double)x : 0.0, dxx = dx * dx;
                for (int o = 16; o > 0; o >>= 1) {
                    dx += __shfl_xor_sync(FULL, dx, o);
                    dxx += __shfl_xor_sync(FULL, dxx, o);
                }
                if ((int)lane == l0) adaptiveCommit(job.ad, sid0, dx, dxx, (unsigned)__popc(same));
                rem &= ~same;
            }
            /* regenerate: free lanes take the next experiments of the warp's ticket */
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                const unsigned need = __ballot_sync(FULL, st == F_IDLE);
                if (need == 0u) break;
                if (wQuota == 0u) {
                    uint32_t sidNew = AD_NONE, base = 0;
                    if (lane == 0) sidNew = adaptiveTicket(job.ad, base);
                    sidNew = __shfl_sync(FULL, sidNew, 0);
                    base = __shfl_sync(FULL, base, 0);
                    if (sidNew == AD_ALLDONE) {
                        if (st == F_IDLE) st = F_OFF;
                        break;
                    }
                    if (sidNew == AD_NONE) break; /* every ticket drawn was a closed sample: try again next round */
                    wSample = sidNew;
                    wNext = base;
                    wQuota = job.ad.quota;
                }
                const unsigned take = min((unsigned)__popc(need), wQuota);
                const unsigned rank = (unsigned)__popc(need & laneLt);
                if (st == F_IDLE && rank < take) {
                    const DsPointRadianceTask* task = job.tasks + wSample;
                    const V3 o = mk(task->position[0], task->position[1], task->position[2]);
                    const V3 d = mk(task->direction[0], task->direction[1], task->direction[2]);
                    s.out = wSample;
                    st = beginPathFast<MODE>(sc, k, job, sCdfPad, sGuideA, sGuideB, o, d, wSample * 4096u, wNext + rank + 1u, s);
                    nPaths++;
                    if (PIPE && st == F_MARCH) DS_ISSUE_TAPS(s.nf, 0.0f);
                }
                wNext += take;
                wQuota -= take;
            }
            cnt = laneCounts(st);
        }
        if (!ADAPT && DS_N_FREE(cnt) && ((int)DS_N_FREE(cnt) >= job.regenMin || (DS_N_BUSY(cnt) == 0u && (int)DS_N_SKIP(cnt) < max(job.skipMin, 1)))) {
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                if (st == F_DONE) {
                    writeResultFast(sc, job, s);
                    st = F_IDLE;
                }
                const unsigned need = __ballot_sync(FULL, st == F_IDLE);
                if (need == 0u) break;
                unsigned long long base = 0;
                const int leader = __ffs(need) - 1;
                if ((int)lane == leader) base = atomicAdd(job.queue, (unsigned long long)__popc(need));
                base = __shfl_sync(FULL, base, leader);
                if (st == F_IDLE) {
                    const unsigned long long idx = base + __popc(need & laneLt);
                    if (idx >= job.total) {
                        st = F_OFF;
                    } else {
                        bool valid;
                        st = beginItemFast<MODE>(sc, k, job, sCdfPad, sGuideA, sGuideB, idx, s, valid);
                        if (valid) nPaths++;
                        if (PIPE && st == F_MARCH) DS_ISSUE_TAPS(s.nf, 0.0f);
                    }
                }
                cnt = laneCounts(st);
                if ((int)DS_N_FREE(cnt) < job.regenMin) break;
            }
        }
        if (cnt == 0u) break; /* every lane off: the queue is exhausted */

        /* ---- B: empty-space phase: the tap at the current position fell into an empty cell ---- */
        if (SKIP && DS_N_SKIP(cnt) && ((int)DS_N_SKIP(cnt) >= job.skipMin || DS_N_BUSY(cnt) == 0u)) {
            if (st == F_SKIP) {
                bool more, leaves;
                const V3 qs = posAt(s, s.nf);
                float kf = emptySteps(sc, k, qs, s.sv, job.skipMaxIters, more, leaves);
                if (leaves) kf = stepsToLeaveBox(k, qs, s.sv);
                s.nf += kf;
                if (!BOXTEST && leaves && !inBoxTs(k, posAt(s, s.nf))) {
                    /* nothing but zeros up to the box exit, and the landing position is outside: the path ends here.  The reference
                     * stops at the first position outside the box (a landing position that rounding left inside marches on below) */
                    float n = s.nf;
#pragma unroll 1
                    while (n >= 2.0f && !inBoxTs(k, posAt(s, n - 1.0f))) n -= 1.0f;
                    nSteps += (uint32_t)n;
                    st = F_DONE;
                } else {
                    /* a walk cut short lands in an empty cell and continues next round; otherwise march on */
                    st = (more && kf >= 1.0f) ? F_SKIP : F_MARCH;
                    if (PIPE && st == F_MARCH) DS_ISSUE_TAPS(s.nf, 0.0f);
                }
            }
            cnt = laneCounts(st);
        }

        /* ---- C: march phase (CU/cloud.cuh:87-104) ---- */
        if (DS_N_BUSY(cnt)) {
            const int keep = ((int)DS_N_BUSY(cnt) * job.marchKeep32) >> 5; /* leave when at most this many lanes still march */
            unsigned mMarch = 0;
#pragma unroll 1
            for (int it = 0; it < job.marchMaxIters; it += UNROLL) {
                if (PIPE) {
                    if (st == F_MARCH) {
                        const float n1 = s.nf + 1.0f, n2 = s.nf + 2.0f;
                        const float t1 = fmaf(pd1, k.c1, s.tau);
                        const bool hit1 = t1 > s.tauStar;
                        const bool one = hit1 || pd2 < 0.0f; /* only the first step is consumed: it collided, or the second tap was not fetched */
                        const float t2 = one ? t1 : fmaf(pd2, k.c1, t1); /* pd2 >= 0: t2 >= t1 */
                        s.nf = one ? n1 : n2;
                        s.tau = t2;
                        lastDensity = one ? pd1 : pd2;
                        if (t2 > s.tauStar) {
                            st = F_EVENT;
                        } else {
                            DS_ISSUE_TAPS(s.nf, lastDensity);
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                        if (st == F_MARCH) {
                            if (BOXTEST && !inBoxTs(k, posAt(s, s.nf))) {
                                nSteps += (uint32_t)s.nf;
                                st = F_DONE;
                            } else {
                                s.nf += 1.0f;
                                lastDensity = tapDensity<FUSED>(sc, posAt(s, s.nf));
                                s.tau = fmaf(lastDensity, k.c1, s.tau);
                                if (s.tau > s.tauStar) st = F_EVENT;
                                nTaps++;
                            }
                        }
                    }
                }
                mMarch = __ballot_sync(FULL, st == F_MARCH);
                if (__popc(mMarch) <= keep) break;
            }

            /* once per round: lanes whose last tap read 0 */
            /* (with zeroCheckMin > 1 the look is postponed until that many lanes wait, but never beyond 4 rounds and
             * never when every marching lane waits: a lane past the cloud only wastes zero taps meanwhile) */
            const bool zero = st == F_MARCH && lastDensity == 0.0f;
            const unsigned mZero = __ballot_sync(FULL, zero);
            /* mMarch is the vote that ended the loop: the lanes still marching */
            if (mZero && (__popc(mZero) >= job.zeroCheckMin || mZero == mMarch || (++roundIdx & 3u) == 0u)) {
                if (zero) {
                    const V3 q = posAt(s, s.nf);
                    if (!BOXTEST && !inBoxTs(k, q)) {
                        /* left the box: the reference stops at the first position outside it (its taps beyond read 0 too) */
                        float n = s.nf;
#pragma unroll 1
                        while (n >= 2.0f && !inBoxTs(k, posAt(s, n - 1.0f))) n -= 1.0f;
                        nSteps += (uint32_t)n;
                        st = F_DONE;
                    } else if (SKIP && tapCellDistance(sc, k, q) >= job.skipOpenDist) {
                        /* open space (no occupied cell within skipOpenDist - 1 cells): leap; pockets next to the cloud
                         * are cheaper to march through */
                        st = F_SKIP;
                    }
                }
            }
        }

        /* ---- D: event phase (cloudRadianceMaterials.cu:49-61) ---- */
        if (st == F_EVENT) {
            s.rad = fmaf(s.pendT, s.pendP, s.rad); /* the previous event's estimate */
            s.pendT = s.pendP = 0.0f;
            nSteps += (uint32_t)s.nf;
            /* cloud.cuh:99: scatterPos = pos - dir * log(xi / T) / sigma, with log(xi / T) = tau - tauStar; in march
             * steps: (tau - tauStar) / (density * mult * step) */
            const float back = __fdividef((s.tau - s.tauStar) * k.backScale, lastDensity);
            s.q0 = posAt(s, s.nf - back);
            /* The reference's loop tests isInBox(scatterPos) before it goes on (cloudRadianceMaterials.cu:30).  Without BOXTEST the test
             * cannot fail and is not compiled: the grid's faces are zero (borderEmpty), so the step that collided (density > 0) lies strictly
             * inside the grid, the collision point lies less than one step (<= 0.01 of the box, checked by the launcher) before it, and the
             * box slab reaches 0.01 beyond the grid (cloud.cuh:40-44) */
            if (BOXTEST && !inBoxTs(k, s.q0)) {
                st = F_DONE;
            } else {
                s.pendT = tapSun<FUSED>(sc, s.q0); /* consumed at the next event */
                const float cosLightAngle = -dot(k.lightTs, s.sv);
                const float u = (cosLightAngle + 1.0f) * 0.5f;
                if (mode == DS_MODE_SUN_MULTIPLE_SCATTER || (mode == DS_MODE_SUN_AND_SKY_ALL_SCATTER && s.depth != 1))
                    s.pendP = tableLerp(sChopped, u);
                else
                    s.pendP = tableLerp(sc.mie, u);
                nEvents++;
                if (mode == DS_MODE_SUN_SINGLE_SCATTER) {
                    st = F_DONE;
                } else {
                    const V3 dir = newDirectionFast(sCdfPad, sGuideA, sGuideB, s.seed, s.sv * k.invStepTs);
                    s.sv = dir * k.stepTs;
                    st = beginFlight<false>(k, s) ? F_MARCH : F_DONE; /* q0 is the scatter position just verified in-box */
                    if (PIPE && st == F_MARCH) DS_ISSUE_TAPS(0.0f, lastDensity);
                }
            }
        }
    }

    unsigned long long c[4] = {nPaths, nEvents, nSteps, nTaps};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        unsigned long long v = c[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
        if (lane == 0 && v) atomicAdd(job.stats + i, v);
    }
}

template <bool SKIP, bool BOXTEST, int UNROLL, int MODE, bool ADAPT = false, bool FUSED = false>
static cudaError_t launchFast(const DevScene& sc, const TraceJob& job, int blocks, int threads, size_t smem, cudaStream_t st, int carveout = -1)
{
    cudaError_t e = cudaFuncSetAttribute(k_trace_fast<SKIP, BOXTEST, UNROLL, MODE, ADAPT, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (carveout >= 0) {
        e = cudaFuncSetAttribute(k_trace_fast<SKIP, BOXTEST, UNROLL, MODE, ADAPT, FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        if (e != cudaSuccess) return e;
    }
    k_trace_fast<SKIP, BOXTEST, UNROLL, MODE, ADAPT, FUSED><<<blocks, threads, smem, st>>>(sc, job, makeConsts(sc));
    return cudaGetLastError();
}

template <>
cudaError_t KernelSet<true>::trace(const DevScene& sc, const TraceJob& job, const LaunchConfig& cfg, cudaStream_t st)
{
    if (cfg.variant == 1) return traceGeneric<true>(sc, job, cfg, st);
    const size_t smem = FAST_TABLE_BYTES;
    const int threads = cfg.blockThreads > FAST_MAX_THREADS ? FAST_MAX_THREADS : cfg.blockThreads;
    const unsigned long long wantBlocks = (job.total + threads - 1) / threads;
    const unsigned long long maxBlocks = (unsigned long long)cfg.smCount * cfg.blocksPerSm;
    const int blocks = (int)(wantBlocks < maxBlocks ? (wantBlocks ? wantBlocks : 1) : maxBlocks);
    /* the variants without a per-step box test rely on zero faces and on a sampling step within the slack of the box test (0.01) */
    const bool boxtest = sc.borderEmpty == 0 || !(sc.step <= 0.01f);
    if (job.kind == JOB_ADAPTIVE) {
        /* the collector always runs multipleScatterSunRadiance (Tasks.cpp:134); the host falls back to the update loop for
         * grids with non-zero faces */
        if (boxtest || !cfg.skipEmpty || job.mode != DS_MODE_SUN_MULTIPLE_SCATTER) return cudaErrorInvalidValue;
        if (sc.fusedTex) return launchFast<true, false, 2, DS_MODE_SUN_MULTIPLE_SCATTER, true, true>(sc, job, (int)maxBlocks, threads, smem, st);
        return launchFast<true, false, 2, DS_MODE_SUN_MULTIPLE_SCATTER, true, false>(sc, job, (int)maxBlocks, threads, smem, st);
    }
    if (!cfg.skipEmpty || boxtest) {
        /* uncommon configurations (grids with non-zero faces, empty-space skipping switched off): estimator read at run time */
        if (cfg.skipEmpty) return launchFast<true, true, 1, -1>(sc, job, blocks, threads, smem, st);
        if (boxtest) return launchFast<false, true, 1, -1>(sc, job, blocks, threads, smem, st);
        return launchFast<false, false, 1, -1>(sc, job, blocks, threads, smem, st);
    }
    const bool u2 = cfg.marchUnroll >= 2;
    const bool fused = sc.fusedTex != 0;
#define DS_LAUNCH_MODE(M)                                                                                                           \
    (fused ? (u2 ? launchFast<true, false, 2, M, false, true>(sc, job, blocks, threads, smem, st, cfg.smemCarveout)                 \
                 : launchFast<true, false, 1, M, false, true>(sc, job, blocks, threads, smem, st, cfg.smemCarveout))                \
           : (u2 ? launchFast<true, false, 2, M, false, false>(sc, job, blocks, threads, smem, st, cfg.smemCarveout)                \
                 : launchFast<true, false, 1, M, false, false>(sc, job, blocks, threads, smem, st, cfg.smemCarveout)))
    switch (job.mode) {
    case DS_MODE_SUN_AND_SKY_ALL_SCATTER: return DS_LAUNCH_MODE(DS_MODE_SUN_AND_SKY_ALL_SCATTER);
    case DS_MODE_SUN_MULTIPLE_SCATTER: return DS_LAUNCH_MODE(DS_MODE_SUN_MULTIPLE_SCATTER);
    default: return DS_LAUNCH_MODE(DS_MODE_SUN_SINGLE_SCATTER);
    }
#undef DS_LAUNCH_MODE
}

/*
 * Primary-ray pre-pass: one thread per pixel.  Walks the camera ray from the box entry through empty cells and
 * records how many march steps precede the first tap that can be non-zero; pixels whose ray never reaches an
 * occupied cell are marked ENTRY_MISS (their radiance sample is exactly 0 for every subframe) and the number of
 * march steps the reference algorithm would spend on them is summed for the work counters.
 * Pixels are enumerated in SUPER-TILE order (64 x 64 pixels = 8 x 16 tiles of 8 x 4, row-major inside, super-tiles row-major), and
 * the hitting pixels are compacted in exactly that order (per-block counts -> exclusive scan -> ordered scatter), so a run of
 * consecutive hit-list entries is a compact patch of the image and the list is the same on every run.
 */
constexpr uint32_t SUPER_TILES_X = 8, SUPER_TILES_Y = 16, SUPER_TILES = SUPER_TILES_X * SUPER_TILES_Y;

__device__ __forceinline__ bool prepassPixel(const TraceJob& job, uint32_t rem, uint32_t& px, uint32_t& py)
{
    const uint32_t tile = rem >> 5, within = rem & 31u;
    const uint32_t superCols = ((uint32_t)job.tilesX + SUPER_TILES_X - 1) / SUPER_TILES_X;
    const uint32_t sup = tile / SUPER_TILES, w = tile % SUPER_TILES;
    const uint32_t tx = (sup % superCols) * SUPER_TILES_X + (w % SUPER_TILES_X);
    const uint32_t ty = (sup / superCols) * SUPER_TILES_Y + (w / SUPER_TILES_X);
    px = tx * 8u + (within & 7u);
    py = ty * 4u + (within >> 3);
    return tx < (uint32_t)job.tilesX && px < (uint32_t)job.width && py < (uint32_t)job.height;
}

static unsigned long long prepassItems(const TraceJob& job)
{
    const unsigned long long tilesY = ((unsigned long long)job.height + 3) / 4;
    const unsigned long long superCols = ((unsigned long long)job.tilesX + SUPER_TILES_X - 1) / SUPER_TILES_X;
    const unsigned long long superRows = (tilesY + SUPER_TILES_Y - 1) / SUPER_TILES_Y;
    return superCols * superRows * SUPER_TILES * 32ull;
}

__global__ void __launch_bounds__(256) k_primary_prepass(const DevScene sc, const TraceJob job, const FastConsts k, uint32_t* __restrict__ entrySteps,
                                                         uint32_t* __restrict__ blockCounts, unsigned long long* __restrict__ counts)
{
    const uint32_t rem = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t px, py;
    const bool valid = prepassPixel(job, rem, px, py);
    bool hit = false;
    uint32_t steps = 0;
    const uint32_t pixel = py * (uint32_t)job.width + px;
    if (valid) {
        const float dx = (float)px / (float)job.width * 2.f - 1.f;
        const float dy = (float)py / (float)job.height * 2.f - 1.f;
        const V3 U = mk(job.U[0], job.U[1], job.U[2]), V = mk(job.V[0], job.V[1], job.V[2]), W = mk(job.W[0], job.W[1], job.W[2]);
        const V3 o = mk(job.eye[0], job.eye[1], job.eye[2]);
        const V3 d = normalize<true>(dx * U + dy * V + W);
        const float tHit = intersectBox(sc, o, d);
        if (tHit >= 0.0f) {
            V3 hitp = o + tHit * d;
            hitp = hitp + 0.5f * sc.bbox;
            FastState f;
            f.q0 = hitp * sc.texScale;
            f.sv = normalize<true>(d) * k.stepTs;
            /* march like the reference (cloud.cuh:87-89) but only look at occupancy; positions by step index, exactly
             * as k_trace_fast computes them */
            float n = 0.0f;
            while (inBoxTs(k, posAt(f, n))) {
                if (!tapCellEmpty(sc, k, sc.occ, posAt(f, n + 1.0f))) {
                    hit = true; /* the next step's tap may be non-zero: the traced path starts at step n */
                    break;
                }
                n += 1.0f;
                bool more, leaves;
                const V3 qs = posAt(f, n);
                const float kf = emptySteps(sc, k, qs, f.sv, 256, more, leaves);
                n += leaves ? stepsToLeaveBox(k, qs, f.sv) : kf;
            }
            /* a ray that leaves the box: the reference stops at the first position outside it */
            while (!hit && n >= 2.0f && !inBoxTs(k, posAt(f, n - 1.0f))) n -= 1.0f;
            steps = (uint32_t)n;
        }
        entrySteps[pixel] = hit ? steps : ENTRY_MISS;
    }
    const int blockHits = __syncthreads_count(hit);
    if (threadIdx.x == 0) blockCounts[blockIdx.x] = (uint32_t)blockHits;
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long missSteps = (valid && !hit) ? steps : 0ull;
    for (int o = 16; o > 0; o >>= 1) missSteps += __shfl_down_sync(0xffffffffu, missSteps, o);
    if (lane == 0 && missSteps) atomicAdd(counts + 1, missSteps);
}

/* exclusive scan of the per-block hit counts, in place; counts[0] = total.  One block. */
__global__ void __launch_bounds__(1024) k_primary_scan(uint32_t* __restrict__ blockCounts, uint32_t nBlocks, unsigned long long* __restrict__ counts)
{
    __shared__ uint32_t partial[1024];
    const uint32_t per = (nBlocks + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, nBlocks);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += blockCounts[i];
    partial[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) { /* Hillis-Steele inclusive scan */
        const uint32_t v = threadIdx.x >= off ? partial[threadIdx.x - off] : 0u;
        __syncthreads();
        partial[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = partial[threadIdx.x] - sum;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = blockCounts[i];
        blockCounts[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) counts[0] = partial[1023];
}

/* ordered scatter of the hitting pixels: same enumeration as k_primary_prepass */
__global__ void __launch_bounds__(256) k_primary_compact(const TraceJob job, const uint32_t* __restrict__ entrySteps, const uint32_t* __restrict__ blockBase,
                                                         uint32_t* __restrict__ hitList)
{
    __shared__ uint32_t warpBase[8];
    const uint32_t rem = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t px, py;
    const bool valid = prepassPixel(job, rem, px, py);
    const uint32_t pixel = py * (uint32_t)job.width + px;
    const bool hit = valid && entrySteps[pixel] != ENTRY_MISS;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warpBase[warp] = (uint32_t)__popc(m);
    __syncthreads();
    uint32_t base = blockBase[blockIdx.x];
    for (unsigned w = 0; w < warp; w++) base += warpBase[w];
    if (hit) hitList[base + __popc(m & ((1u << lane) - 1u))] = pixel;
}

template <>
cudaError_t KernelSet<true>::primaryPrepass(const DevScene& sc, const TraceJob& cam, uint32_t* entrySteps, uint32_t* hitList,
                                            unsigned long long* counts, cudaStream_t st)
{
    const unsigned long long items = prepassItems(cam);
    const unsigned nBlocks = (unsigned)((items + 255) / 256);
    uint32_t* blockCounts = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&blockCounts, (size_t)nBlocks * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    k_primary_prepass<<<nBlocks, 256, 0, st>>>(sc, cam, makeConsts(sc), entrySteps, blockCounts, counts);
    k_primary_scan<<<1, 1024, 0, st>>>(blockCounts, nBlocks, counts);
    k_primary_compact<<<nBlocks, 256, 0, st>>>(cam, entrySteps, blockCounts, hitList);
    e = cudaGetLastError();
    cudaFreeAsync(blockCounts, st);
    return e;
}

template struct KernelSet<true>;

/* introspection (ds_invert_phase_cdf): the inversion of the chopped-Mie CDF exactly as k_trace_fast runs it, tables staged the same way */
__global__ void __launch_bounds__(256) k_invert_cdf(const __grid_constant__ DevScene sc, const float* __restrict__ val, uint32_t n, float* __restrict__ cosTheta,
                                                    float* __restrict__ phase)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float* sCdfPad;
    __half* sChopped;
    uint16_t *sGuideA, *sGuideB;
    stageFastTables(sc, smemRaw, sCdfPad, sChopped, sGuideA, sGuideB);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        cosTheta[i] = invertCdf(sCdfPad, sGuideA, sGuideB, val[i]);
        phase[i] = tableLerp(sChopped, val[i]); /* the chopped phase sampler at u = val */
    }
}

cudaError_t launchInvertCdf(const DevScene& sc, const float* val, uint32_t n, float* cosTheta, float* phase, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    k_invert_cdf<<<148, 256, FAST_TABLE_BYTES, st>>>(sc, val, n, cosTheta, phase);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_interleave(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int nx, int ny, int nz,
                                                    cudaSurfaceObject_t surf)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, z = blockIdx.z;
    if (x >= nx) return;
    const size_t i = ((size_t)z * ny + y) * nx + x;
    surf3Dwrite(make_uchar2(a[i], b[i]), surf, x * (int)sizeof(uchar2), y, z);
}

cudaError_t launchInterleave(const uint8_t* a, const uint8_t* b, int nx, int ny, int nz, cudaSurfaceObject_t surf, cudaStream_t st)
{
    if (ny > 65535 || nz > 65535) return cudaErrorInvalidValue;
    k_interleave<<<dim3((nx + 255) / 256, ny, nz), 256, 0, st>>>(a, b, nx, ny, nz, surf);
    return cudaGetLastError();
}

} // namespace dsk
