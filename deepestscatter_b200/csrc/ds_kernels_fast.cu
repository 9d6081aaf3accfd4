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
 *   - empty space: a lane whose tap returned exactly 0 (and every new path) looks up the occupancy bit of its
 *     tap cell (shared memory) and, if the cell is empty, walks the ray through the occupancy cells with a 3-D
 *     DDA until the next occupied cell; it then advances k whole march steps at once, k chosen so that every
 *     skipped tap lies in cells that are known to be empty.  Skipped steps read density 0, i.e. tau and the collision test are unchanged;
 *     the step counter still advances by k (the reference algorithm performs those steps);
 *   - the 16-step bisection of the chopped-Mie CDF (cloud.cuh:167-178) is replaced by the closed-form
 *     inverse of the same piecewise-linear CDF: guide table -> short binary search for the table cell ->
 *     linear solve.  The bisection converges to that root within 2^-16;
 *   - radiance is accumulated as the scalar sum of Tsun*phase and scaled by lightColor*lightIntensity*ratio
 *     once per path.
 */
#include "ds_kernels.cuh"

namespace dsk {

struct FastState {
    V3 q;   /* position in texture coordinates */
    V3 dir; /* unit direction (world / box-local axes) */
    float rad;
    float tau;     /* optical depth accumulated in the current free flight */
    float tauStar; /* -ln(xi) */
    uint32_t seed;
    int depth;
    unsigned long long out;
};

struct FastConsts {
    V3 stepTs;  /* sampleStep * textureScale */
    V3 half;    /* 0.5 + 0.01 * textureScale: half extent of the in-box slab in texture space */
    float c1;   /* densityMultiplier * sampleStep */
    float nxf, nyf, nzf;
};

__device__ __forceinline__ bool inBoxTs(const FastConsts& k, V3 q)
{
    return fabsf(q.x - 0.5f) <= k.half.x && fabsf(q.y - 0.5f) <= k.half.y && fabsf(q.z - 0.5f) <= k.half.z;
}

/* closed-form inverse of the piecewise-linear CDF that cloud.cuh:167-178 bisects */
__device__ __forceinline__ float invertCdf(const float* sCdf, const uint16_t* sGuide, float val)
{
    const int k = min((int)(val * (float)GUIDE_N), GUIDE_N - 1);
    int lo = sGuide[k], hi = sGuide[k + 1];
    while (lo < hi) { /* first index with cdf[i] >= val */
        const int mid = (lo + hi) >> 1;
        if (sCdf[mid] < val)
            lo = mid + 1;
        else
            hi = mid;
    }
    float u;
    if (lo == 0) {
        u = 0.0f;
    } else if (lo >= MIE_N) {
        u = 1.0f;
    } else {
        const float a = sCdf[lo - 1], b = sCdf[lo];
        const float t = __fdividef(val - a, b - a);
        u = ((float)lo - 0.5f + t) * (1.0f / (float)MIE_N);
    }
    return 2.0f * u - 1.0f;
}

/* CU/cloud.cuh:160-188 */
__device__ __forceinline__ V3 newDirectionFast(const float* sCdf, const uint16_t* sGuide, uint32_t& seed, V3 prev)
{
    const float val = rnd(seed);
    const float cosTheta = invertCdf(sCdf, sGuide, val);
    const float phi = rnd(seed) * (PI_F * 2.0f);
    const float s2 = fmaxf(fmaf(-cosTheta, cosTheta, 1.0f), 0.0f);
    const float sinTheta = s2 * rsqrtf(fmaxf(s2, 1.0e-30f));
    float s, c;
    __sincosf(phi, &s, &c);
    V3 d = onbInverseTransform<true>(prev, mk(sinTheta * c, sinTheta * s, cosTheta));
    return d * rsqrtf(dot(d, d));
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

template <bool CHECK_BOX>
__device__ __forceinline__ int loopTopFast(const FastConsts& k, FastState& s)
{
    if (CHECK_BOX && !inBoxTs(k, s.q)) return ST_DONE;
    s.depth++;
    if (s.depth == MAX_DEPTH) return ST_DONE;
    const float xi = rnd(s.seed);
    s.tauStar = -__logf(xi); /* xi == 0 -> +inf: never collides, as `0 > T` in the reference */
    s.tau = 0.0f;
    return ST_MARCH;
}

/* The tap at q is about to be taken.  If its cell is empty the tap is skipped (it would read 0) and the lane
 * walks the ray through the occupancy cells with a 3-D DDA until the next cell is occupied or the grid ends; it
 * then advances the largest whole number of march steps whose taps all lie in the empty cells just walked.
 * Coordinates: x = u*N - 0.5 is the voxel coordinate whose floor is the low corner of the trilinear footprint;
 * the occupancy bit of cell c covers voxels [c*2^s, c*2^s + 2^s], i.e. every footprint with floor(x) in c.
 * Returns true when the tap is skipped. */
__device__ __forceinline__ bool skipEmpty(const DevScene& sc, const FastConsts& k, const uint32_t* sOcc, FastState& s, uint32_t& nSteps)
{
    const float x = fmaf(s.q.x, k.nxf, -0.5f), y = fmaf(s.q.y, k.nyf, -0.5f), z = fmaf(s.q.z, k.nzf, -0.5f);
    const int fx = __float2int_rd(x), fy = __float2int_rd(y), fz = __float2int_rd(z);
    int cx = min(max(fx, 0), sc.nx - 1) >> sc.occShift;
    int cy = min(max(fy, 0), sc.ny - 1) >> sc.occShift;
    int cz = min(max(fz, 0), sc.nz - 1) >> sc.occShift;
    int cell = (cz * sc.ocy + cy) * sc.ocx + cx;
    if ((sOcc[cell >> 5] >> (cell & 31)) & 1u) return false;
    /* outside the grid the footprint is clamped to edge voxels: no walk, just skip this tap */
    if ((unsigned)fx >= (unsigned)(sc.nx - 1) || (unsigned)fy >= (unsigned)(sc.ny - 1) || (unsigned)fz >= (unsigned)(sc.nz - 1)) return true;
    const float cs = (float)(1 << sc.occShift);
    const float vx = s.dir.x * k.stepTs.x * k.nxf, vy = s.dir.y * k.stepTs.y * k.nyf, vz = s.dir.z * k.stepTs.z * k.nzf; /* voxels per step */
    const float big = 3.0e38f;
    const float ix = fabsf(vx) > 1e-12f ? __fdividef(1.0f, vx) : big, iy = fabsf(vy) > 1e-12f ? __fdividef(1.0f, vy) : big,
                iz = fabsf(vz) > 1e-12f ? __fdividef(1.0f, vz) : big;
    const int sx = vx > 0.0f ? 1 : -1, sy = vy > 0.0f ? 1 : -1, sz = vz > 0.0f ? 1 : -1;
    /* steps until the ray crosses into the neighbouring cell along each axis, and per-cell increments */
    float tx = fabsf(ix) >= big ? big : ((float)(cx + (sx > 0 ? 1 : 0)) * cs - x) * ix;
    float ty = fabsf(iy) >= big ? big : ((float)(cy + (sy > 0 ? 1 : 0)) * cs - y) * iy;
    float tz = fabsf(iz) >= big ? big : ((float)(cz + (sz > 0 ? 1 : 0)) * cs - z) * iz;
    const float dtx = fabsf(ix) >= big ? 0.0f : cs * fabsf(ix), dty = fabsf(iy) >= big ? 0.0f : cs * fabsf(iy),
                dtz = fabsf(iz) >= big ? 0.0f : cs * fabsf(iz);
    const int strideY = sc.ocx, strideZ = sc.ocx * sc.ocy;
    float t = 0.0f, margin = 0.0f;
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
        const bool ax = tx <= ty && tx <= tz;
        const bool ay = !ax && ty <= tz;
        if (ax) {
            t = tx;
            margin = fabsf(ix);
            tx += dtx;
            cx += sx;
            cell += sx;
            if ((unsigned)cx >= (unsigned)sc.ocx) break;
        } else if (ay) {
            t = ty;
            margin = fabsf(iy);
            ty += dty;
            cy += sy;
            cell += sy * strideY;
            if ((unsigned)cy >= (unsigned)sc.ocy) break;
        } else {
            t = tz;
            margin = fabsf(iz);
            tz += dtz;
            cz += sz;
            cell += sz * strideZ;
            if ((unsigned)cz >= (unsigned)sc.ocz) break;
        }
        if ((sOcc[cell >> 5] >> (cell & 31)) & 1u) break;
    }
    /* stay 0.01 voxel short of the cell boundary that stopped the walk */
    const float kf = floorf(fminf(t - 0.01f * margin, 8192.0f));
    if (kf >= 1.0f) {
        s.q.x = fmaf(kf * s.dir.x, k.stepTs.x, s.q.x);
        s.q.y = fmaf(kf * s.dir.y, k.stepTs.y, s.q.y);
        s.q.z = fmaf(kf * s.dir.z, k.stepTs.z, s.q.z);
        nSteps += (uint32_t)kf;
    }
    return true;
}

template <bool SKIP>
__device__ __forceinline__ int beginItemFast(const DevScene& sc, const FastConsts& k, const TraceJob& job, const float* sCdf,
                                             const uint16_t* sGuide, unsigned long long idx, FastState& s, bool& valid)
{
    V3 o, d;
    uint32_t val0, stream;
    valid = true;
    if (job.kind == JOB_RENDER) {
        const unsigned long long sub = idx / job.itemsPerSubframe;
        const uint32_t rem = (uint32_t)(idx - sub * job.itemsPerSubframe);
        const uint32_t tile = rem >> 5, within = rem & 31u;
        const uint32_t px = (tile % (uint32_t)job.tilesX) * 8u + (within & 7u);
        const uint32_t py = (tile / (uint32_t)job.tilesX) * 4u + (within >> 3);
        if (px >= (uint32_t)job.width || py >= (uint32_t)job.height) {
            valid = false;
            return ST_IDLE;
        }
        const float dx = (float)px / (float)job.width * 2.f - 1.f;
        const float dy = (float)py / (float)job.height * 2.f - 1.f;
        const V3 U = mk(job.U[0], job.U[1], job.U[2]), V = mk(job.V[0], job.V[1], job.V[2]), W = mk(job.W[0], job.W[1], job.W[2]);
        o = mk(job.eye[0], job.eye[1], job.eye[2]);
        d = normalize<true>(dx * U + dy * V + W);
        val0 = px * 4096u + py;
        stream = job.firstSubframe + (uint32_t)sub;
        s.out = sub * (unsigned long long)job.width * job.height + (unsigned long long)py * job.width + px;
    } else if (job.kind == JOB_POINT) {
        const uint32_t t = (uint32_t)(idx / job.launches);
        const uint32_t l = (uint32_t)(idx - (unsigned long long)t * job.launches);
        const DsPointRadianceTask* task = job.tasks + t;
        o = mk(task->position[0], task->position[1], task->position[2]);
        d = mk(task->direction[0], task->direction[1], task->direction[2]);
        val0 = t * 4096u;
        stream = job.frame0 + l + 1u;
        s.out = idx;
    } else {
        o = mk(job.origins[3 * idx], job.origins[3 * idx + 1], job.origins[3 * idx + 2]);
        d = mk(job.dirs[3 * idx], job.dirs[3 * idx + 1], job.dirs[3 * idx + 2]);
        val0 = job.seedVal0[idx];
        stream = job.stream[idx];
        s.out = idx;
    }
    s.rad = 0.0f;
    const float tHit = intersectBox(sc, o, d);
    if (tHit < 0.0f) return ST_DONE;
    V3 hit = o + tHit * d;
    hit = hit + 0.5f * sc.bbox;
    s.q = hit * sc.texScale;
    s.dir = normalize<true>(d);
    s.seed = tea4(val0, stream);
    s.depth = 0;
    if (job.mode == DS_MODE_SUN_MULTIPLE_SCATTER) s.dir = newDirectionFast(sCdf, sGuide, s.seed, s.dir);
    return loopTopFast<true>(k, s);
}

__device__ __forceinline__ void writeResultFast(const DevScene& sc, const TraceJob& job, const FastState& s, uint32_t& nonfinite)
{
    const float scale = sc.lightIntensity * SUN_TO_SPHERE * s.rad;
    const float r = sc.lightColor.x * scale, g = sc.lightColor.y * scale, b = sc.lightColor.z * scale;
    if (!(fabsf(r + g + b) <= 3.0e38f)) nonfinite++;
    if (job.kind == JOB_RENDER) {
        job.staging[s.out] = make_float4(r, g, b, 1.0f);
    } else if (job.kind == JOB_POINT) {
        job.xOut[s.out] = r;
    } else {
        job.radianceOut[3 * s.out] = r;
        job.radianceOut[3 * s.out + 1] = g;
        job.radianceOut[3 * s.out + 2] = b;
    }
}

/* lane states of k_trace_fast */
enum FastLaneState { F_IDLE = 0, F_SKIP = 1, F_MARCH = 2, F_EVENT = 3, F_DONE = 4 };

/*
 * Warp-level schedule.  Each lane owns one path; a warp cycles through four phases and runs a phase only when
 * enough of its lanes want it, so the (long, divergent) code of each phase executes with many lanes active:
 *   A  retire + regenerate   lanes whose path ended write their sample and pull the next work item from the
 *                            device queue (one warp-aggregated atomic); runs when >= regenMin lanes are free
 *   B  empty-space phase     lanes travelling through empty cells (every new path starts here, and paths
 *                            leaving the cloud return here): step, test the tap cell, jump; runs when >= skipMin
 *                            lanes are in it
 *   C  march phase           lanes inside the cloud: step + density tap + collision test
 *   D  event phase           lanes that collided: next-event estimate + new direction
 * A waiting lane costs nothing but its slot; a phase entered with two lanes costs the whole warp its full
 * instruction stream, which is what the thresholds avoid.  Thresholds are ignored when nothing else can run.
 */
template <bool SKIP>
__global__ void __launch_bounds__(640, 2) k_trace_fast(const DevScene sc, const TraceJob job)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float* sChopped = reinterpret_cast<float*>(smemRaw);
    float* sCdf = sChopped + MIE_N;
    uint16_t* sGuide = reinterpret_cast<uint16_t*>(sCdf + MIE_N);
    uint32_t* sOcc = reinterpret_cast<uint32_t*>(sGuide + GUIDE_N + 2);
    for (int i = threadIdx.x; i < MIE_N; i += blockDim.x) {
        sChopped[i] = sc.chopped[i];
        sCdf[i] = sc.cdf[i];
    }
    for (int i = threadIdx.x; i <= GUIDE_N; i += blockDim.x) sGuide[i] = sc.guide[i];
    for (int i = threadIdx.x; i < sc.occWords; i += blockDim.x) sOcc[i] = sc.occ[i];
    __syncthreads();

    FastConsts k;
    k.stepTs = sc.texScale * sc.step;
    k.half = mk(0.5f + 0.01f * sc.texScale.x, 0.5f + 0.01f * sc.texScale.y, 0.5f + 0.01f * sc.texScale.z);
    k.c1 = sc.mult * sc.step;
    k.nxf = (float)sc.nx;
    k.nyf = (float)sc.ny;
    k.nzf = (float)sc.nz;

    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneLt = (1u << lane) - 1u;

    FastState s;
    s.q = s.dir = mk(0.f, 0.f, 0.f);
    s.rad = s.tau = s.tauStar = 0.f;
    s.seed = 0;
    s.depth = 0;
    s.out = 0;
    int st = F_IDLE;
    bool exhausted = false;
    uint32_t nPaths = 0, nEvents = 0, nSteps = 0, nTaps = 0, nNonfinite = 0;
    float lastDensity = 0.0f;

    for (;;) {
        unsigned mBusy = __ballot_sync(FULL, st == F_MARCH || st == F_EVENT);
        unsigned mSkip = __ballot_sync(FULL, st == F_SKIP);
        unsigned mFree = __ballot_sync(FULL, st == F_DONE || (st == F_IDLE && !exhausted));

        /* ---- A: retire + regenerate ---- */
        if (mFree && (__popc(mFree) >= job.regenMin || (mBusy == 0u && (mSkip == 0u || __popc(mSkip) < job.skipMin)))) {
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                if (st == F_DONE) {
                    writeResultFast(sc, job, s, nNonfinite);
                    st = F_IDLE;
                }
                const unsigned need = __ballot_sync(FULL, st == F_IDLE && !exhausted);
                if (need == 0u) break;
                unsigned long long base = 0;
                const int leader = __ffs(need) - 1;
                if ((int)lane == leader) base = atomicAdd(job.queue, (unsigned long long)__popc(need));
                base = __shfl_sync(FULL, base, leader);
                if (st == F_IDLE && !exhausted) {
                    const unsigned long long idx = base + __popc(need & laneLt);
                    if (idx >= job.total) {
                        exhausted = true;
                    } else {
                        bool valid;
                        const int g = beginItemFast<SKIP>(sc, k, job, sCdf, sGuide, idx, s, valid);
                        st = g == ST_MARCH ? (SKIP ? F_SKIP : F_MARCH) : (g == ST_DONE ? F_DONE : F_IDLE);
                        if (valid) nPaths++;
                    }
                }
                mFree = __ballot_sync(FULL, st == F_DONE || (st == F_IDLE && !exhausted));
                if (__popc(mFree) < job.regenMin) break;
            }
            mSkip = __ballot_sync(FULL, st == F_SKIP);
            mBusy = __ballot_sync(FULL, st == F_MARCH || st == F_EVENT);
            mFree = __ballot_sync(FULL, st == F_DONE || (st == F_IDLE && !exhausted));
        }
        if (mBusy == 0u && mSkip == 0u && mFree == 0u) break; /* every lane idle and the queue exhausted */

        /* ---- B: empty-space phase ---- */
        if (SKIP && mSkip && (__popc(mSkip) >= job.skipMin || mBusy == 0u)) {
#pragma unroll 1
            for (int it = 0; it < job.skipMaxIters; ++it) {
                if (st == F_SKIP) {
                    if (!inBoxTs(k, s.q)) {
                        st = F_DONE;
                    } else {
                        s.q.x = fmaf(s.dir.x, k.stepTs.x, s.q.x);
                        s.q.y = fmaf(s.dir.y, k.stepTs.y, s.q.y);
                        s.q.z = fmaf(s.dir.z, k.stepTs.z, s.q.z);
                        nSteps++;
                        if (!skipEmpty(sc, k, sOcc, s, nSteps)) {
                            nTaps++;
                            lastDensity = tex3D<float>(sc.densityTex, s.q.x, s.q.y, s.q.z);
                            if (lastDensity != 0.0f) {
                                s.tau = fmaf(lastDensity, k.c1, s.tau);
                                st = s.tau > s.tauStar ? F_EVENT : F_MARCH;
                            }
                        }
                    }
                }
                const int nSkip = __popc(__ballot_sync(FULL, st == F_SKIP));
                const bool busyNow = __any_sync(FULL, st == F_MARCH || st == F_EVENT);
                if (nSkip == 0 || (busyNow && nSkip < job.skipKeep)) break;
            }
        }

        /* ---- C: march phase (CU/cloud.cuh:87-104) ---- */
        const int nBusy = __popc(__ballot_sync(FULL, st == F_MARCH || st == F_EVENT));
        if (nBusy) {
#pragma unroll 1
            for (int it = 0; it < job.marchMaxIters; ++it) {
                if (st == F_MARCH) {
                    if (!inBoxTs(k, s.q)) {
                        st = F_DONE;
                    } else {
                        s.q.x = fmaf(s.dir.x, k.stepTs.x, s.q.x);
                        s.q.y = fmaf(s.dir.y, k.stepTs.y, s.q.y);
                        s.q.z = fmaf(s.dir.z, k.stepTs.z, s.q.z);
                        nSteps++;
                        nTaps++;
                        lastDensity = tex3D<float>(sc.densityTex, s.q.x, s.q.y, s.q.z);
                        if (SKIP && lastDensity == 0.0f) {
                            st = F_SKIP; /* left the cloud (or a hole in it): continue in the empty-space phase */
                        } else {
                            s.tau = fmaf(lastDensity, k.c1, s.tau);
                            if (s.tau > s.tauStar) st = F_EVENT;
                        }
                    }
                }
                const int nMarch = __popc(__ballot_sync(FULL, st == F_MARCH));
                if (nMarch * 32 <= nBusy * job.marchKeep32) break;
            }
        }

        /* ---- D: event phase (cloudRadianceMaterials.cu:49-61) ---- */
        if (st == F_EVENT) {
            /* cloud.cuh:99: scatterPos = pos - dir * log(xi / T) / sigma, with log(xi / T) = tau - tauStar */
            const float back = __fdividef(s.tau - s.tauStar, lastDensity * sc.mult);
            s.q.x = fmaf(-back * s.dir.x, sc.texScale.x, s.q.x);
            s.q.y = fmaf(-back * s.dir.y, sc.texScale.y, s.q.y);
            s.q.z = fmaf(-back * s.dir.z, sc.texScale.z, s.q.z);
            if (!inBoxTs(k, s.q)) {
                st = F_DONE;
            } else {
                const float cosLightAngle = -dot(sc.light, s.dir);
                const bool choppedPhase = (job.mode == DS_MODE_SUN_AND_SKY_ALL_SCATTER) ? (s.depth != 1) : (job.mode == DS_MODE_SUN_MULTIPLE_SCATTER);
                const float u = (cosLightAngle + 1.0f) * 0.5f;
                const float phase = choppedPhase ? tableLerp(sChopped, u) : tableLerp(sc.mie, u);
                const float tsun = tex3D<float>(sc.inscatterTex, s.q.x, s.q.y, s.q.z);
                s.rad = fmaf(tsun, phase, s.rad);
                nEvents++;
                if (job.mode == DS_MODE_SUN_SINGLE_SCATTER) {
                    st = F_DONE;
                } else {
                    s.dir = newDirectionFast(sCdf, sGuide, s.seed, s.dir);
                    const int g = loopTopFast<false>(k, s); /* q is the scatter position just verified in-box */
                    st = g == ST_MARCH ? F_MARCH : F_DONE;
                }
            }
        }
    }

    unsigned long long c[5] = {nPaths, nEvents, nSteps, nTaps, nNonfinite};
#pragma unroll
    for (int i = 0; i < 5; i++) {
        unsigned long long v = c[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
        if (lane == 0 && v) atomicAdd(job.stats + i, v);
    }
}

template <>
cudaError_t KernelSet<true>::trace(const DevScene& sc, const TraceJob& job, const LaunchConfig& cfg, cudaStream_t st)
{
    if (cfg.variant == 1) return traceGeneric<true>(sc, job, cfg, st);
    const size_t smem = (size_t)(2 * MIE_N) * 4 + (size_t)(GUIDE_N + 2) * 2 + (size_t)sc.occWords * 4;
    const int threads = cfg.blockThreads > 640 ? 640 : cfg.blockThreads;
    const unsigned long long wantBlocks = (job.total + threads - 1) / threads;
    const unsigned long long maxBlocks = (unsigned long long)cfg.smCount * cfg.blocksPerSm;
    const int blocks = (int)(wantBlocks < maxBlocks ? (wantBlocks ? wantBlocks : 1) : maxBlocks);
    cudaError_t e;
    if (cfg.skipEmpty) {
        e = cudaFuncSetAttribute(k_trace_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_trace_fast<true><<<blocks, threads, smem, st>>>(sc, job);
    } else {
        e = cudaFuncSetAttribute(k_trace_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_trace_fast<false><<<blocks, threads, smem, st>>>(sc, job);
    }
    return cudaGetLastError();
}

template struct KernelSet<true>;

} // namespace dsk
