/*
 * ds_kernels.cuh -- the estimator kernels, templated on the arithmetic flavour.
 *
 * k_trace is a persistent-threads kernel: every lane owns one light path at a time and, when the
 * path ends, pulls the next work item (pixel sample, explicit ray, or point-radiance experiment)
 * from a device-side queue with one warp-aggregated atomic ("path regeneration").  Per-path RNG
 * streams depend only on the work item, never on the lane that runs it, so results do not depend
 * on scheduling.  Inside a warp the loop alternates between a MARCH phase (one ray-march step per
 * iteration for lanes in free flight) and an EVENT phase (next-event estimate + Mie direction
 * sampling for lanes that collided); the warp leaves the march phase by vote, which keeps both
 * phases populated instead of serialising the two nested data-dependent loops of the reference
 * (CU/cloudRadianceMaterials.cu:28-62 around CU/cloud.cuh:87-104).
 *
 * Shared memory per block: chopped-Mie phase table (16 KiB), chopped-Mie CDF (16 KiB) and the
 * empty-space occupancy bit mask (<= 32 KiB).
 */
#pragma once

#include "ds_kernels.h"

namespace dsk {

enum LaneState { ST_IDLE = 0, ST_MARCH = 1, ST_EVENT = 2, ST_DONE = 3 };

struct PathState {
    V3 p;    /* marching position / scatter position (box-local) */
    V3 dir;
    V3 rad;
    float T; /* transmittance accumulated in the current free flight */
    float xi;
    uint32_t seed;
    int depth;
    unsigned long long out; /* output slot of the work item */
};

/* start of an iteration of the reference's `while (isInBox(pos))` loop (cloudRadianceMaterials.cu:28-35) */
__device__ __forceinline__ int loopTop(const DevScene& sc, PathState& s)
{
    if (!isInBox(sc, s.p)) return ST_DONE;
    s.depth++;
    if (s.depth == MAX_DEPTH) return ST_DONE;
    s.xi = rnd(s.seed); /* getNextScatteringEvent(seed, ...), cloud.cuh:120 */
    s.T = 1.0f;
    return ST_MARCH;
}

/* Decode work item `idx`, generate its ray and intersect the cloud box.  Returns the lane state. */
template <bool FAST>
__device__ __forceinline__ int beginItem(const DevScene& sc, const TraceJob& job, const float* sCdf, unsigned long long idx, PathState& s,
                                         bool& valid)
{
    V3 o, d;
    uint32_t val0, stream;
    valid = true;
    if (job.kind == JOB_RENDER) {
        /* CU/pathTracingCamera.cu:12-21 + CU/cameraCommon.cuh:19-29 */
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
        d = normalize<FAST>(dx * U + dy * V + W);
        val0 = px * 4096u + py; /* cloudRadianceMaterials.cu:21 */
        stream = job.firstSubframe + (uint32_t)sub;
        s.out = sub * (unsigned long long)job.width * job.height + (unsigned long long)py * job.width + px;
    } else if (job.kind == JOB_POINT) {
        /* CU/pointEmissionCamera.cu:20-33: thread t, launch l -> tea<4>(t*4096 + 0, frameId) */
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
    s.rad = mk(0.f, 0.f, 0.f);
    const float tHit = intersectBox(sc, o, d);
    if (tHit < 0.0f) return ST_DONE; /* miss program is a no-op (progressive.cu:44-46) */
    /* closest hit prologue, cloudRadianceMaterials.cu:11-21 */
    V3 hit = o + tHit * d;
    hit = hit + 0.5f * sc.bbox;
    s.p = hit;
    s.dir = normalize<FAST>(d);
    s.seed = tea4(val0, stream);
    s.depth = 0;
    if (job.mode == DS_MODE_SUN_MULTIPLE_SCATTER) {
        s.dir = getNewDirection<FAST>(sCdf, s.seed, s.dir); /* cloudRadianceMaterials.cu:86 */
    }
    return loopTop(sc, s);
}

__device__ __forceinline__ void writeResult(const TraceJob& job, const PathState& s, uint32_t& nonfinite)
{
    const float sum = s.rad.x + s.rad.y + s.rad.z;
    if (!(fabsf(sum) <= 3.0e38f)) nonfinite++;
    if (job.kind == JOB_RENDER) {
        job.staging[s.out] = make_float4(s.rad.x, s.rad.y, s.rad.z, 1.0f); /* pathTracingCamera.cu:20 */
    } else if (job.kind == JOB_POINT) {
        job.xOut[s.out] = s.rad.x; /* pointEmissionCamera.cu:32 */
    } else {
        job.radianceOut[3 * s.out] = s.rad.x;
        job.radianceOut[3 * s.out + 1] = s.rad.y;
        job.radianceOut[3 * s.out + 2] = s.rad.z;
    }
}

template <bool FAST, bool SKIP>
__global__ void __launch_bounds__(512, 2) k_trace(const DevScene sc, const TraceJob job)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float* sChopped = reinterpret_cast<float*>(smemRaw);
    float* sCdf = sChopped + MIE_N;
    uint32_t* sOcc = reinterpret_cast<uint32_t*>(sCdf + MIE_N);
    for (int i = threadIdx.x; i < MIE_N; i += blockDim.x) {
        sChopped[i] = sc.chopped[i];
        sCdf[i] = sc.cdf[i];
    }
    if (SKIP) {
        for (int i = threadIdx.x; i < sc.occWords; i += blockDim.x) sOcc[i] = sc.occ[i];
    }
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneLt = (1u << lane) - 1u;

    PathState s;
    s.p = s.dir = s.rad = mk(0.f, 0.f, 0.f);
    s.T = 1.f;
    s.xi = 0.f;
    s.seed = 0;
    s.depth = 0;
    s.out = 0;
    int st = ST_IDLE;
    bool exhausted = false;
    uint32_t nPaths = 0, nEvents = 0, nSteps = 0, nTaps = 0, nNonfinite = 0;

    for (;;) {
        /* ---- finish + regenerate ---- */
        if (st == ST_DONE) {
            writeResult(job, s, nNonfinite);
            st = ST_IDLE;
        }
        const unsigned need = __ballot_sync(FULL, st == ST_IDLE && !exhausted);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if ((int)lane == leader) base = atomicAdd(job.queue, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (st == ST_IDLE && !exhausted) {
                const unsigned long long idx = base + __popc(need & laneLt);
                if (idx >= job.total) {
                    exhausted = true;
                } else {
                    bool valid;
                    st = beginItem<FAST>(sc, job, sCdf, idx, s, valid);
                    if (valid) nPaths++;
                }
            }
        }
        const unsigned alive = __ballot_sync(FULL, st != ST_IDLE);
        if (alive == 0u) {
            if (__all_sync(FULL, exhausted)) break;
            continue;
        }
        const int nAlive = __popc(alive);

        /* ---- march phase: CU/cloud.cuh:87-104, one step per iteration ---- */
#pragma unroll 1
        for (int it = 0; it < job.marchMaxIters; ++it) {
            if (st == ST_MARCH) {
                if (!isInBox(sc, s.p)) {
                    st = ST_DONE; /* left the box without colliding */
                } else {
                    s.p = s.p + s.dir * sc.step;
                    nSteps++;
                    if (!(SKIP && tapIsEmpty(sc, sOcc, s.p))) {
                        nTaps++;
                        const float density = sampleCloud<FAST>(sc, s.p) * sc.mult;
                        const float extinction = density * sc.step;
                        s.T *= expNeg<FAST>(-extinction);
                        if (s.xi > s.T) {
                            const float lg = logPos<FAST>(FAST ? __fdividef(s.xi, s.T) : s.xi / s.T);
                            const float inv = FAST ? __fdividef(1.0f, density) : 1.0f / density;
                            s.p = s.p - (s.dir * lg) * inv; /* cloud.cuh:99 */
                            st = ST_EVENT;
                        }
                    }
                }
            }
            const int nMarch = __popc(__ballot_sync(FULL, st == ST_MARCH));
            if (nMarch * 4 <= nAlive * job.marchKeepQuarters) break;
        }

        /* ---- event phase: cloudRadianceMaterials.cu:49-61 ---- */
        if (st == ST_EVENT) {
            if (!isInBox(sc, s.p)) {
                st = ST_DONE;
            } else {
                /* getInScattering, cloud.cuh:146-158 */
                const float cosLightAngle = dot(-sc.light, s.dir);
                const bool choppedPhase = (job.mode == DS_MODE_SUN_AND_SKY_ALL_SCATTER) ? (s.depth != 1) : (job.mode == DS_MODE_SUN_MULTIPLE_SCATTER);
                const float u = (cosLightAngle + 1) / 2;
                const float phase = choppedPhase ? tex1dSoft(sChopped, u) : tex1dSoft(sc.mie, u);
                const float tsun = sampleInScatter<FAST>(sc, s.p);
                const V3 li = sc.lightColor * sc.lightIntensity * tsun * phase * SUN_TO_SPHERE;
                s.rad = s.rad + li;
                nEvents++;
                if (job.mode == DS_MODE_SUN_SINGLE_SCATTER) {
                    st = ST_DONE;
                } else {
                    s.dir = getNewDirection<FAST>(sCdf, s.seed, s.dir);
                    st = loopTop(sc, s);
                }
            }
        }
    }

    /* fold the per-lane counters */
    unsigned long long c[5] = {nPaths, nEvents, nSteps, nTaps, nNonfinite};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned long long v = c[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
        if (lane == 0 && v) atomicAdd(job.stats + k, v);
    }
}

template <bool FAST>
cudaError_t traceGeneric(const DevScene& sc, const TraceJob& job, const LaunchConfig& cfg, cudaStream_t st)
{
    const size_t smem = (size_t)(2 * MIE_N + (cfg.skipEmpty ? sc.occWords : 0)) * 4;
    const int threads = cfg.blockThreads > 512 ? 512 : cfg.blockThreads; /* __launch_bounds__(512, 2) */
    unsigned long long wantBlocks = (job.total + threads - 1) / threads;
    /* the options describe threads per SM (block_threads x blocks_per_sm, tuned for k_trace_fast); keep that many resident here */
    const int perSm = (cfg.blockThreads * cfg.blocksPerSm + threads - 1) / threads;
    const unsigned long long maxBlocks = (unsigned long long)cfg.smCount * (perSm < 1 ? 1 : perSm > 2 ? 2 : perSm);
    const int blocks = (int)(wantBlocks < maxBlocks ? (wantBlocks ? wantBlocks : 1) : maxBlocks);
    cudaError_t e;
    if (cfg.skipEmpty) {
        e = cudaFuncSetAttribute(k_trace<FAST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_trace<FAST, true><<<blocks, threads, smem, st>>>(sc, job);
    } else {
        e = cudaFuncSetAttribute(k_trace<FAST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_trace<FAST, false><<<blocks, threads, smem, st>>>(sc, job);
    }
    return cudaGetLastError();
}

/* ---- CU/inScatter.cu:40-66: one thread per voxel, x fastest ---- */
template <bool FAST, bool SKIP>
__global__ void __launch_bounds__(256) k_bake(const DevScene sc, uint8_t* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int z = blockIdx.z;
    if (x >= sc.nx) return;
    const int maxN = max(max(sc.nx, sc.ny), sc.nz);
    const float minScale = fminf(fminf(sc.texScale.x, sc.texScale.y), sc.texScale.z);
    const float invMinScale = 1.0f / minScale;
    V3 pos = mk(((float)x / (float)maxN) * invMinScale, ((float)y / (float)maxN) * invMinScale, ((float)z / (float)maxN) * invMinScale);
    const V3 stepToLight = (-normalize<FAST>(sc.light)) * sc.step;
    const int stepCount = (int)(1 / sc.step);
    float transmittance = 1;
    for (int i = 0; i < stepCount; i++) {
        if (!(SKIP && tapIsEmpty(sc, sc.occ, pos))) {
            const float density = sampleCloud<FAST>(sc, pos) * sc.mult;
            const float extinction = density * sc.step;
            transmittance *= expNeg<FAST>(-extinction);
        }
        pos = pos + stepToLight;
        if (transmittance * 255.f < 1.f) break;
    }
    out[((size_t)z * sc.ny + y) * sc.nx + x] = (uint8_t)(transmittance * 255.f);
}

template <bool FAST>
cudaError_t KernelSet<FAST>::bake(const DevScene& sc, uint8_t* out, int skipEmpty, cudaStream_t st)
{
    dim3 block(256, 1, 1);
    dim3 grid((sc.nx + 255) / 256, sc.ny, sc.nz);
    if (skipEmpty)
        k_bake<FAST, true><<<grid, block, 0, st>>>(sc, out);
    else
        k_bake<FAST, false><<<grid, block, 0, st>>>(sc, out);
    return cudaGetLastError();
}

/* ---- CU/pointGeneratorCamera.cu:20-42 + CU/cloudFirstScatterMaterial.cu:8-28 ---- */
template <bool FAST>
__global__ void __launch_bounds__(128) k_generate_points(const DevScene sc, uint32_t firstIndex, uint32_t n, uint32_t stream, float* pos,
                                                         float* dir, unsigned long long* stats)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t launchID = firstIndex + i;
    uint32_t seed = tea4(launchID, stream);
    uint32_t attempt = 0;
    unsigned long long steps = 0;
    for (;;) {
        attempt++;
        const V3 discNormal = uniformOnSphere<FAST>(seed);
        const float discRadius = sqrtf(3.0f) / 2;
        const V3 position = uniformOnDisc<FAST>(seed, discNormal) * discRadius;
        const V3 origin = position + discNormal * 2;
        const V3 direction = -discNormal;
        const float tHit = intersectBox(sc, origin, direction);
        if (tHit < 0.0f) continue;
        V3 p = origin + tHit * direction;
        p = p + 0.5f * sc.bbox;
        const V3 d = normalize<FAST>(direction);
        uint32_t hitSeed = tea4(launchID * 4096u, attempt);
        const float xi = rnd(hitSeed);
        float T = 1.0f;
        bool scattered = false;
        while (isInBox(sc, p)) {
            p = p + d * sc.step;
            steps++;
            const float density = sampleCloud<FAST>(sc, p) * sc.mult;
            const float extinction = density * sc.step;
            T *= expNeg<FAST>(-extinction);
            if (xi > T) {
                const float lg = logPos<FAST>(xi / T);
                const float inv = 1.0f / density;
                p = p - (d * lg) * inv;
                scattered = true;
                break;
            }
        }
        if (scattered && isInBox(sc, p)) {
            const V3 w = p - 0.5f * sc.bbox;
            pos[3 * i] = w.x;
            pos[3 * i + 1] = w.y;
            pos[3 * i + 2] = w.z;
            dir[3 * i] = -discNormal.x;
            dir[3 * i + 1] = -discNormal.y;
            dir[3 * i + 2] = -discNormal.z;
            break;
        }
    }
    if (stats) atomicAdd(stats + CNT_STEPS, steps);
}

template <bool FAST>
cudaError_t KernelSet<FAST>::generatePoints(const DevScene& sc, uint32_t firstIndex, uint32_t n, uint32_t stream, float* pos, float* dir,
                                            unsigned long long* stats, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    k_generate_points<FAST><<<(n + 127) / 128, 128, 0, st>>>(sc, firstIndex, n, stream, pos, dir, stats);
    return cudaGetLastError();
}

/* ---- CU/disneyCamera.cu:20-36 (pinholeCamera) + CU/disneyDescriptorMaterial.cu:14-46 (sampleDisneyDescriptor) ----
 * one thread per pixel of the rectangle: transmittance of the whole ray, a collision forced inside the cloud, the direct
 * sun radiance there; the float descriptor of the scattered pixels is gathered by k_descriptors afterwards */
template <bool FAST>
__global__ void __launch_bounds__(128) k_network_info(const DevScene sc, const TraceJob cam, int rectX, int rectY, int rectW, int rectH, uint32_t stream,
                                                      int tile, const uint32_t* __restrict__ entrySteps, float* __restrict__ info,
                                                      float* __restrict__ pos, float* __restrict__ dir, float* __restrict__ angleOut,
                                                      uint8_t* __restrict__ active, unsigned long long* stats)
{
    /* tile == 0: one rectangle of renderRect, outputs indexed by the rectangle-local pixel;
     * tile  > 0: the whole frame in one launch (rectW x rectH = frame, rectX = rectY = 0), outputs indexed by the frame pixel; every pixel
     *            behaves as in the launch of its own tile x tile rectangle: rectangle-local launch index in the seed and the stream of
     *            rectangle k = (px / tile) * rectsY + py / tile, the order DisneyRenderer::render visits them (DisneyRenderer.cpp:72-78) */
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inRange = i < (uint32_t)(rectW * rectH);
    uint32_t lx = inRange ? i % (uint32_t)rectW : 0u, ly = inRange ? i / (uint32_t)rectW : 0u;
    const uint32_t px = lx + (uint32_t)rectX, py = ly + (uint32_t)rectY;
    if (tile > 0) {
        const uint32_t rectsY = ((uint32_t)rectH + (uint32_t)tile - 1u) / (uint32_t)tile;
        stream += (px / (uint32_t)tile) * rectsY + py / (uint32_t)tile;
        lx = px % (uint32_t)tile;
        ly = py % (uint32_t)tile;
    }
    unsigned long long steps = 0, events = 0;
    if (inRange) {
    const float dx = (float)px / (float)cam.width * 2.f - 1.f;
    const float dy = (float)py / (float)cam.height * 2.f - 1.f;
    const V3 U = mk(cam.U[0], cam.U[1], cam.U[2]), V = mk(cam.V[0], cam.V[1], cam.V[2]), W = mk(cam.W[0], cam.W[1], cam.W[2]);
    const V3 o = mk(cam.eye[0], cam.eye[1], cam.eye[2]);
    const V3 rayDirection = normalize<FAST>(dx * U + dy * V + W);
    angleOut[i] = acosf(dot(sc.light, rayDirection)); /* disneyCamera.cu:31 */
    float r = 0.f, g = 0.f, b = 0.f, transmittance = 1.0f;
    bool has = false;
    V3 world = mk(0.f, 0.f, 0.f), d = rayDirection;
    const float tHit = intersectBox(sc, o, rayDirection);
    /* entrySteps (frame-wide launches of the FAST flavour): march steps of the camera ray whose taps are known to read 0 (k_primary_prepass);
     * ENTRY_MISS = the ray never reaches an occupied cell: transmittance 1, nothing scatters */
    const uint32_t entry = entrySteps ? entrySteps[i] : 0u;
    if (tHit >= 0.0f && entry != ENTRY_MISS) {
        V3 hit = o + tHit * rayDirection;
        hit = hit + 0.5f * sc.bbox;
        d = normalize<FAST>(rayDirection);
        hit = hit + d * (sc.step * (float)entry); /* zero-density steps change neither the transmittance nor the collision */
        steps += entry + entry;                   /* both marches below would have taken them */
        uint32_t seed = tea4(lx * 4096u + ly, stream);
        /* getNextScatteringEvent(seed, pos, direction, false).transmittance (cloud.cuh:77-122): the march does not stop */
        {
            const float xi = rnd(seed);
            (void)xi; /* the scatter position of this pass is not used */
            V3 p = hit;
            float T = 1.0f;
            while (isInBox(sc, p)) {
                p = p + d * sc.step;
                steps++;
                const float density = sampleCloud<FAST>(sc, p) * sc.mult;
                T *= expNeg<FAST>(-(density * sc.step));
            }
            transmittance = T;
        }
        /* getNextScatteringEvent(1 - rnd(seed) * (1 - transmittance), pos, direction) */
        const float xi = 1 - rnd(seed) * (1 - transmittance);
        V3 p = hit;
        float T = 1.0f;
        bool scattered = false;
        while (isInBox(sc, p)) {
            p = p + d * sc.step;
            steps++;
            const float density = sampleCloud<FAST>(sc, p) * sc.mult;
            T *= expNeg<FAST>(-(density * sc.step));
            if (xi > T) {
                const float lg = logPos<FAST>(xi / T);
                const float inv = 1.0f / density;
                p = p - (d * lg) * inv;
                scattered = true;
                break;
            }
        }
        if (scattered && isInBox(sc, p)) {
            has = true;
            const float cosLightAngle = dot(-sc.light, d);
            const float phase = tex1dSoft(sc.mie, (cosLightAngle + 1) / 2); /* full Mie phase: getInScattering(scatter, direction, false) */
            const float tsun = sampleInScatter<FAST>(sc, p);
            const V3 li = sc.lightColor * sc.lightIntensity * tsun * phase * SUN_TO_SPHERE;
            r = li.x;
            g = li.y;
            b = li.z;
            events++;
            world = p - 0.5f * sc.bbox;
        }
    }
    info[5 * (size_t)i] = r;
    info[5 * (size_t)i + 1] = g;
    info[5 * (size_t)i + 2] = b;
    info[5 * (size_t)i + 3] = transmittance;
    info[5 * (size_t)i + 4] = __uint_as_float(has ? 1u : 0u); /* IntersectionInfo::hasScattered: a bool in a 4-byte slot (rayData.cuh:28-33) */
    active[i] = has ? 1 : 0;
    pos[3 * (size_t)i] = world.x;
    pos[3 * (size_t)i + 1] = world.y;
    pos[3 * (size_t)i + 2] = world.z;
    dir[3 * (size_t)i] = d.x;
    dir[3 * (size_t)i + 1] = d.y;
    dir[3 * (size_t)i + 2] = d.z;
    }
    /* work counters: one atomic per warp and counter */
    unsigned long long paths = inRange ? 1ull : 0ull;
    for (int o = 16; o > 0; o >>= 1) {
        paths += __shfl_down_sync(0xffffffffu, paths, o);
        steps += __shfl_down_sync(0xffffffffu, steps, o);
        events += __shfl_down_sync(0xffffffffu, events, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (paths) atomicAdd(stats + CNT_PATHS, paths);
        if (steps) atomicAdd(stats + CNT_STEPS, steps);
        if (events) atomicAdd(stats + CNT_EVENTS, events);
    }
}

template <bool FAST>
cudaError_t KernelSet<FAST>::networkInfo(const DevScene& sc, const TraceJob& cam, int rectX, int rectY, int rectW, int rectH, uint32_t stream, float* info,
                                         float* pos, float* dir, float* angle, uint8_t* active, unsigned long long* stats, cudaStream_t st, int tile,
                                         const uint32_t* entrySteps)
{
    const int n = rectW * rectH;
    if (n <= 0) return cudaSuccess;
    k_network_info<FAST><<<(n + 127) / 128, 128, 0, st>>>(sc, cam, rectX, rectY, rectW, rectH, stream, tile, entrySteps, info, pos, dir, angle, active, stats);
    return cudaGetLastError();
}

/* ---- CU/DisneyDescriptor.cuh:38-112 + CU/disneyDescriptorCollector.cu:22-29 ---- */

/* rtTex3DLod: clamp lod to [0, L-1], trilinear in floor(lod) and floor(lod)+1, lerp */
__device__ __forceinline__ float tex3dLod(const LevelTable& lv, float u, float v, float w, float lod, int& l0Out)
{
    const int last = lv.count - 1;
    const float l = fminf(fmaxf(lod, 0.0f), (float)last);
    const float lf = floorf(l);
    const int l0 = (int)lf;
    const float t = l - lf;
    l0Out = l0;
    const float a = tex3dSoft(lv.data[l0], lv.nx[l0], lv.ny[l0], lv.nz[l0], u, v, w);
    if (l0 >= last || t == 0.0f) return a;
    const float b = tex3dSoft(lv.data[l0 + 1], lv.nx[l0 + 1], lv.ny[l0 + 1], lv.nz[l0 + 1], u, v, w);
    return fmaf(t, b - a, a);
}

/* DisneyDescriptor.cuh:48-55 */
__device__ __forceinline__ float distanceToBox(const DevScene& sc, V3 pos, float voxelSize)
{
    V3 dist = pos - sc.bbox * 0.5f;
    dist = mk(fabsf(dist.x), fabsf(dist.y), fabsf(dist.z));
    const V3 c = sc.bbox * 0.5f - mk(voxelSize, voxelSize, voxelSize) * 0.5f;
    const V3 boxCorner = mk(fmaxf(c.x, 0.0f), fmaxf(c.y, 0.0f), fmaxf(c.z, 0.0f));
    dist = dist - boxCorner;
    dist = mk(fmaxf(dist.x, 0.0f), fmaxf(dist.y, 0.0f), fmaxf(dist.z, 0.0f));
    return sqrtf(dot(dist, dist));
}

/* one block per sample; thread t < 225 is stencil tap (x, y, z) = (t%5-2, (t/5)%5-2, t/25-2), i.e. the
 * reference's sampleId (z outermost, x innermost, DisneyDescriptor.cuh:89-93); 10 layers per thread */
template <bool FAST>
__global__ void __launch_bounds__(256) k_descriptors(const __grid_constant__ DevScene sc, const __grid_constant__ LevelTable lv,
                                                     const __grid_constant__ DescriptorLayers layers,
                                                     const float* __restrict__ positions, const float* __restrict__ directions, uint32_t n,
                                                     uint8_t* __restrict__ outU8, float* __restrict__ outF32, int32_t* __restrict__ tapIndex,
                                                     int layerStride, const float* __restrict__ angle, const uint8_t* __restrict__ active,
                                                     const uint32_t* __restrict__ gather, cudaTextureObject_t mipTex)
{
    const uint32_t i = blockIdx.x;
    const int t = threadIdx.x;
    const bool tiled = layerStride <= 0, bf16 = layerStride < 0 /* 16-bit tiles: -1 bfloat16, -2 IEEE half */, f16 = layerStride == -2;
    /* tiles: [tile][layer][K group][row 128][16 bytes]; 58 K groups of 4 floats (tf32 operands) or 30 K groups of 8 bf16 per layer */
    const uint32_t KG = bf16 ? 30u : 58u;
    unsigned char* const tileRow = reinterpret_cast<unsigned char*>(outF32) + (size_t)(i >> 7) * (10u * KG * 2048u) + (size_t)(i & 127u) * 16;
    if (tiled) {
        /* constants of the tile layout, and the rows that pad the last tile */
        if (i >= n) {
            for (uint32_t g = t; g < 10u * KG; g += blockDim.x) *reinterpret_cast<uint4*>(tileRow + (size_t)g * 2048) = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
        if (t >= 225 && t < 235) {
            const uint32_t layer = (uint32_t)(t - 225);
            if (bf16) {
                /* K group 28 = k 224..231 (224 density, 225 angle, 226 and 227 the bias-carrying ones), group 29 = k 232..239 */
                unsigned char* p = tileRow + (size_t)(layer * KG + 28) * 2048;
                *reinterpret_cast<uint32_t*>(p + 4) = f16 ? 0x3c003c00u : 0x3f803f80u;
                *reinterpret_cast<uint2*>(p + 8) = make_uint2(0u, 0u);
                *reinterpret_cast<uint4*>(p + 2048) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                /* K group 56 = k 224..227, group 57 = k 228..231 */
                float* p = reinterpret_cast<float*>(tileRow + (size_t)(layer * KG + 56) * 2048);
                p[2] = 1.0f;
                p[3] = 1.0f;
                *reinterpret_cast<float4*>(p + 512) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    if (i >= n || t >= 225) return;
    const uint32_t src = gather ? gather[i] : i; /* input sample of output row i */
    const size_t sampleStride = (size_t)(layerStride > 0 ? layerStride : 0) * 10;
    /* element k of layer `layer` of output row i: a float of the row-major layouts, or an element of the tile rounded to the operand type */
    auto storeF = [&](int layer, int k, float v) {
        if (!tiled) {
            outF32[(size_t)i * sampleStride + (size_t)layer * layerStride + k] = v;
        } else if (bf16) {
            uint16_t h16;
            if (f16) {
                h16 = __half_as_ushort(__float2half_rn(v)); /* densities and angles are far inside the range of a half */
            } else {
                uint32_t u = __float_as_uint(v); /* round to nearest even */
                u += 0x7fffu + ((u >> 16) & 1u);
                h16 = (uint16_t)(u >> 16);
            }
            *reinterpret_cast<uint16_t*>(tileRow + (size_t)((uint32_t)layer * KG + (uint32_t)(k >> 3)) * 2048 + (size_t)(k & 7) * 2) = h16;
        } else {
            /* tf32 (10 mantissa bits), round to nearest: the tensor core would otherwise truncate */
            *reinterpret_cast<uint32_t*>(tileRow + (size_t)((uint32_t)layer * KG + (uint32_t)(k >> 2)) * 2048 + (size_t)(k & 3) * 4) = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
        }
    };
    if (angle && outF32 && t < 10) storeF(t, 225, angle[src]); /* disneyCamera.cu:32-35 */
    if (active && !active[src]) {
        if (outF32)
            for (int layer = 0; layer < 10; layer++) storeF(layer, t, 0.0f);
        return;
    }
    const V3 worldPos = mk(positions[3 * (size_t)src], positions[3 * (size_t)src + 1], positions[3 * (size_t)src + 2]);
    const V3 viewDirection = mk(directions[3 * (size_t)src], directions[3 * (size_t)src + 1], directions[3 * (size_t)src + 2]);
    const V3 eZ = normalize<false>(-sc.light);
    const V3 eX = normalize<false>(cross(eZ, viewDirection));
    const V3 eY = cross(eX, eZ);
    const V3 origin = worldPos + 0.5f * sc.bbox;
    const float x = (float)(t % 5 - 2), y = (float)((t / 5) % 5 - 2), z = (float)(t / 25 - 2);
    /* FAST: the ten layers are independent texture fetches -- unrolled so that they are all in flight together */
#pragma unroll(FAST ? 10 : 1)
    for (int layer = 0; layer < 10; layer++) {
        const V3 offset = (eX * x + eY * y + eZ * z) * layers.scale[layer];
        const V3 pos = origin + offset;
        const V3 uvw = pos * sc.texScale;
        int l0 = 0;
        const float mipVoxelSize = layers.mipVoxelSize[layer];
        const float distance = distanceToBox(sc, pos, mipVoxelSize);
        const float tt = fminf(fmaxf(distance / mipVoxelSize, 0.0f), 1.0f);
        /* a tap more than one mip voxel outside the box fades to exactly zero: lerp(d, 0, 1) = d + 1 * (0 - d) = +0 for every finite d,
         * so its (up to 16) texel reads are skipped -- most taps of the outer layers */
        float density = 0.0f;
        if (tt < 1.0f || tapIndex) {
            if (FAST)
                density = tex3DLod<float>(mipTex, uvw.x, uvw.y, uvw.z, layers.lod[layer]); /* rtTex3DLod, DisneyDescriptor.cuh:41 */
            else
                density = tex3dLod(lv, uvw.x, uvw.y, uvw.z, layers.lod[layer], l0);
            density = density + tt * (0.0f - density); /* lerp(density, 0, t) */
        }
        const size_t o = (size_t)i * 2250 + (size_t)layer * 225 + t;
        if (outU8) outU8[o] = (uint8_t)(density * 255.0f); /* TFromFloat<uint8_t>, DisneyDescriptor.cuh:66-69 */
        if (outF32) storeF(layer, t, density);
        if (tapIndex) {
            const int nx = lv.nx[l0], ny = lv.ny[l0], nz = lv.nz[l0];
            tapIndex[4 * o + 0] = (int)fminf(fmaxf(floorf(uvw.x * (float)nx - 0.5f), -2.0f), (float)nx + 1.0f);
            tapIndex[4 * o + 1] = (int)fminf(fmaxf(floorf(uvw.y * (float)ny - 0.5f), -2.0f), (float)ny + 1.0f);
            tapIndex[4 * o + 2] = (int)fminf(fmaxf(floorf(uvw.z * (float)nz - 0.5f), -2.0f), (float)nz + 1.0f);
            tapIndex[4 * o + 3] = l0;
        }
    }
}

template <bool FAST>
cudaError_t KernelSet<FAST>::descriptors(const DevScene& sc, const LevelTable& lv, const DescriptorLayers& layers, const float* pos, const float* dir,
                                         uint32_t n, uint8_t* outU8, float* outF32, int32_t* tapIndex, cudaStream_t st, int layerStride, const float* angle,
                                         const uint8_t* active, const uint32_t* gather, cudaTextureObject_t mipTex)
{
    if (n == 0) return cudaSuccess;
    if (FAST && !mipTex) return cudaErrorInvalidValue; /* the FAST instantiation samples the mip-mapped texture */
    const uint32_t blocks = layerStride <= 0 ? (n + 127u) / 128u * 128u : n; /* tiled output: whole tiles */
    k_descriptors<FAST><<<blocks, 256, 0, st>>>(sc, lv, layers, pos, dir, n, outU8, outF32, tapIndex, layerStride, angle, active, gather, mipTex);
    return cudaGetLastError();
}

} // namespace dsk
