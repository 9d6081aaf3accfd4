/*
 * ds_kernels.h -- host-callable launchers of the sm_100a kernels (internal to the library).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ds_abi.h"
#include "ds_device.cuh"

namespace dsk {

enum JobKind { JOB_RENDER = 0, JOB_PATHS = 1, JOB_POINT = 2, JOB_ADAPTIVE = 3 };

/* JOB_ADAPTIVE: device-resident radiance collector of the FAST flavour.  One launch per batch of samples:
 * warps draw tickets (sample, `quota` experiment ids), lanes run the experiments, finished paths are summed into the
 * sample's accumulators and the lane that commits applies the reference's convergence rule (RadianceCollector.cpp:108-118,
 * PointRadianceTask.h:23-36); a converged sample is closed and receives no more tickets. */
struct AdaptiveCollector {
    double* sum;               /* [n] sum of the radiance samples */
    double* sumSq;             /* [n] sum of their squares */
    unsigned long long* count; /* [n] experiments accumulated */
    uint32_t* issued;          /* [n] experiment ids handed out */
    uint32_t* flag;            /* [n] 0 = open, 1 = converged, 2 = closed by the experiment cap */
    uint32_t* closed;          /* number of samples with flag != 0 */
    unsigned long long* ticket;
    uint32_t nSamples;
    uint32_t quota;            /* experiments per ticket */
    uint32_t minExperiments;   /* no convergence test before this many (the reference's first test comes after repeat x 100) */
    uint32_t maxExperiments;   /* 0 = unlimited */
    uint32_t zeroMin;          /* zero-radiance samples need more than this many experiments */
    float relCI, absCI;
    /* The reference's runningVariance is the SUM of per-thread Welford M2's: each thread accumulates launches_per_update = L
     * experiments per update and the merge drops the between-thread term (PointRadianceTask.h:54-68), so its expectation is
     * (L - 1) / L of the true total M2.  The collector keeps the exact total and scales it by that factor wherever the
     * reference's statistic is meant (the convergence rule, the reported running_variance). */
    float m2Scale;
};

/* index of the 64-bit work counters on the device */
enum { CNT_PATHS = 0, CNT_EVENTS = 1, CNT_STEPS = 2, CNT_TAPS = 3, CNT_NONFINITE = 4, CNT_COUNT = 8 };

/* One launch of the path-tracing kernel: `total` work items pulled from a device-side queue. */
struct TraceJob {
    int kind;
    int mode;
    unsigned long long total;
    unsigned long long* queue;  /* work counter, zeroed before launch */
    unsigned long long* stats;  /* CNT_COUNT counters */
    int marchKeepQuarters;      /* generic kernel: leave the march phase when marching*4 <= alive*q */
    int marchKeep32;            /* fast kernel: leave the march phase when marching*32 <= alive*k */
    int marchMaxIters;
    int regenMin;      /* fast kernel: regenerate when at least this many lanes of the warp are free */
    int skipMin;       /* fast kernel: enter the empty-space phase with at least this many lanes */
    int skipMaxIters;
    int zeroCheckMin;  /* fast kernel: look at lanes marching through zero density (box exit / open space) when at least this many wait */
    int skipOpenDist;  /* fast kernel: leave the march phase for the empty-space phase when the tap cell is at least this far (cells) from the cloud */
    int specPercent;   /* fast kernel, two-tap pipeline: skip the speculative second tap when lastDensity * c1 * specPercent / 100 exceeds the optical
                          depth left to the collision (0 = always fetch both) */
    /* JOB_RENDER: item = (subframe, 8x4 pixel tile, pixel in tile) */
    float eye[3], U[3], V[3], W[3];
    int width, height, tilesX;
    uint32_t firstSubframe;
    unsigned long long itemsPerSubframe;
    float4* staging; /* [n][H][W] */
    /* primary-ray cache (FAST flavour): pixels whose camera ray reaches an occupied cell, and the
     * number of march steps each of them spends in empty cells before that */
    const uint32_t* hitList;
    uint32_t nHit;
    /* item order over the hit list: regionSize == 0: subframe-major (all hitting pixels of subframe 0, then subframe 1, ...);
     * regionSize > 0: REGION-major -- the hit list is cut into runs of regionSize pixels (k_primary_prepass lists them in
     * 64x64-pixel super-tile order, so a run is a compact patch of the image) and all nSub subframes of a run are handed out
     * before the next run starts.  The ~130k paths in flight on the GPU then share a column of the volume that fits the L2. */
    uint32_t regionSize, nSub;
    const uint32_t* entrySteps; /* [H*W], ENTRY_MISS for pixels that never reach an occupied cell */
    /* JOB_PATHS */
    const float* origins;
    const float* dirs;
    const uint32_t* seedVal0;
    const uint32_t* stream;
    float* radianceOut;
    /* JOB_POINT: item = (thread t, launch l) */
    const DsPointRadianceTask* tasks;
    uint32_t launches;
    uint32_t frame0;
    float* xOut; /* [threads][launches] */
    /* JOB_ADAPTIVE: sample i = tasks[i].position / direction */
    AdaptiveCollector ad;
};

constexpr uint32_t ENTRY_MISS = 0xffffffffu;

struct LevelTable {
    const uint8_t* data[MAX_LEVELS];
    int nx[MAX_LEVELS], ny[MAX_LEVELS], nz[MAX_LEVELS];
    int count;
};

struct DescriptorLayers {
    float scale[10];
    float lod[10];        /* max(0, mipmapLevel) */
    float mipVoxelSize[10];
};

struct LaunchConfig {
    int blockThreads;
    int blocksPerSm;
    int smCount;
    int skipEmpty;
    int variant; /* FAST flavour: 0 = optimised k_trace_fast, 1 = generic k_trace<true, SKIP> (round-1 baseline) */
    int smemCarveout; /* k_trace_fast: cudaFuncAttributePreferredSharedMemoryCarveout in percent, -1 = leave the driver default */
    int marchUnroll; /* k_trace_fast: march steps per vote (1, or 2 = the taps of two steps in flight together) */
};

template <bool FAST>
struct KernelSet {
    static cudaError_t trace(const DevScene& sc, const TraceJob& job, const LaunchConfig& cfg, cudaStream_t st);
    /* per pixel: entry steps / ENTRY_MISS, compacted list of hitting pixels; counts[0] = nHit, counts[1] = total march
     * steps the reference algorithm spends on the missing pixels (per subframe) */
    static cudaError_t primaryPrepass(const DevScene& sc, const TraceJob& cam, uint32_t* entrySteps, uint32_t* hitList,
                                      unsigned long long* counts, cudaStream_t st);
    static cudaError_t bake(const DevScene& sc, uint8_t* out, int skipEmpty, cudaStream_t st);
    /* first half of the neural renderers' frame (disneyCamera.cu pinholeCamera + disneyDescriptorMaterial.cu): per pixel of the
     * rectangle info = {radiance rgb, transmittance, hasScattered}, the collision point (centred) and view direction, the
     * light / view angle and the hasScattered byte */
    static cudaError_t networkInfo(const DevScene& sc, const TraceJob& cam, int rectX, int rectY, int rectW, int rectH, uint32_t stream, float* info,
                                   float* pos, float* dir, float* angle, uint8_t* active, unsigned long long* stats, cudaStream_t st, int tile = 0,
                                   const uint32_t* entrySteps = nullptr);
    /* the hierarchical stencil descriptor (launchDescriptors below picks the instantiation) */
    static cudaError_t descriptors(const DevScene& sc, const struct LevelTable& lv, const struct DescriptorLayers& layers, const float* pos, const float* dir,
                                   uint32_t n, uint8_t* outU8, float* outF32, int32_t* tapIndex, cudaStream_t st, int layerStride, const float* angle,
                                   const uint8_t* active, const uint32_t* gather, cudaTextureObject_t mipTex);
    static cudaError_t generatePoints(const DevScene& sc, uint32_t firstIndex, uint32_t n, uint32_t stream, float* pos, float* dir,
                                      unsigned long long* stats, cudaStream_t st);
};

/* exact-arithmetic helpers (compiled in the -fmad=false translation unit only) */
cudaError_t launchSynth(uint8_t* out, int n, int kind, uint32_t seed, cudaStream_t st);
cudaError_t launchQuantize(const float* in, size_t count, double maxDensity, uint8_t* out, cudaStream_t st);
cudaError_t launchMip(const uint8_t* prev, int pnx, int pny, int pnz, uint8_t* cur, int cnx, int cny, int cnz, cudaStream_t st);
cudaError_t launchCellDistance(const uint32_t* occBits, int ocx, int ocy, int ocz, uint8_t* dist, uint8_t* tmp, cudaStream_t st);
cudaError_t launchBorderCount(const uint8_t* density, int nx, int ny, int nz, uint32_t* count, cudaStream_t st);
cudaError_t launchOccupancy(const uint8_t* density, int nx, int ny, int nz, int shift, int ocx, int ocy, int ocz, uint32_t* bits,
                            cudaStream_t st);
cudaError_t launchUpdateFrame(const float4* staging, const uint32_t* entrySteps, float4* progressive, float4* variance, size_t pixels,
                              uint32_t firstSubframe, uint32_t n, cudaStream_t st);
cudaError_t launchFillMissing(float4* staging, const uint32_t* entrySteps, size_t pixels, cudaStream_t st);
cudaError_t launchTonemap(const float4* progressive, int w, int h, float exposure, float* columns, float* average, uchar4* screen,
                          cudaStream_t st);
cudaError_t launchUnconverged(const float4* progressive, const float4* variance, size_t pixels, uint32_t subframeId, uint32_t* count,
                              cudaStream_t st);
cudaError_t launchExportMoments(const float4* progressive, const float4* variance, size_t pixels, uint32_t n, double* out, cudaStream_t st);
cudaError_t launchImportMoments(const double* in, size_t pixels, uint32_t nTotal, float4* progressive, float4* variance, cudaStream_t st);
/* one 128-row tile of network input in the layout the tensor-core model kernel consumes (ds_mlp.h): [layer 10][K group 58][row 128][4 floats],
 * k = 0..224 densities, 225 the angle, 226 and 227 the constant 1 that carries the biases, 228..231 zero */
constexpr size_t NETWORK_TILE_FLOATS = (size_t)10 * 58 * 128 * 4;

/* layerStride 225: DisneyDescriptor layout [n][10][225]; 226: DisneyNetworkInput layout [n][10][226] whose last element per layer is
 * angle[i] (may be NULL) -- samples with active[i] == 0 (active may be NULL) get all-zero densities; gather (may be NULL): output row i
 * is computed from input sample gather[i]; layerStride 0 / -1 / -2: outF32 is written as 128-row tiles for the tensor-core model kernel, rounded to
 * tf32 (NETWORK_TILE_FLOATS floats each) / bfloat16 / IEEE half (half that), the rows that pad the last tile zeroed; mipTex (0 = none): mip-mapped density texture, the taps then run on the texture units (what the
 * reference's rtTex3DLod does) instead of the exact software fetch */
cudaError_t launchDescriptors(const DevScene& sc, const LevelTable& lv, const DescriptorLayers& layers, const float* pos, const float* dir,
                              uint32_t n, uint8_t* outU8, float* outF32, int32_t* tapIndex, cudaStream_t st, int layerStride = 225,
                              const float* angle = nullptr, const uint8_t* active = nullptr, const uint32_t* gather = nullptr,
                              cudaTextureObject_t mipTex = 0);
/* RG8 array of {a, b} texels from two u8 volumes of the same shape (the fused volume of k_trace_fast) */
cudaError_t launchInterleave(const uint8_t* a, const uint8_t* b, int nx, int ny, int nz, cudaSurfaceObject_t surf, cudaStream_t st);
/* introspection: k_trace_fast's inversion of the chopped-Mie CDF and its half-precision phase sampler on n values in [0, 1) */
cudaError_t launchInvertCdf(const DevScene& sc, const float* val, uint32_t n, float* cosTheta, float* phase, cudaStream_t st);
cudaError_t launchTaskWelford(DsPointRadianceTask* tasks, const float* x, uint32_t nThreads, uint32_t launches, cudaStream_t st);

} // namespace dsk
