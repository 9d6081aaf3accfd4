/*
 * ds_abi.cu -- implementation of include/ds_abi.h: context, volume, scene, progressive renderer and
 * dataset-generation entry points on top of the sm_100a kernels.  No CPU fallback: every entry point
 * that computes needs a CUDA device and fails loudly without one.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "ds_kernels.h"
#include "ds_mlp.h"

extern "C" {
extern const unsigned char ds_mie_blob[];     /* deepestscatter_b200/data/mie_tables.f32 (ds_mie_blob.S) */
extern const unsigned char ds_mie_blob_end[];
}

using namespace dsk;

struct DsContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = true;
    cudaDeviceProp prop{};
    std::string err;
    std::string desc;
    std::map<std::string, int> opt;

    /* volume */
    std::vector<uint8_t*> levels;
    std::vector<int> lnx, lny, lnz;
    uint8_t* inscatter = nullptr;
    bool baked = false;
    cudaArray_t densityArr = nullptr, inscatterArr = nullptr;
    cudaTextureObject_t densityTex = 0, inscatterTex = 0;
    cudaArray_t fusedArr = nullptr; /* RG8 {density, sun transmittance}: what k_trace_fast reads when option fused_volume is on */
    cudaTextureObject_t fusedTex = 0;
    bool fusedValid = false;
    cudaMipmappedArray_t densityMip = nullptr; /* the whole mip chain as one mip-mapped array (FAST descriptor gather of the neural renderer) */
    cudaTextureObject_t densityMipTex = 0;
    uint32_t* occ = nullptr;
    uint8_t* cellDist = nullptr;
    uint8_t* cellEscape = nullptr; /* per cell: bit o set when every cell of octant o seen from the cell is empty (k_trace_fast, emptySteps) */
    int occShift = 0, ocx = 0, ocy = 0, ocz = 0, occWords = 0;
    int borderEmpty = 0;

    /* scene */
    DsSceneParams params{};
    bool sceneSet = false;
    float derived[12] = {0};
    float* mie = nullptr;     /* 3 * 4096 floats: mie, chopped, cdf */
    uint16_t* guide = nullptr; /* GUIDE_A_N + GUIDE_B_N guide entries (DevScene::guideA / guideB), then the chopped phase sampler as MIE_N halves */

    /* frame */
    int width = 0, height = 0;
    float4 *progressive = nullptr, *variance = nullptr, *staging = nullptr;
    size_t stagingSubframes = 0;
    uchar4* screen = nullptr;
    float *columns = nullptr, *average = nullptr;
    uint32_t* unconv = nullptr;

    /* primary-ray cache of the FAST flavour (k_primary_prepass): valid for one (volume, camera, frame size) */
    uint32_t *entrySteps = nullptr, *hitList = nullptr;
    unsigned long long* primaryCounts = nullptr; /* device: nHit, missSteps */
    bool primaryValid = false;
    DsCamera primaryCam{};
    uint32_t nHit = 0;
    unsigned long long missSteps = 0;
    unsigned long long extraPaths = 0, extraSteps = 0; /* work of untraced (missing) pixels, added to the counters */

    /* ds_render_subframes_host: the frame buffers travel on a second stream while the first trace kernel runs */
    cudaStream_t copyStream = nullptr;
    cudaEvent_t copyFence = nullptr, copyDone = nullptr;
    bool uploadPending = false; /* copyDone must be waited for before the accumulation buffers are touched */

    /* counters + queue */
    unsigned long long* stats = nullptr; /* CNT_COUNT */
    unsigned long long* queue = nullptr;

    /* launch accounting */
    unsigned long long launches = 0;       /* kernels launched by this context since the last reset */
    std::vector<cudaEvent_t> traceEvents;  /* start/stop pairs around k_trace launches ("profile_events") */
    size_t traceEventsUsed = 0;

    /* multi-GPU reduce (ds_comm_*, ds_frame_reduce) */
    void* comm = nullptr;       /* ncclComm_t */
    int commRanks = 0, commRank = -1;
    double* moments = nullptr;  /* 8 * W * H doubles */
    size_t momentsPixels = 0;

    /* scratch */
    void* scratch[8] = {nullptr};
    size_t scratchSize[8] = {0};

    /* radiance-predicting network of the neural renderer (ds_disney_model_load) */
    DisneyModelDev model;
    void* mlpScratch[7] = {nullptr}; /* network-input upload, predictions, compacted row indices + count, frame result, network-input tiles,
                                        entry steps + hit list of the neural renderer's primary-ray cache */
    size_t mlpScratchSize[7] = {0};
    bool disneyPrimaryValid = false; /* entry steps valid for (disneyPrimaryCam, disneyPrimaryW x H, the current volume and scene) */
    DsCamera disneyPrimaryCam{};
    uint32_t disneyPrimaryW = 0, disneyPrimaryH = 0;
};

static thread_local std::string g_createError;

#define DS_FAIL(ctx, code, ...)                         \
    do {                                                \
        char _b[512];                                   \
        snprintf(_b, sizeof(_b), __VA_ARGS__);          \
        (ctx)->err = _b;                                \
        return (code);                                  \
    } while (0)

#define DS_CUDA(ctx, call)                                                                              \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) DS_FAIL(ctx, DS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

/* ds_frame_upload copies on a second stream; every entry point orders the context's stream behind that copy before it does anything else --
 * except ds_render_subframes, which defers the wait to its first accumulation kernel so that the copy runs beside the trace kernel */
static inline void settleUpload(DsContext* ctx)
{
    if (ctx->uploadPending) {
        cudaStreamWaitEvent(ctx->stream, ctx->copyDone, 0);
        ctx->uploadPending = false;
    }
}
#define DS_CHECK_CTX_DEFER(ctx)        \
    if (!(ctx)) return DS_ERR_INVALID; \
    cudaSetDevice((ctx)->device)
#define DS_CHECK_CTX(ctx)   \
    DS_CHECK_CTX_DEFER(ctx); \
    settleUpload(ctx)

/* With option "profile_events" on, the device time of a kernel (CUDA events on the context's stream) is published as the
 * read-only option `key` in microseconds: what bench.py's roofline legs divide the algorithmic bytes by. */
struct KernelTimer {
    DsContext* ctx;
    const char* key;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    KernelTimer(DsContext* c, const char* k) : ctx(c), key(k)
    {
        if (ctx->opt["profile_events"] && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, ctx->stream);
    }
    void stop()
    {
        if (!e1) return;
        float ms = 0.0f;
        cudaEventRecord(e1, ctx->stream);
        if (cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) ctx->opt[key] = (int)(ms * 1000.0f + 0.5f);
    }
    ~KernelTimer()
    {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

static int ensureScratch(DsContext* ctx, int slot, size_t bytes)
{
    if (ctx->scratchSize[slot] >= bytes) return DS_OK;
    if (ctx->scratch[slot]) cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratchSize[slot] = 0;
    DS_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], bytes));
    ctx->scratchSize[slot] = bytes;
    return DS_OK;
}

static int ensureMlpScratch(DsContext* ctx, int slot, size_t bytes)
{
    if (ctx->mlpScratchSize[slot] >= bytes) return DS_OK;
    if (ctx->mlpScratch[slot]) cudaFree(ctx->mlpScratch[slot]);
    ctx->mlpScratch[slot] = nullptr;
    ctx->mlpScratchSize[slot] = 0;
    DS_CUDA(ctx, cudaMalloc(&ctx->mlpScratch[slot], bytes));
    ctx->mlpScratchSize[slot] = bytes;
    return DS_OK;
}

static void freeDisneyModel(DsContext* ctx)
{
    DisneyModelDev& m = ctx->model;
    cudaFree(m.wT);
    cudaFree(m.bias);
    cudaFree(m.w4b4);
    cudaFree(m.stream);
    freeMlpProgram(m.program);
    freeMlpProgram(m.programBf16);
    cudaFree(m.streamBf16);
    cudaFree(m.streamF16);
    cudaFree(m.error);
    cudaFree(m.prof);
    m = DisneyModelDev();
}

static void freeVolume(DsContext* ctx)
{
    for (uint8_t* p : ctx->levels) cudaFree(p);
    ctx->levels.clear();
    ctx->lnx.clear();
    ctx->lny.clear();
    ctx->lnz.clear();
    if (ctx->inscatter) cudaFree(ctx->inscatter);
    ctx->inscatter = nullptr;
    ctx->baked = false;
    if (ctx->densityTex) cudaDestroyTextureObject(ctx->densityTex);
    if (ctx->inscatterTex) cudaDestroyTextureObject(ctx->inscatterTex);
    ctx->densityTex = ctx->inscatterTex = 0;
    if (ctx->densityMipTex) cudaDestroyTextureObject(ctx->densityMipTex);
    ctx->densityMipTex = 0;
    if (ctx->densityMip) cudaFreeMipmappedArray(ctx->densityMip);
    ctx->densityMip = nullptr;
    if (ctx->fusedTex) cudaDestroyTextureObject(ctx->fusedTex);
    ctx->fusedTex = 0;
    ctx->fusedValid = false;
    if (ctx->densityArr) cudaFreeArray(ctx->densityArr);
    if (ctx->inscatterArr) cudaFreeArray(ctx->inscatterArr);
    if (ctx->fusedArr) cudaFreeArray(ctx->fusedArr);
    ctx->densityArr = ctx->inscatterArr = ctx->fusedArr = nullptr;
    if (ctx->occ) cudaFree(ctx->occ);
    ctx->occ = nullptr;
    if (ctx->cellDist) cudaFree(ctx->cellDist);
    ctx->cellDist = nullptr;
    if (ctx->cellEscape) cudaFree(ctx->cellEscape);
    ctx->cellEscape = nullptr;
}

static void freeFrame(DsContext* ctx)
{
    cudaFree(ctx->progressive);
    cudaFree(ctx->variance);
    cudaFree(ctx->staging);
    cudaFree(ctx->screen);
    cudaFree(ctx->columns);
    cudaFree(ctx->average);
    cudaFree(ctx->unconv);
    ctx->progressive = ctx->variance = ctx->staging = nullptr;
    ctx->screen = nullptr;
    ctx->columns = ctx->average = nullptr;
    ctx->unconv = nullptr;
    ctx->stagingSubframes = 0;
    ctx->width = ctx->height = 0;
    cudaFree(ctx->entrySteps);
    cudaFree(ctx->hitList);
    cudaFree(ctx->primaryCounts);
    ctx->entrySteps = ctx->hitList = nullptr;
    ctx->primaryCounts = nullptr;
    ctx->primaryValid = false;
}

/* DG/Mie.cpp:8206-8282: phase samplers = table / mean(table); integral = running sum of table / sum(table) */
static void buildMieSamplers(const float* mieRaw, const float* choppedRaw, float* out /* 3*4096 */)
{
    auto phase = [](const float* src, float* dst) {
        float average = 0;
        for (int i = 0; i < MIE_N; i++) average += src[i];
        average /= MIE_N;
        for (int i = 0; i < MIE_N; i++) dst[i] = src[i] / average;
    };
    phase(mieRaw, out);
    phase(choppedRaw, out + MIE_N);
    float sum = 0;
    for (int i = 0; i < MIE_N; i++) sum += choppedRaw[i];
    float integral = 0;
    for (int i = 0; i < MIE_N; i++) {
        integral += choppedRaw[i] / sum;
        out[2 * MIE_N + i] = integral;
    }
}

static int makeTexture(DsContext* ctx, const uint8_t* linear, int nx, int ny, int nz, cudaArray_t* arr, cudaTextureObject_t* tex)
{
    if (*tex) {
        cudaDestroyTextureObject(*tex);
        *tex = 0;
    }
    if (!*arr) {
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
        DS_CUDA(ctx, cudaMalloc3DArray(arr, &cd, make_cudaExtent(nx, ny, nz)));
    }
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr((void*)linear, (size_t)nx, (size_t)nx, (size_t)ny);
    cp.dstArray = *arr;
    cp.extent = make_cudaExtent(nx, ny, nz);
    cp.kind = cudaMemcpyDeviceToDevice;
    DS_CUDA(ctx, cudaMemcpy3DAsync(&cp, ctx->stream));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = *arr;
    cudaTextureDesc td = {};
    /* VDBCloud::createSamplerForBuffer3D (VDBCloud.cpp:119-137) */
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    DS_CUDA(ctx, cudaCreateTextureObject(tex, &rd, &td, nullptr));
    return DS_OK;
}

/* The volume k_trace_fast marches through when option fused_volume is on: one block-linear RG8 array whose texel is {density, sun
 * transmittance}, filled on the device from the two u8 volumes (surface writes) after every bake.  Same sampler as makeTexture. */
static int ensureFusedTexture(DsContext* ctx)
{
    if (ctx->fusedValid) return DS_OK;
    const int nx = ctx->lnx[0], ny = ctx->lny[0], nz = ctx->lnz[0];
    if (!ctx->fusedArr) {
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar2>();
        DS_CUDA(ctx, cudaMalloc3DArray(&ctx->fusedArr, &cd, make_cudaExtent(nx, ny, nz), cudaArraySurfaceLoadStore));
    }
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = ctx->fusedArr;
    cudaSurfaceObject_t surf = 0;
    DS_CUDA(ctx, cudaCreateSurfaceObject(&surf, &rd));
    cudaError_t e = launchInterleave(ctx->levels[0], ctx->inscatter, nx, ny, nz, surf, ctx->stream);
    ctx->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaDestroySurfaceObject(surf);
    DS_CUDA(ctx, e);
    if (!ctx->fusedTex) {
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        DS_CUDA(ctx, cudaCreateTextureObject(&ctx->fusedTex, &rd, &td, nullptr));
    }
    ctx->fusedValid = true;
    return DS_OK;
}

/* rtTex3DLod's texture (DisneyDescriptor.cuh:38-42; sampler of VDBCloud.cpp:119-137 with the mip chain of Resources.cpp:169-209): the u8 levels
 * this library built, copied into one mip-mapped array; trilinear within a level, linear between levels, clamp, normalised.  Built on first
 * use, dropped with the volume. */
static int ensureMipTexture(DsContext* ctx)
{
    if (ctx->densityMipTex) return DS_OK;
    const int count = (int)ctx->levels.size();
    if (count < 1) DS_FAIL(ctx, DS_ERR_STATE, "no volume");
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
    DS_CUDA(ctx, cudaMallocMipmappedArray(&ctx->densityMip, &cd, make_cudaExtent(ctx->lnx[0], ctx->lny[0], ctx->lnz[0]), (unsigned)count));
    for (int l = 0; l < count; ++l) {
        cudaArray_t level = nullptr;
        DS_CUDA(ctx, cudaGetMipmappedArrayLevel(&level, ctx->densityMip, (unsigned)l));
        cudaExtent ext{};
        cudaChannelFormatDesc got{};
        DS_CUDA(ctx, cudaArrayGetInfo(&got, &ext, nullptr, level));
        if ((int)ext.width != ctx->lnx[l] || (int)std::max<size_t>(ext.height, 1) != ctx->lny[l] || (int)std::max<size_t>(ext.depth, 1) != ctx->lnz[l])
            DS_FAIL(ctx, DS_ERR_CUDA, "mip level %d: array is %zux%zux%zu, expected %dx%dx%d", l, ext.width, ext.height, ext.depth, ctx->lnx[l], ctx->lny[l],
                    ctx->lnz[l]);
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr((void*)ctx->levels[l], (size_t)ctx->lnx[l], (size_t)ctx->lnx[l], (size_t)ctx->lny[l]);
        cp.dstArray = level;
        cp.extent = make_cudaExtent(ctx->lnx[l], ctx->lny[l], ctx->lnz[l]);
        cp.kind = cudaMemcpyDeviceToDevice;
        DS_CUDA(ctx, cudaMemcpy3DAsync(&cp, ctx->stream));
    }
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeMipmappedArray;
    rd.res.mipmap.mipmap = ctx->densityMip;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.mipmapFilterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    td.minMipmapLevelClamp = 0.0f;
    td.maxMipmapLevelClamp = (float)(count - 1);
    td.disableTrilinearOptimization = 1; /* the full linear blend between levels for every LOD fraction (and the same one in every process) */
    DS_CUDA(ctx, cudaCreateTextureObject(&ctx->densityMipTex, &rd, &td, nullptr));
    return DS_OK;
}

/* the descriptor gather of the neural renderer runs on the texture units in the FAST flavour (what the reference's rtTex3DLod does); the
 * dataset collectors always use the exact software fetch.  Option descriptor_hw: -1 = that rule, 0 = never, 1 = also the float collector */
static int descriptorTexture(DsContext* ctx, bool networkInput, cudaTextureObject_t* tex)
{
    *tex = 0;
    const int opt = ctx->opt["descriptor_hw"];
    const bool want = opt == 1 || (opt < 0 && networkInput && ctx->opt["precision"] == DS_PRECISION_FAST);
    if (!want || ctx->levels.size() < 2) return DS_OK;
    int rc = ensureMipTexture(ctx);
    if (rc) return rc;
    *tex = ctx->densityMipTex;
    return DS_OK;
}

static void fillDevScene(DsContext* ctx, DevScene& sc)
{
    memset(&sc, 0, sizeof(sc));
    sc.density = ctx->levels.empty() ? nullptr : ctx->levels[0];
    sc.inscatter = ctx->inscatter;
    sc.densityTex = ctx->densityTex;
    sc.inscatterTex = ctx->inscatterTex;
    sc.fusedTex = 0; /* runTrace sets it */
    if (!ctx->levels.empty()) {
        sc.nx = ctx->lnx[0];
        sc.ny = ctx->lny[0];
        sc.nz = ctx->lnz[0];
    }
    const float* d = ctx->derived;
    sc.bbox = V3{d[0], d[1], d[2]};
    sc.texScale = V3{d[3], d[4], d[5]};
    sc.mult = d[6];
    sc.step = ctx->params.sample_step;
    sc.minRay = ctx->params.minimal_ray_distance;
    sc.light = V3{d[9], d[10], d[11]};
    sc.lightColor = V3{ctx->params.light_color[0], ctx->params.light_color[1], ctx->params.light_color[2]};
    sc.lightIntensity = ctx->params.light_intensity;
    sc.mie = ctx->mie;
    sc.chopped = ctx->mie + MIE_N;
    sc.cdf = ctx->mie + 2 * MIE_N;
    sc.occ = ctx->occ;
    sc.occShift = ctx->occShift;
    sc.ocx = ctx->ocx;
    sc.ocy = ctx->ocy;
    sc.ocz = ctx->ocz;
    sc.occWords = ctx->occWords;
    sc.cellDist = ctx->cellDist;
    sc.cellEscape = ctx->opt["escape_octants"] ? ctx->cellEscape : nullptr;
    sc.guideA = ctx->guide;
    sc.guideB = ctx->guide + GUIDE_A_N;
    sc.choppedHalf = ctx->guide + GUIDE_A_N + GUIDE_B_N;
    sc.borderEmpty = ctx->borderEmpty;
}

/* VDBCloud::setupVolumeVariables / setupVariables (VDBCloud.cpp:88-117) + Sun (SceneDescription.h:17-19) */
static void computeDerived(DsContext* ctx)
{
    float* d = ctx->derived;
    if (!ctx->levels.empty()) {
        const float fx = (float)ctx->lnx[0], fy = (float)ctx->lny[0], fz = (float)ctx->lnz[0];
        const float maxSize = std::max({fx, fy, fz});
        d[0] = fx / maxSize;
        d[1] = fy / maxSize;
        d[2] = fz / maxSize;
        d[3] = maxSize / fx;
        d[4] = maxSize / fy;
        d[5] = maxSize / fz;
        const size_t mx = (size_t)std::max({ctx->lnx[0], ctx->lny[0], ctx->lnz[0]});
        d[7] = ctx->params.cloud_size_m / mx;
    }
    d[6] = ctx->params.cloud_size_m / ctx->params.mean_free_path_m;
    d[8] = d[7] / ctx->params.mean_free_path_m;
    const float lx = ctx->params.light_direction[0], ly = ctx->params.light_direction[1], lz = ctx->params.light_direction[2];
    const float invLen = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);
    d[9] = lx * invLen;
    d[10] = ly * invLen;
    d[11] = lz * invLen;
}

static int finishVolume(DsContext* ctx, int buildMips)
{
    const int nx = ctx->lnx[0], ny = ctx->lny[0], nz = ctx->lnz[0];
    if (buildMips) {
        /* Resources.cpp:110-115 level count; optix getMipLevelSize = max(1, n >> level) */
        int maxSize = std::max({nx, ny, nz});
        int levelCount = 1;
        while (maxSize /= 2) levelCount++;
        if (levelCount > MAX_LEVELS) DS_FAIL(ctx, DS_ERR_INVALID, "volume too large: %d mip levels", levelCount);
        for (int l = 1; l < levelCount; l++) {
            const int cx = std::max(1, nx >> l), cy = std::max(1, ny >> l), cz = std::max(1, nz >> l);
            uint8_t* p = nullptr;
            DS_CUDA(ctx, cudaMalloc(&p, (size_t)cx * cy * cz));
            ctx->levels.push_back(p);
            ctx->lnx.push_back(cx);
            ctx->lny.push_back(cy);
            ctx->lnz.push_back(cz);
            DS_CUDA(ctx, launchMip(ctx->levels[l - 1], ctx->lnx[l - 1], ctx->lny[l - 1], ctx->lnz[l - 1], p, cx, cy, cz, ctx->stream));
        }
    }
    /* occupancy mask: smallest power-of-two cell with <= 2^18 cells (<= 32 KiB of bits in shared memory) */
    int shift = 0;
    for (;; shift++) {
        const long long c = 1ll << shift;
        const long long cells = ((nx + c - 1) / c) * ((ny + c - 1) / c) * ((nz + c - 1) / c);
        if (cells <= (1ll << 18)) break;
    }
    const int c = 1 << shift;
    ctx->occShift = shift;
    ctx->ocx = (nx + c - 1) / c;
    ctx->ocy = (ny + c - 1) / c;
    ctx->ocz = (nz + c - 1) / c;
    ctx->occWords = (ctx->ocx * ctx->ocy * ctx->ocz + 31) / 32;
    DS_CUDA(ctx, cudaMalloc(&ctx->occ, (size_t)ctx->occWords * 4));
    DS_CUDA(ctx, cudaMemsetAsync(ctx->occ, 0, (size_t)ctx->occWords * 4, ctx->stream));
    DS_CUDA(ctx, launchOccupancy(ctx->levels[0], nx, ny, nz, shift, ctx->ocx, ctx->ocy, ctx->ocz, ctx->occ, ctx->stream));
    {
        const size_t cells = (size_t)ctx->ocx * ctx->ocy * ctx->ocz;
        DS_CUDA(ctx, cudaMalloc(&ctx->cellDist, cells));
        int rcs = ensureScratch(ctx, 7, cells);
        if (rcs) return rcs;
        DS_CUDA(ctx, launchCellDistance(ctx->occ, ctx->ocx, ctx->ocy, ctx->ocz, ctx->cellDist, (uint8_t*)ctx->scratch[7], ctx->stream));
    }
    int rc = makeTexture(ctx, ctx->levels[0], nx, ny, nz, &ctx->densityArr, &ctx->densityTex);
    if (rc) return rc;
    DS_CUDA(ctx, cudaMalloc(&ctx->inscatter, (size_t)nx * ny * nz));
    DS_CUDA(ctx, cudaMemsetAsync(ctx->inscatter, 0, (size_t)nx * ny * nz, ctx->stream));
    computeDerived(ctx);
    {
        /* are all face voxels zero?  (decides whether clamped taps outside the grid can be skipped) */
        int rcs = ensureScratch(ctx, 6, sizeof(uint32_t));
        if (rcs) return rcs;
        uint32_t h = 1;
        DS_CUDA(ctx, cudaMemsetAsync(ctx->scratch[6], 0, sizeof(uint32_t), ctx->stream));
        DS_CUDA(ctx, launchBorderCount(ctx->levels[0], nx, ny, nz, (uint32_t*)ctx->scratch[6], ctx->stream));
        DS_CUDA(ctx, cudaMemcpyAsync(&h, ctx->scratch[6], sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->borderEmpty = h == 0 ? 1 : 0;
    }
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->borderEmpty) {
        /* escape octants: bit o = (dx > 0) | (dy > 0) << 1 | (dz > 0) << 2 of cell c is set when c and every cell on its far side in all
         * three directions of octant o are empty.  A ray in c with those direction signs only visits such cells and then leaves a grid
         * whose faces are zero: every tap up to the box exit reads 0.  Three suffix-AND sweeps per octant over <= 2^18 cells, on the host. */
        const int ox = ctx->ocx, oy = ctx->ocy, oz = ctx->ocz;
        const size_t cells = (size_t)ox * oy * oz;
        std::vector<uint32_t> bits(ctx->occWords);
        DS_CUDA(ctx, cudaMemcpy(bits.data(), ctx->occ, (size_t)ctx->occWords * 4, cudaMemcpyDeviceToHost));
        std::vector<uint8_t> esc(cells, 0), r(cells);
        for (int o = 0; o < 8; o++) {
            for (size_t i = 0; i < cells; i++) r[i] = ((bits[i >> 5] >> (i & 31)) & 1u) ? 0 : 1;
            const int sx = (o & 1) ? 1 : -1, sy = (o & 2) ? 1 : -1, sz = (o & 4) ? 1 : -1;
            auto at = [&](int x, int y, int z) -> uint8_t& { return r[((size_t)z * oy + y) * ox + x]; };
            for (int z = 0; z < oz; z++)
                for (int y = 0; y < oy; y++)
                    for (int i = 1; i < ox; i++) {
                        const int x = sx > 0 ? ox - 1 - i : i;
                        at(x, y, z) &= at(x + sx, y, z);
                    }
            for (int z = 0; z < oz; z++)
                for (int i = 1; i < oy; i++) {
                    const int y = sy > 0 ? oy - 1 - i : i;
                    for (int x = 0; x < ox; x++) at(x, y, z) &= at(x, y + sy, z);
                }
            for (int i = 1; i < oz; i++) {
                const int z = sz > 0 ? oz - 1 - i : i;
                for (int y = 0; y < oy; y++)
                    for (int x = 0; x < ox; x++) at(x, y, z) &= at(x, y, z + sz);
            }
            for (size_t i = 0; i < cells; i++) esc[i] |= (uint8_t)(r[i] << o);
        }
        DS_CUDA(ctx, cudaMalloc(&ctx->cellEscape, cells));
        DS_CUDA(ctx, cudaMemcpy(ctx->cellEscape, esc.data(), cells, cudaMemcpyHostToDevice));
    }
    return DS_OK;
}

static int beginVolume(DsContext* ctx, int nx, int ny, int nz)
{
    if (nx <= 0 || ny <= 0 || nz <= 0) DS_FAIL(ctx, DS_ERR_INVALID, "bad volume size %dx%dx%d", nx, ny, nz);
    freeVolume(ctx);
    ctx->primaryValid = false;
    ctx->disneyPrimaryValid = false;
    ctx->opt["volume_generation"]++; /* read by the importer's cache (host/CloudImporter.hpp) */
    uint8_t* p = nullptr;
    DS_CUDA(ctx, cudaMalloc(&p, (size_t)nx * ny * nz));
    ctx->levels.push_back(p);
    ctx->lnx.push_back(nx);
    ctx->lny.push_back(ny);
    ctx->lnz.push_back(nz);
    return DS_OK;
}

static LaunchConfig launchConfig(DsContext* ctx)
{
    LaunchConfig cfg;
    cfg.blockThreads = ctx->opt["block_threads"];
    cfg.blocksPerSm = ctx->opt["blocks_per_sm"];
    cfg.smCount = ctx->prop.multiProcessorCount;
    cfg.skipEmpty = ctx->opt["skip_empty"];
    cfg.variant = ctx->opt["variant"];
    cfg.smemCarveout = ctx->opt["smem_carveout"];
    cfg.marchUnroll = ctx->opt["march_unroll"];
    if (cfg.marchUnroll == 0) {
        /* auto: the two-tap pipeline (second tap guarded by spec_percent, DS_ISSUE_TAPS) pays while the taps are L2 hits (C2:
         * 268 MB of volumes, 99.7 % L2 hit rate: 1177 vs ~1000 Mpaths/s single-tap); once the volumes dwarf the L2 every wasted
         * tap is a DRAM transaction and a march step is two voxels long, so pairs share no sectors (C4, 2.1 GB: 629 single-tap vs
         * 595 guarded pairs vs 419 unconditional pairs, profiles/r02c_*, r02j_*, r02g_*) */
        const size_t volumeBytes = ctx->levels.empty() ? 0 : 2 * (size_t)ctx->lnx[0] * ctx->lny[0] * ctx->lnz[0];
        /* with the fused RG8 volume the pair wins there as well (C4: 561 vs 544 Mpaths/s at 16 spp per launch, profiles/r02r_*) */
        const bool fused = ctx->opt["fused_volume"] != 0 && ctx->opt["precision"] == DS_PRECISION_FAST;
        cfg.marchUnroll = (volumeBytes > 4 * (size_t)ctx->prop.l2CacheSize && !fused) ? 1 : 2;
    }
    return cfg;
}

static int runTrace(DsContext* ctx, TraceJob& job)
{
    DevScene sc;
    fillDevScene(ctx, sc);
    job.queue = ctx->queue;
    job.stats = ctx->stats;
    job.marchKeepQuarters = ctx->opt["march_keep_quarters"];
    job.marchMaxIters = ctx->opt["march_max_iters"];
    job.marchKeep32 = ctx->opt["march_keep32"];
    job.regenMin = ctx->opt["regen_min"];
    job.skipMin = ctx->opt["skip_min"];
    job.skipMaxIters = ctx->opt["skip_max_iters"];
    job.skipOpenDist = ctx->opt["skip_open_dist"];
    job.zeroCheckMin = ctx->opt["zero_check_min"];
    job.specPercent = ctx->opt["spec_percent"];
    DS_CUDA(ctx, cudaMemsetAsync(ctx->queue, 0, sizeof(unsigned long long), ctx->stream));
    const LaunchConfig cfg = launchConfig(ctx);
    if (ctx->opt["precision"] == DS_PRECISION_FAST && cfg.variant == 0 && ctx->opt["fused_volume"] != 0 && ctx->baked) {
        const int rc = ensureFusedTexture(ctx);
        if (rc) return rc;
        sc.fusedTex = ctx->fusedTex;
    }
    const bool prof = ctx->opt["profile_events"] != 0;
    if (prof) {
        while (ctx->traceEvents.size() < ctx->traceEventsUsed + 2) {
            cudaEvent_t e;
            DS_CUDA(ctx, cudaEventCreate(&e));
            ctx->traceEvents.push_back(e);
        }
        DS_CUDA(ctx, cudaEventRecord(ctx->traceEvents[ctx->traceEventsUsed], ctx->stream));
    }
    if (ctx->opt["precision"] == DS_PRECISION_FAST)
        DS_CUDA(ctx, KernelSet<true>::trace(sc, job, cfg, ctx->stream));
    else
        DS_CUDA(ctx, KernelSet<false>::trace(sc, job, cfg, ctx->stream));
    ctx->launches++;
    if (prof) {
        DS_CUDA(ctx, cudaEventRecord(ctx->traceEvents[ctx->traceEventsUsed + 1], ctx->stream));
        ctx->traceEventsUsed += 2;
    }
    return DS_OK;
}

static int requireScene(DsContext* ctx, bool needBake)
{
    if (ctx->levels.empty()) DS_FAIL(ctx, DS_ERR_STATE, "no volume uploaded");
    if (!ctx->sceneSet) DS_FAIL(ctx, DS_ERR_STATE, "ds_scene_set has not been called");
    if (needBake && !ctx->baked) DS_FAIL(ctx, DS_ERR_STATE, "sun transmittance volume not baked (ds_bake_sun_transmittance)");
    return DS_OK;
}

extern "C" {

/* ================================================================ context */

int ds_context_create(int device, DsContext** out)
{
    if (!out) return DS_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_createError = std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        return DS_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_createError = "device index out of range";
        return DS_ERR_INVALID;
    }
    DsContext* ctx = new DsContext();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess) {
        g_createError = "cudaSetDevice / cudaGetDeviceProperties failed";
        delete ctx;
        return DS_ERR_CUDA;
    }
    if (ctx->prop.major < 10) {
        char b[256];
        snprintf(b, sizeof(b), "device %d (%s, sm_%d%d) is not a Blackwell sm_100 GPU; kernels are built for sm_100a only", device,
                 ctx->prop.name, ctx->prop.major, ctx->prop.minor);
        g_createError = b;
        delete ctx;
        return DS_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_createError = "cudaStreamCreate failed";
        delete ctx;
        return DS_ERR_CUDA;
    }
    ctx->opt["precision"] = DS_PRECISION_FAST;
    ctx->opt["variant"] = 0;
    ctx->opt["block_threads"] = 1024;
    ctx->opt["blocks_per_sm"] = 1;
    ctx->opt["skip_empty"] = 1;
    ctx->opt["march_keep_quarters"] = 2;
    ctx->opt["march_max_iters"] = 64;
    ctx->opt["march_keep32"] = 11;
    ctx->opt["regen_min"] = 4;
    ctx->opt["skip_min"] = 10;
    ctx->opt["skip_max_iters"] = 32;
    ctx->opt["skip_open_dist"] = 1;
    ctx->opt["zero_check_min"] = 8;
    ctx->opt["radiance_scheduler"] = 1;
    ctx->opt["smem_carveout"] = -1;
    ctx->opt["volume_generation"] = 0;
    ctx->opt["radiance_quota"] = 256;
    ctx->opt["march_unroll"] = 0; /* auto */
    ctx->opt["staging_subframes"] = 16;
    ctx->opt["stream_offset"] = 0;
    ctx->opt["profile_events"] = 0;
    ctx->opt["primary_cache"] = 1;
    ctx->opt["spec_percent"] = 100; /* FAST render, two-tap pipeline: threshold of the speculative second tap (0 = always fetch it) */
    ctx->opt["escape_octants"] = 1; /* FAST estimator: a path in an empty cell whose whole octant ahead is empty ends without walking the leap DDA */
    ctx->opt["fused_volume"] = 1; /* FAST estimator: march through one RG8 {density, sun transmittance} array instead of two R8 arrays */
    ctx->opt["region_pixels"] = -1; /* FAST render: hit-list pixels per region of the region-major item order (0 = subframe-major; -1 = auto =
                                       1024: against 4096, C4 730 -> 757 and C2 1408 -> 1421 Mpaths/s, profiles/r04a_*, r04c_*) */
    ctx->opt["descriptor_hw"] = -1;
    ctx->opt["mlp_fp16"] = 1; /* FAST flavour of the model on IEEE half operands (default): the MMA rate and operand bytes of bf16 with the 10 mantissa
                                 bits of tf32 (656 vs 357 TFLOP/s, max error against the fp32 model 1.2e-3 either way); 0 = tf32 operands.  Activations
                                 are converted with .satfinite: beyond 65504 they would clamp, where tf32 would carry on */
    ctx->opt["mlp_bf16"] = 0; /* FAST flavour of the model: 0 = tf32 operands (default), 1 = bf16 operands (twice the MMA rate, half the operand bytes) */
    ctx->opt["compact_reverse"] = 0; /* test hook: neural renderer processes the scattering pixels in the opposite order */
    ctx->opt["mlp_last_us"] = 0; /* read-only: device time of the last model launch when profile_events is on */
    ctx->opt["descriptors_last_us"] = 0; /* read-only: device time of the last descriptor-gather kernel (profile_events) */
    ctx->opt["bake_last_us"] = 0;        /* read-only: device time of the last sun-transmittance bake kernel (profile_events) */
    ds_scene_params_default(&ctx->params);
    bool ok = cudaMalloc(&ctx->stats, CNT_COUNT * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&ctx->queue, sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemset(ctx->stats, 0, CNT_COUNT * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&ctx->mie, 3 * MIE_N * sizeof(float)) == cudaSuccess;
    if (ok) {
        if ((size_t)(ds_mie_blob_end - ds_mie_blob) != 2 * MIE_N * sizeof(float)) {
            g_createError = "embedded Mie table blob has the wrong size";
            ok = false;
        } else {
            std::vector<float> raw(2 * MIE_N), samplers(3 * MIE_N);
            memcpy(raw.data(), ds_mie_blob, raw.size() * sizeof(float));
            buildMieSamplers(raw.data(), raw.data() + MIE_N, samplers.data());
            ok = cudaMemcpy(ctx->mie, samplers.data(), samplers.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
            /* two-level guide of the CDF inversion (DevScene::guideA / guideB): entry = first index i with cdf[i] >= bucket start; followed by
             * the chopped phase sampler as IEEE halves (DevScene::choppedHalf) */
            std::vector<uint16_t> guide(GUIDE_A_N + GUIDE_B_N + MIE_N);
            const float* cdf = samplers.data() + 2 * MIE_N;
            auto fill = [&](uint16_t* dst, int buckets, float limit) {
                int idx = 0, maxKnots = 0;
                std::vector<int> first(buckets + 1);
                for (int k = 0; k <= buckets; k++) {
                    const float v = (float)k / (float)buckets * limit; /* exact: powers of two */
                    while (idx < MIE_N && cdf[idx] < v) idx++;
                    first[k] = idx;
                }
                for (int k = 0; k < buckets; k++) {
                    /* buckets of guideA below the limit are never looked up (guideB serves them) */
                    const bool used = limit < 1.0f || (float)(k + 1) / (float)buckets > GUIDE_B_LIMIT;
                    if (used) maxKnots = std::max(maxKnots, first[k + 1] - first[k]);
                    dst[k] = (uint16_t)std::min(first[k], MIE_N - 1);
                }
                return maxKnots;
            };
            const int mA = fill(guide.data(), GUIDE_A_N, 1.0f), mB = fill(guide.data() + GUIDE_A_N, GUIDE_B_N, GUIDE_B_LIMIT);
            if (mA > GUIDE_MAX_KNOTS || mB > GUIDE_MAX_KNOTS || !(cdf[MIE_N - 1] >= 1.0f)) {
                g_createError = "chopped-Mie CDF does not fit the two-level guide (more than 8 knots in a bucket)";
                ok = false;
            }
            for (int i = 0; i < MIE_N; i++) guide[GUIDE_A_N + GUIDE_B_N + i] = __half_as_ushort(__float2half_rn(samplers[MIE_N + i]));
            ok = ok && cudaMalloc(&ctx->guide, guide.size() * sizeof(uint16_t)) == cudaSuccess &&
                 cudaMemcpy(ctx->guide, guide.data(), guide.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) == cudaSuccess;
        }
    }
    if (!ok) {
        if (g_createError.empty()) g_createError = "device allocation failed";
        ds_context_destroy(ctx);
        return DS_ERR_CUDA;
    }
    *out = ctx;
    return DS_OK;
}

int ds_context_destroy(DsContext* ctx)
{
    ds_cloud_forget(ctx); /* the importer's cache entry dies with the context */
    if (!ctx) return DS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ds_comm_destroy(ctx);
    cudaFree(ctx->moments);
    freeVolume(ctx);
    freeFrame(ctx);
    cudaFree(ctx->stats);
    cudaFree(ctx->queue);
    cudaFree(ctx->mie);
    cudaFree(ctx->guide);
    for (int i = 0; i < 8; i++) cudaFree(ctx->scratch[i]);
    for (int i = 0; i < 7; i++) cudaFree(ctx->mlpScratch[i]);
    freeDisneyModel(ctx);
    for (cudaEvent_t e : ctx->traceEvents) cudaEventDestroy(e);
    if (ctx->copyStream) {
        cudaStreamDestroy(ctx->copyStream);
        cudaEventDestroy(ctx->copyFence);
        cudaEventDestroy(ctx->copyDone);
    }
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DS_OK;
}

const char* ds_last_error(DsContext* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int ds_context_set_stream(DsContext* ctx, void* cuda_stream)
{
    DS_CHECK_CTX(ctx);
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->ownStream = false;
    return DS_OK;
}

int ds_sync(DsContext* ctx)
{
    DS_CHECK_CTX(ctx);
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

int ds_set_option(DsContext* ctx, const char* name, int value)
{
    DS_CHECK_CTX(ctx);
    if (!name || ctx->opt.find(name) == ctx->opt.end()) DS_FAIL(ctx, DS_ERR_INVALID, "unknown option '%s'", name ? name : "(null)");
    const std::string n = name;
    if (n == "block_threads" && (value < 32 || value > 1024 || value % 32)) DS_FAIL(ctx, DS_ERR_INVALID, "block_threads must be 32..1024, multiple of 32");
    if (n == "blocks_per_sm" && (value < 1 || value > 32)) DS_FAIL(ctx, DS_ERR_INVALID, "blocks_per_sm must be 1..32");
    if (n == "precision" && value != DS_PRECISION_EXACT && value != DS_PRECISION_FAST) DS_FAIL(ctx, DS_ERR_INVALID, "precision must be 0 or 1");
    if (n == "march_unroll" && (value < 0 || value > 2)) DS_FAIL(ctx, DS_ERR_INVALID, "march_unroll must be 0 (auto), 1 or 2");
    if (n == "skip_open_dist" && value < 1) DS_FAIL(ctx, DS_ERR_INVALID, "skip_open_dist must be >= 1 (0 would leap out of occupied cells)");
    if (n == "staging_subframes" && value < 1) DS_FAIL(ctx, DS_ERR_INVALID, "staging_subframes must be >= 1");
    if (n == "region_pixels" && (value < -1 || value > (1 << 20))) DS_FAIL(ctx, DS_ERR_INVALID, "region_pixels must be -1 (auto) or 0 .. 2^20");
    if (n == "spec_percent" && (value < 0 || value > 1000)) DS_FAIL(ctx, DS_ERR_INVALID, "spec_percent must be 0 .. 1000");
    if (n == "march_max_iters" && value < 1) DS_FAIL(ctx, DS_ERR_INVALID, "march_max_iters must be >= 1");
    ctx->opt[n] = value;
    return DS_OK;
}

int ds_get_option(DsContext* ctx, const char* name, int* value)
{
    DS_CHECK_CTX(ctx);
    if (!name || !value || ctx->opt.find(name) == ctx->opt.end()) DS_FAIL(ctx, DS_ERR_INVALID, "unknown option '%s'", name ? name : "(null)");
    *value = ctx->opt[name];
    return DS_OK;
}

int ds_get_counters(DsContext* ctx, DsCounters* out)
{
    DS_CHECK_CTX(ctx);
    if (!out) return DS_ERR_INVALID;
    unsigned long long h[CNT_COUNT];
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DS_CUDA(ctx, cudaMemcpy(h, ctx->stats, sizeof(h), cudaMemcpyDeviceToHost));
    out->paths = h[CNT_PATHS] + ctx->extraPaths;
    out->events = h[CNT_EVENTS];
    out->steps = h[CNT_STEPS] + ctx->extraSteps;
    out->density_taps = h[CNT_TAPS];
    out->nonfinite = h[CNT_NONFINITE];
    out->untraced_paths = ctx->extraPaths;
    out->untraced_steps = ctx->extraSteps;
    return DS_OK;
}

int ds_reset_counters(DsContext* ctx)
{
    DS_CHECK_CTX(ctx);
    DS_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, CNT_COUNT * sizeof(unsigned long long), ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches = 0;
    ctx->traceEventsUsed = 0;
    ctx->extraPaths = ctx->extraSteps = 0;
    return DS_OK;
}

int ds_get_launch_stats(DsContext* ctx, uint64_t* kernel_launches, uint64_t* trace_launches_timed, double* trace_ms_total)
{
    DS_CHECK_CTX(ctx);
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double total = 0;
    for (size_t i = 0; i + 1 < ctx->traceEventsUsed; i += 2) {
        float ms = 0;
        DS_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->traceEvents[i], ctx->traceEvents[i + 1]));
        total += ms;
    }
    if (kernel_launches) *kernel_launches = ctx->launches;
    if (trace_launches_timed) *trace_launches_timed = ctx->traceEventsUsed / 2;
    if (trace_ms_total) *trace_ms_total = total;
    return DS_OK;
}

const char* ds_describe(DsContext* ctx)
{
    if (!ctx) return "{}";
    char b[512];
    snprintf(b, sizeof(b),
             "{\"library\": \"deepestscatter_b200\", \"device\": \"%s\", \"sm\": %d%d, \"sm_count\": %d, \"l2_bytes\": %d, "
             "\"global_mem_bytes\": %zu, \"arch\": \"sm_100a\"}",
             ctx->prop.name, ctx->prop.major, ctx->prop.minor, ctx->prop.multiProcessorCount, ctx->prop.l2CacheSize,
             (size_t)ctx->prop.totalGlobalMem);
    ctx->desc = b;
    return ctx->desc.c_str();
}

/* ================================================================ volume */

int ds_volume_upload(DsContext* ctx, const uint8_t* level0, int nx, int ny, int nz, int build_mips)
{
    DS_CHECK_CTX(ctx);
    if (!level0) DS_FAIL(ctx, DS_ERR_INVALID, "level0 is NULL");
    int rc = beginVolume(ctx, nx, ny, nz);
    if (rc) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->levels[0], level0, (size_t)nx * ny * nz, cudaMemcpyHostToDevice, ctx->stream));
    return finishVolume(ctx, build_mips);
}

int ds_volume_upload_float(DsContext* ctx, const float* dense, int nx, int ny, int nz, double max_density, int build_mips)
{
    DS_CHECK_CTX(ctx);
    if (!dense) DS_FAIL(ctx, DS_ERR_INVALID, "dense is NULL");
    if (!(max_density > 0)) DS_FAIL(ctx, DS_ERR_INVALID, "max_density must be positive");
    int rc = beginVolume(ctx, nx, ny, nz);
    if (rc) return rc;
    const size_t count = (size_t)nx * ny * nz;
    rc = ensureScratch(ctx, 0, count * sizeof(float));
    if (rc) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[0], dense, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    DS_CUDA(ctx, launchQuantize((const float*)ctx->scratch[0], count, max_density, ctx->levels[0], ctx->stream));
    return finishVolume(ctx, build_mips);
}

int ds_volume_synth(DsContext* ctx, int n, int kind, uint32_t seed, int build_mips)
{
    DS_CHECK_CTX(ctx);
    if (n < 4 || n > 2048) DS_FAIL(ctx, DS_ERR_INVALID, "synthetic grid size %d out of range [4, 2048]", n);
    if (kind < 0 || kind > 2) DS_FAIL(ctx, DS_ERR_INVALID, "unknown synthetic kind %d", kind);
    int rc = beginVolume(ctx, n, n, n);
    if (rc) return rc;
    DS_CUDA(ctx, launchSynth(ctx->levels[0], n, kind, seed, ctx->stream));
    return finishVolume(ctx, build_mips);
}

int ds_volume_level_count(DsContext* ctx, int* count)
{
    DS_CHECK_CTX(ctx);
    if (!count) return DS_ERR_INVALID;
    *count = (int)ctx->levels.size();
    return DS_OK;
}

int ds_volume_level_dims(DsContext* ctx, int level, int dims[3])
{
    DS_CHECK_CTX(ctx);
    if (level < 0 || level >= (int)ctx->levels.size()) DS_FAIL(ctx, DS_ERR_INVALID, "mip level %d out of range", level);
    dims[0] = ctx->lnx[level];
    dims[1] = ctx->lny[level];
    dims[2] = ctx->lnz[level];
    return DS_OK;
}

int ds_volume_download_level(DsContext* ctx, int level, uint8_t* out)
{
    DS_CHECK_CTX(ctx);
    if (level < 0 || level >= (int)ctx->levels.size() || !out) DS_FAIL(ctx, DS_ERR_INVALID, "mip level %d out of range", level);
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DS_CUDA(ctx, cudaMemcpy(out, ctx->levels[level], (size_t)ctx->lnx[level] * ctx->lny[level] * ctx->lnz[level], cudaMemcpyDeviceToHost));
    return DS_OK;
}

/* ================================================================ scene */

void ds_scene_params_default(DsSceneParams* p)
{
    if (!p) return;
    p->cloud_size_m = 7000.0f;      /* main.cpp:63 */
    p->mean_free_path_m = 10.0f;    /* SceneDescription.h:80 */
    p->sample_step = 1.0f / 512.f;  /* installers.cpp:86 */
    p->light_direction[0] = -0.03f; /* LightDirection::Side, Tasks.cpp:58 */
    p->light_direction[1] = -0.25f;
    p->light_direction[2] = 0.8f;
    p->light_color[0] = p->light_color[1] = p->light_color[2] = 1.0f; /* installers.cpp:99 */
    p->light_intensity = 1e6f;                                           /* installers.cpp:100 */
    p->minimal_ray_distance = 0.000001f;                                 /* CloudMaterial.cpp:23 */
}

int ds_scene_set(DsContext* ctx, const DsSceneParams* p)
{
    DS_CHECK_CTX(ctx);
    if (!p) DS_FAIL(ctx, DS_ERR_INVALID, "params is NULL");
    if (!(p->cloud_size_m > 0) || !(p->mean_free_path_m > 0) || !(p->sample_step > 0) || p->sample_step > 1)
        DS_FAIL(ctx, DS_ERR_INVALID, "cloud_size_m, mean_free_path_m must be > 0 and sample_step in (0, 1]");
    const float l2 = p->light_direction[0] * p->light_direction[0] + p->light_direction[1] * p->light_direction[1] +
                     p->light_direction[2] * p->light_direction[2];
    if (!(l2 > 0)) DS_FAIL(ctx, DS_ERR_INVALID, "light_direction is zero");
    const bool lightChanged = !ctx->sceneSet || memcmp(p->light_direction, ctx->params.light_direction, 12) != 0 ||
                              p->cloud_size_m != ctx->params.cloud_size_m || p->mean_free_path_m != ctx->params.mean_free_path_m ||
                              p->sample_step != ctx->params.sample_step;
    ctx->params = *p;
    ctx->sceneSet = true;
    ctx->primaryValid = false; /* entry steps are counted in units of sample_step */
    ctx->disneyPrimaryValid = false;
    if (lightChanged) ctx->baked = false;
    computeDerived(ctx);
    return DS_OK;
}

int ds_scene_get_derived(DsContext* ctx, float out[12])
{
    DS_CHECK_CTX(ctx);
    memcpy(out, ctx->derived, sizeof(ctx->derived));
    return DS_OK;
}

int ds_bake_sun_transmittance(DsContext* ctx)
{
    DS_CHECK_CTX(ctx);
    int rc = requireScene(ctx, false);
    if (rc) return rc;
    DevScene sc;
    fillDevScene(ctx, sc);
    KernelTimer timer(ctx, "bake_last_us");
    if (ctx->opt["precision"] == DS_PRECISION_FAST)
        DS_CUDA(ctx, KernelSet<true>::bake(sc, ctx->inscatter, ctx->opt["skip_empty"], ctx->stream));
    else
        DS_CUDA(ctx, KernelSet<false>::bake(sc, ctx->inscatter, ctx->opt["skip_empty"], ctx->stream));
    timer.stop();
    rc = makeTexture(ctx, ctx->inscatter, ctx->lnx[0], ctx->lny[0], ctx->lnz[0], &ctx->inscatterArr, &ctx->inscatterTex);
    if (rc) return rc;
    ctx->fusedValid = false;
    ctx->baked = true;
    return DS_OK;
}

int ds_invert_phase_cdf(DsContext* ctx, const float* values, uint32_t n, float* cos_theta_out, float* phase_out)
{
    DS_CHECK_CTX(ctx);
    if (!values || !cos_theta_out || !phase_out) return DS_ERR_INVALID;
    if (n == 0) return DS_OK;
    DevScene sc;
    fillDevScene(ctx, sc);
    float* d = nullptr;
    DS_CUDA(ctx, cudaMalloc(&d, (size_t)n * 3 * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(d, values, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launchInvertCdf(sc, d, n, d + n, d + 2 * (size_t)n, ctx->stream);
    ctx->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(cos_theta_out, d + n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(phase_out, d + 2 * (size_t)n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    DS_CUDA(ctx, e);
    return DS_OK;
}

int ds_inscatter_download(DsContext* ctx, uint8_t* out)
{
    DS_CHECK_CTX(ctx);
    if (ctx->levels.empty() || !out) DS_FAIL(ctx, DS_ERR_STATE, "no volume");
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DS_CUDA(ctx, cudaMemcpy(out, ctx->inscatter, (size_t)ctx->lnx[0] * ctx->lny[0] * ctx->lnz[0], cudaMemcpyDeviceToHost));
    return DS_OK;
}

int ds_inscatter_upload(DsContext* ctx, const uint8_t* in)
{
    DS_CHECK_CTX(ctx);
    if (ctx->levels.empty() || !in) DS_FAIL(ctx, DS_ERR_STATE, "no volume");
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->inscatter, in, (size_t)ctx->lnx[0] * ctx->lny[0] * ctx->lnz[0], cudaMemcpyHostToDevice, ctx->stream));
    int rc = makeTexture(ctx, ctx->inscatter, ctx->lnx[0], ctx->lny[0], ctx->lnz[0], &ctx->inscatterArr, &ctx->inscatterTex);
    if (rc) return rc;
    ctx->fusedValid = false;
    ctx->baked = true;
    return DS_OK;
}

void ds_camera_look_at(const float eye[3], const float lookat[3], const float up[3], float hfov_deg, float aspect, DsCamera* out)
{
    /* sutil::calculateCameraVariables, fov_is_vertical = false */
    float W[3] = {lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]};
    const float wlen = sqrtf(W[0] * W[0] + W[1] * W[1] + W[2] * W[2]);
    auto cross3 = [](const float* a, const float* b, float* c) {
        c[0] = a[1] * b[2] - a[2] * b[1];
        c[1] = a[2] * b[0] - a[0] * b[2];
        c[2] = a[0] * b[1] - a[1] * b[0];
    };
    auto norm3 = [](float* v) {
        const float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        v[0] *= inv;
        v[1] *= inv;
        v[2] *= inv;
    };
    float U[3], V[3];
    cross3(W, up, U);
    norm3(U);
    cross3(U, W, V);
    norm3(V);
    const float ulen = wlen * tanf(0.5f * hfov_deg * PI_F / 180.0f);
    const float vlen = ulen / aspect;
    for (int i = 0; i < 3; i++) {
        out->eye[i] = eye[i];
        out->U[i] = U[i] * ulen;
        out->V[i] = V[i] * vlen;
        out->W[i] = W[i];
    }
}

void ds_camera_default(int width, int height, DsCamera* out)
{
    const float eye[3] = {2.5f, -0.4f, 0.0f}, lookat[3] = {0, 0, 0}, up[3] = {0, 1, 0};
    ds_camera_look_at(eye, lookat, up, 30.0f, (float)width / (float)height, out);
}

/* ================================================================ progressive renderer */

int ds_frame_create(DsContext* ctx, int width, int height)
{
    DS_CHECK_CTX(ctx);
    if (width <= 0 || height <= 0 || width > 4096 || height > 4096)
        DS_FAIL(ctx, DS_ERR_INVALID, "frame size %dx%d out of range (seed packs x*4096+y, cloudRadianceMaterials.cu:21)", width, height);
    freeFrame(ctx);
    const size_t px = (size_t)width * height;
    DS_CUDA(ctx, cudaMalloc(&ctx->progressive, px * sizeof(float4)));
    DS_CUDA(ctx, cudaMalloc(&ctx->variance, px * sizeof(float4)));
    DS_CUDA(ctx, cudaMalloc(&ctx->screen, px * sizeof(uchar4)));
    DS_CUDA(ctx, cudaMalloc(&ctx->columns, (size_t)width * sizeof(float)));
    DS_CUDA(ctx, cudaMalloc(&ctx->average, sizeof(float)));
    DS_CUDA(ctx, cudaMalloc(&ctx->unconv, sizeof(uint32_t)));
    DS_CUDA(ctx, cudaMalloc(&ctx->entrySteps, px * sizeof(uint32_t)));
    DS_CUDA(ctx, cudaMalloc(&ctx->hitList, px * sizeof(uint32_t)));
    DS_CUDA(ctx, cudaMalloc(&ctx->primaryCounts, 2 * sizeof(unsigned long long)));
    ctx->width = width;
    ctx->height = height;
    return ds_frame_clear(ctx);
}

int ds_frame_clear(DsContext* ctx)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    const size_t px = (size_t)ctx->width * ctx->height;
    DS_CUDA(ctx, cudaMemsetAsync(ctx->progressive, 0, px * sizeof(float4), ctx->stream));
    DS_CUDA(ctx, cudaMemsetAsync(ctx->variance, 0, px * sizeof(float4), ctx->stream));
    return DS_OK;
}

static int ensureStaging(DsContext* ctx, size_t subframes)
{
    if (ctx->stagingSubframes >= subframes) return DS_OK;
    if (subframes * (size_t)ctx->width * ctx->height >= (1ull << 32))
        DS_FAIL(ctx, DS_ERR_INVALID, "staging_subframes x width x height must stay below 2^32 samples");
    cudaFree(ctx->staging);
    ctx->staging = nullptr;
    ctx->stagingSubframes = 0;
    DS_CUDA(ctx, cudaMalloc(&ctx->staging, subframes * (size_t)ctx->width * ctx->height * sizeof(float4)));
    ctx->stagingSubframes = subframes;
    return DS_OK;
}

static void fillRenderJob(DsContext* ctx, TraceJob& job, const DsCamera* cam, DsMode mode, uint32_t first, uint32_t n)
{
    memset(&job, 0, sizeof(job));
    job.kind = JOB_RENDER;
    job.mode = (int)mode;
    memcpy(job.eye, cam->eye, 12);
    memcpy(job.U, cam->U, 12);
    memcpy(job.V, cam->V, 12);
    memcpy(job.W, cam->W, 12);
    job.width = ctx->width;
    job.height = ctx->height;
    job.tilesX = (ctx->width + 7) / 8;
    const int tilesY = (ctx->height + 3) / 4;
    job.itemsPerSubframe = (unsigned long long)job.tilesX * tilesY * 32ull;
    /* RNG stream id of a subframe; stream_offset lets a rank render global subframes [offset+1, offset+n]
     * while its local Welford weights still run 1/1, 1/2, ... (multi-GPU split) */
    job.firstSubframe = first + (uint32_t)ctx->opt["stream_offset"];
    job.total = job.itemsPerSubframe * n;
    job.staging = ctx->staging;
}

/* The primary-ray cache applies to the optimised FAST kernel for estimators that follow the camera ray
 * (all-order and single scatter; the multiple-scatter estimator resamples the direction first). */
static bool usePrimaryCache(DsContext* ctx, DsMode mode)
{
    return ctx->opt["precision"] == DS_PRECISION_FAST && ctx->opt["variant"] == 0 && ctx->opt["skip_empty"] != 0 &&
           ctx->opt["primary_cache"] != 0 && mode != DS_MODE_SUN_MULTIPLE_SCATTER;
}

static int ensurePrimary(DsContext* ctx, const DsCamera* cam)
{
    if (ctx->primaryValid && memcmp(&ctx->primaryCam, cam, sizeof(DsCamera)) == 0) return DS_OK;
    DevScene sc;
    fillDevScene(ctx, sc);
    TraceJob job;
    fillRenderJob(ctx, job, cam, DS_MODE_SUN_AND_SKY_ALL_SCATTER, 1, 1);
    DS_CUDA(ctx, cudaMemsetAsync(ctx->primaryCounts, 0, 2 * sizeof(unsigned long long), ctx->stream));
    DS_CUDA(ctx, KernelSet<true>::primaryPrepass(sc, job, ctx->entrySteps, ctx->hitList, ctx->primaryCounts, ctx->stream));
    ctx->launches++;
    unsigned long long h[2];
    DS_CUDA(ctx, cudaMemcpyAsync(h, ctx->primaryCounts, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->nHit = (uint32_t)h[0];
    ctx->missSteps = h[1];
    ctx->primaryCam = *cam;
    ctx->primaryValid = true;
    return DS_OK;
}

/* trace `n` subframes starting at `first` into the staging buffer */
static int traceSubframes(DsContext* ctx, const DsCamera* cam, DsMode mode, uint32_t first, uint32_t n, bool* cached)
{
    TraceJob job;
    fillRenderJob(ctx, job, cam, mode, first, n);
    *cached = usePrimaryCache(ctx, mode);
    if (*cached) {
        int rc = ensurePrimary(ctx, cam);
        if (rc) return rc;
        job.hitList = ctx->hitList;
        job.nHit = ctx->nHit;
        job.entrySteps = ctx->entrySteps;
        job.total = (unsigned long long)ctx->nHit * n;
        int regionPixels = ctx->opt["region_pixels"];
        if (regionPixels < 0) regionPixels = 1024;
        job.regionSize = (uint32_t)regionPixels;
        job.nSub = n;
        ctx->extraPaths += ((unsigned long long)ctx->width * ctx->height - ctx->nHit) * n;
        ctx->extraSteps += ctx->missSteps * n;
        if (ctx->nHit == 0) return DS_OK;
    }
    return runTrace(ctx, job);
}

static int checkRender(DsContext* ctx, const DsCamera* cam, DsMode mode)
{
    int rc = requireScene(ctx, true);
    if (rc) return rc;
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    if (!cam) DS_FAIL(ctx, DS_ERR_INVALID, "camera is NULL");
    if ((int)mode < 0 || (int)mode > 2) DS_FAIL(ctx, DS_ERR_INVALID, "Invalid Render Mode"); /* CloudMaterial.cpp:62 */
    return DS_OK;
}

int ds_render_frame_result(DsContext* ctx, const DsCamera* cam, DsMode mode, uint32_t subframe_id, float* frame_result_out)
{
    DS_CHECK_CTX(ctx);
    int rc = checkRender(ctx, cam, mode);
    if (rc) return rc;
    rc = ensureStaging(ctx, 1);
    if (rc) return rc;
    bool cached;
    rc = traceSubframes(ctx, cam, mode, subframe_id, 1, &cached);
    if (rc) return rc;
    if (cached) {
        DS_CUDA(ctx, launchFillMissing(ctx->staging, ctx->entrySteps, (size_t)ctx->width * ctx->height, ctx->stream));
        ctx->launches++;
    }
    if (frame_result_out) {
        DS_CUDA(ctx, cudaMemcpyAsync(frame_result_out, ctx->staging, (size_t)ctx->width * ctx->height * sizeof(float4), cudaMemcpyDeviceToHost,
                                     ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return DS_OK;
}

int ds_render_subframes(DsContext* ctx, const DsCamera* cam, DsMode mode, uint32_t first_subframe, uint32_t n)
{
    DS_CHECK_CTX_DEFER(ctx);
    int rc = checkRender(ctx, cam, mode);
    if (rc) return rc;
    if (first_subframe == 0) DS_FAIL(ctx, DS_ERR_INVALID, "subframe ids are 1-based (Camera.cpp:191)");
    const uint32_t chunkMax = (uint32_t)ctx->opt["staging_subframes"];
    rc = ensureStaging(ctx, std::min<uint32_t>(chunkMax, n));
    if (rc) return rc;
    const size_t px = (size_t)ctx->width * ctx->height;
    for (uint32_t done = 0; done < n;) {
        const uint32_t chunk = std::min<uint32_t>(chunkMax, n - done);
        bool cached;
        rc = traceSubframes(ctx, cam, mode, first_subframe + done, chunk, &cached);
        if (rc) {
            settleUpload(ctx);
            return rc;
        }
        settleUpload(ctx); /* a frame upload (ds_frame_upload) ran beside the trace kernel; the accumulation needs it now */
        DS_CUDA(ctx, launchUpdateFrame(ctx->staging, cached ? ctx->entrySteps : nullptr, ctx->progressive, ctx->variance, px,
                                       first_subframe + done, chunk, ctx->stream));
        ctx->launches++;
        done += chunk;
    }
    settleUpload(ctx);
    return DS_OK;
}

int ds_render_subframes_host(DsContext* ctx, const DsCamera* cam, DsMode mode, uint32_t first_subframe, uint32_t n, float* progressive_inout,
                             float* variance_inout)
{
    DS_CHECK_CTX(ctx);
    if (!progressive_inout || !variance_inout) DS_FAIL(ctx, DS_ERR_INVALID, "host buffers are NULL");
    int rc = ds_frame_upload(ctx, progressive_inout, variance_inout); /* on the copy stream: overlaps the trace kernel */
    if (rc) return rc;
    rc = ds_render_subframes(ctx, cam, mode, first_subframe, n);
    if (rc) return rc;
    return ds_frame_download(ctx, progressive_inout, variance_inout);
}

int ds_frame_download(DsContext* ctx, float* progressive_out, float* variance_out)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    const size_t bytes = (size_t)ctx->width * ctx->height * sizeof(float4);
    if (progressive_out) DS_CUDA(ctx, cudaMemcpyAsync(progressive_out, ctx->progressive, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (variance_out) DS_CUDA(ctx, cudaMemcpyAsync(variance_out, ctx->variance, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

int ds_frame_upload(DsContext* ctx, const float* progressive, const float* variance)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    const size_t bytes = (size_t)ctx->width * ctx->height * sizeof(float4);
    /* host -> device on a second stream, behind everything queued so far; the next entry point waits for it (settleUpload) -- a following
     * ds_render_subframes only before its first accumulation, so the copy runs beside the trace kernel, which only writes the staging buffer */
    if (!ctx->copyStream) {
        DS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        DS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copyFence, cudaEventDisableTiming));
        DS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copyDone, cudaEventDisableTiming));
    }
    DS_CUDA(ctx, cudaEventRecord(ctx->copyFence, ctx->stream));
    DS_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->copyFence, 0));
    if (progressive) DS_CUDA(ctx, cudaMemcpyAsync(ctx->progressive, progressive, bytes, cudaMemcpyHostToDevice, ctx->copyStream));
    if (variance) DS_CUDA(ctx, cudaMemcpyAsync(ctx->variance, variance, bytes, cudaMemcpyHostToDevice, ctx->copyStream));
    DS_CUDA(ctx, cudaEventRecord(ctx->copyDone, ctx->copyStream));
    ctx->uploadPending = true;
    return DS_OK;
}

int ds_frame_device_ptrs(DsContext* ctx, void** progressive, void** variance)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    if (progressive) *progressive = ctx->progressive;
    if (variance) *variance = ctx->variance;
    return DS_OK;
}

int ds_tonemap(DsContext* ctx, float exposure, uint8_t* screen_out, float* average_luminance_out)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    DS_CUDA(ctx, launchTonemap(ctx->progressive, ctx->width, ctx->height, exposure, ctx->columns, ctx->average, ctx->screen, ctx->stream));
    if (screen_out)
        DS_CUDA(ctx, cudaMemcpyAsync(screen_out, ctx->screen, (size_t)ctx->width * ctx->height * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->stream));
    if (average_luminance_out)
        DS_CUDA(ctx, cudaMemcpyAsync(average_luminance_out, ctx->average, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

int ds_frame_unconverged(DsContext* ctx, uint32_t subframe_id, uint32_t* unconverged_out)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive || !unconverged_out) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    DS_CUDA(ctx, cudaMemsetAsync(ctx->unconv, 0, sizeof(uint32_t), ctx->stream));
    DS_CUDA(ctx, launchUnconverged(ctx->progressive, ctx->variance, (size_t)ctx->width * ctx->height, subframe_id, ctx->unconv, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(unconverged_out, ctx->unconv, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

int ds_frame_export_moments_device(DsContext* ctx, uint32_t n, double* moments_device)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive || !moments_device) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    DS_CUDA(ctx, launchExportMoments(ctx->progressive, ctx->variance, (size_t)ctx->width * ctx->height, n, moments_device, ctx->stream));
    return DS_OK;
}

int ds_frame_import_moments_device(DsContext* ctx, uint32_t n_total, const double* moments_device)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive || !moments_device || n_total == 0) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create) or n_total == 0");
    DS_CUDA(ctx, launchImportMoments(moments_device, (size_t)ctx->width * ctx->height, n_total, ctx->progressive, ctx->variance, ctx->stream));
    return DS_OK;
}

/* ================================================================ multi-GPU reduce over NCCL */

/* The handful of NCCL entry points the reduce needs, resolved at run time.  Types follow nccl.h 2.x (stable ABI):
 * ncclUniqueId is 128 opaque bytes, ncclFloat64 = 8, ncclSum = 0, ncclSuccess = 0. */
namespace {
struct NcclUniqueId {
    char internal[DS_COMM_ID_BYTES];
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};
NcclApi& nccl()
{
    static NcclApi api;
    if (api.lib || !api.why.empty()) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
        if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.lib) {
        api.why = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
        return api;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.Reduce = (decltype(api.Reduce))dlsym(api.lib, "ncclReduce");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Reduce || !api.AllReduce) {
        api.why = "NCCL library lacks an entry point (ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclReduce / ncclAllReduce)";
        dlclose(api.lib);
        api.lib = nullptr;
    }
    return api;
}
const int NCCL_FLOAT64 = 8, NCCL_SUM = 0;
} // namespace

#define DS_NCCL(ctx, call)                                                                                                        \
    do {                                                                                                                          \
        int _r = (call);                                                                                                          \
        if (_r != 0) DS_FAIL(ctx, DS_ERR_CUDA, "%s: %s", #call, nccl().GetErrorString ? nccl().GetErrorString(_r) : "NCCL error"); \
    } while (0)

int ds_comm_unique_id(uint8_t id_out[DS_COMM_ID_BYTES])
{
    if (!id_out) return DS_ERR_INVALID;
    NcclApi& api = nccl();
    if (!api.lib) {
        g_createError = api.why;
        return DS_ERR_STATE;
    }
    NcclUniqueId id;
    if (api.GetUniqueId(&id) != 0) {
        g_createError = "ncclGetUniqueId failed";
        return DS_ERR_CUDA;
    }
    memcpy(id_out, id.internal, DS_COMM_ID_BYTES);
    return DS_OK;
}

int ds_comm_init(DsContext* ctx, int n_ranks, int rank, const uint8_t id[DS_COMM_ID_BYTES])
{
    DS_CHECK_CTX(ctx);
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) DS_FAIL(ctx, DS_ERR_INVALID, "bad rank / n_ranks / id");
    NcclApi& api = nccl();
    if (!api.lib) DS_FAIL(ctx, DS_ERR_STATE, "%s", api.why.c_str());
    ds_comm_destroy(ctx);
    NcclUniqueId uid;
    memcpy(uid.internal, id, DS_COMM_ID_BYTES);
    DS_NCCL(ctx, api.CommInitRank(&ctx->comm, n_ranks, uid, rank));
    ctx->commRanks = n_ranks;
    ctx->commRank = rank;
    /* NCCL sets its channels up lazily inside the first collective (hundreds of milliseconds); pay that here with one 8-byte
     * all-reduce so that the first ds_frame_reduce of a render costs what every later one costs */
    int rc = ensureScratch(ctx, 7, 64);
    if (rc) return rc;
    DS_CUDA(ctx, cudaMemsetAsync(ctx->scratch[7], 0, 64, ctx->stream));
    DS_NCCL(ctx, api.AllReduce(ctx->scratch[7], ctx->scratch[7], 1, NCCL_FLOAT64, NCCL_SUM, ctx->comm, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

int ds_comm_destroy(DsContext* ctx)
{
    if (!ctx) return DS_ERR_INVALID;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        nccl().CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->commRanks = 0;
    ctx->commRank = -1;
    return DS_OK;
}

int ds_frame_reduce(DsContext* ctx, uint32_t n_local, uint32_t n_total, int root)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->comm) DS_FAIL(ctx, DS_ERR_STATE, "no communicator (ds_comm_init)");
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    if (n_total == 0 || n_local > n_total || root >= ctx->commRanks) DS_FAIL(ctx, DS_ERR_INVALID, "bad n_local / n_total / root");
    const size_t pixels = (size_t)ctx->width * ctx->height;
    if (ctx->momentsPixels != pixels) {
        cudaFree(ctx->moments);
        ctx->moments = nullptr;
        ctx->momentsPixels = 0;
        DS_CUDA(ctx, cudaMalloc(&ctx->moments, pixels * 8 * sizeof(double)));
        ctx->momentsPixels = pixels;
    }
    /* float64 moments {n*mean, M2 + n*mean^2}: the import subtracts N*mean^2 from the summed second moment, which cancels
     * catastrophically in fp32 wherever a pixel's variance is small against its squared mean; 64 B/pixel over NVLink is
     * 0.15 ms at 1080p, so exactness costs nothing that shows. */
    DS_CUDA(ctx, launchExportMoments(ctx->progressive, ctx->variance, pixels, n_local, ctx->moments, ctx->stream));
    ctx->launches++;
    NcclApi& api = nccl();
    if (root < 0)
        DS_NCCL(ctx, api.AllReduce(ctx->moments, ctx->moments, pixels * 8, NCCL_FLOAT64, NCCL_SUM, ctx->comm, ctx->stream));
    else
        DS_NCCL(ctx, api.Reduce(ctx->moments, ctx->moments, pixels * 8, NCCL_FLOAT64, NCCL_SUM, root, ctx->comm, ctx->stream));
    if (root < 0 || root == ctx->commRank) {
        DS_CUDA(ctx, launchImportMoments(ctx->moments, pixels, n_total, ctx->progressive, ctx->variance, ctx->stream));
        ctx->launches++;
    }
    return DS_OK;
}

/* ================================================================ generic paths */

int ds_trace_paths(DsContext* ctx, DsMode mode, uint32_t n, const float* origins, const float* directions, const uint32_t* seed_val0,
                   const uint32_t* stream, float* radiance_out)
{
    DS_CHECK_CTX(ctx);
    int rc = requireScene(ctx, true);
    if (rc) return rc;
    if ((int)mode < 0 || (int)mode > 2) DS_FAIL(ctx, DS_ERR_INVALID, "Invalid Render Mode");
    if (n == 0) return DS_OK;
    if (!origins || !directions || !seed_val0 || !stream || !radiance_out) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    const size_t f3 = (size_t)n * 3 * sizeof(float), u1 = (size_t)n * sizeof(uint32_t);
    if ((rc = ensureScratch(ctx, 0, f3)) || (rc = ensureScratch(ctx, 1, f3)) || (rc = ensureScratch(ctx, 2, u1)) ||
        (rc = ensureScratch(ctx, 3, u1)) || (rc = ensureScratch(ctx, 4, f3)))
        return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[0], origins, f3, cudaMemcpyHostToDevice, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[1], directions, f3, cudaMemcpyHostToDevice, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[2], seed_val0, u1, cudaMemcpyHostToDevice, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[3], stream, u1, cudaMemcpyHostToDevice, ctx->stream));
    TraceJob job;
    memset(&job, 0, sizeof(job));
    job.kind = JOB_PATHS;
    job.mode = (int)mode;
    job.total = n;
    job.origins = (const float*)ctx->scratch[0];
    job.dirs = (const float*)ctx->scratch[1];
    job.seedVal0 = (const uint32_t*)ctx->scratch[2];
    job.stream = (const uint32_t*)ctx->scratch[3];
    job.radianceOut = (float*)ctx->scratch[4];
    rc = runTrace(ctx, job);
    if (rc) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(radiance_out, ctx->scratch[4], f3, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

/* ================================================================ dataset generation */

int ds_generate_points(DsContext* ctx, uint32_t first_index, uint32_t n, uint32_t stream, float* positions_out, float* directions_out)
{
    DS_CHECK_CTX(ctx);
    int rc = requireScene(ctx, false);
    if (rc) return rc;
    if (n == 0) return DS_OK;
    if (!positions_out || !directions_out) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    const size_t f3 = (size_t)n * 3 * sizeof(float);
    if ((rc = ensureScratch(ctx, 0, f3)) || (rc = ensureScratch(ctx, 1, f3))) return rc;
    DevScene sc;
    fillDevScene(ctx, sc);
    if (ctx->opt["precision"] == DS_PRECISION_FAST)
        DS_CUDA(ctx, KernelSet<true>::generatePoints(sc, first_index, n, stream, (float*)ctx->scratch[0], (float*)ctx->scratch[1], ctx->stats, ctx->stream));
    else
        DS_CUDA(ctx, KernelSet<false>::generatePoints(sc, first_index, n, stream, (float*)ctx->scratch[0], (float*)ctx->scratch[1], ctx->stats, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(positions_out, ctx->scratch[0], f3, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(directions_out, ctx->scratch[1], f3, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

static void fillDescriptorTables(DsContext* ctx, LevelTable& lv, DescriptorLayers& layers)
{
    memset(&lv, 0, sizeof(lv));
    lv.count = (int)ctx->levels.size();
    for (int l = 0; l < lv.count; l++) {
        lv.data[l] = ctx->levels[l];
        lv.nx[l] = ctx->lnx[l];
        lv.ny[l] = ctx->lny[l];
        lv.nz[l] = ctx->lnz[l];
    }
    /* DisneyDescriptor.cuh:81-88,109-110: per-layer scale, mip level and mip voxel size */
    float scale = 0.5f / ctx->derived[6];
    float mipmapLevel = -ds_log2f(ctx->derived[8]) - 1;
    for (int l = 0; l < 10; l++) {
        layers.scale[l] = scale;
        layers.lod[l] = fmaxf(0.0f, mipmapLevel);
        layers.mipVoxelSize[l] = ds_exp2f(mipmapLevel) * ctx->derived[7] / ctx->params.cloud_size_m;
        scale *= 2;
        mipmapLevel++;
    }
}

static int collectDescriptors(DsContext* ctx, const float* positions, const float* directions, uint32_t n, uint8_t* outU8, float* outF32,
                              int32_t* tapIndex)
{
    int rc = requireScene(ctx, false);
    if (rc) return rc;
    if (n == 0) return DS_OK;
    if (!positions || !directions) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    const size_t f3 = (size_t)n * 3 * sizeof(float), taps = (size_t)n * 2250;
    if ((rc = ensureScratch(ctx, 0, f3)) || (rc = ensureScratch(ctx, 1, f3))) return rc;
    if (outU8 && (rc = ensureScratch(ctx, 2, taps))) return rc;
    if (outF32 && (rc = ensureScratch(ctx, 3, taps * sizeof(float)))) return rc;
    if (tapIndex && (rc = ensureScratch(ctx, 4, taps * 4 * sizeof(int32_t)))) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[0], positions, f3, cudaMemcpyHostToDevice, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[1], directions, f3, cudaMemcpyHostToDevice, ctx->stream));
    DevScene sc;
    fillDevScene(ctx, sc);
    LevelTable lv;
    DescriptorLayers layers;
    fillDescriptorTables(ctx, lv, layers);
    cudaTextureObject_t mipTex = 0; /* the collectors use the exact software fetch unless option descriptor_hw = 1 asks for the float one */
    if (outF32 && !outU8 && !tapIndex && (rc = descriptorTexture(ctx, false, &mipTex))) return rc;
    KernelTimer timer(ctx, "descriptors_last_us");
    DS_CUDA(ctx, launchDescriptors(sc, lv, layers, (const float*)ctx->scratch[0], (const float*)ctx->scratch[1], n,
                                   outU8 ? (uint8_t*)ctx->scratch[2] : nullptr, outF32 ? (float*)ctx->scratch[3] : nullptr,
                                   tapIndex ? (int32_t*)ctx->scratch[4] : nullptr, ctx->stream, 225, nullptr, nullptr, nullptr, mipTex));
    timer.stop();
    if (outU8) DS_CUDA(ctx, cudaMemcpyAsync(outU8, ctx->scratch[2], taps, cudaMemcpyDeviceToHost, ctx->stream));
    if (outF32) DS_CUDA(ctx, cudaMemcpyAsync(outF32, ctx->scratch[3], taps * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (tapIndex) DS_CUDA(ctx, cudaMemcpyAsync(tapIndex, ctx->scratch[4], taps * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

int ds_collect_descriptors(DsContext* ctx, const float* positions, const float* directions, uint32_t n, uint8_t* descriptors_out)
{
    DS_CHECK_CTX(ctx);
    if (!descriptors_out && n) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    return collectDescriptors(ctx, positions, directions, n, descriptors_out, nullptr, nullptr);
}

int ds_collect_descriptors_float(DsContext* ctx, const float* positions, const float* directions, uint32_t n, float* out, int32_t* tap_index_out)
{
    DS_CHECK_CTX(ctx);
    if (!out && n) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    return collectDescriptors(ctx, positions, directions, n, nullptr, out, tap_index_out);
}

/* device side of the first launch of renderRect: fills scratch[3] (network input [n][10][226]), scratch[4] (DsIntersectionInfo [n]) and
 * scratch[2] (angle [n] + hasScattered bytes [n]) for the rectangle */
static int networkInputDevice(DsContext* ctx, const DsCamera* cam, uint32_t frame_width, uint32_t frame_height, uint32_t rect_x, uint32_t rect_y,
                              uint32_t rect_w, uint32_t rect_h, uint32_t stream)
{
    int rc;
    const size_t n = (size_t)rect_w * rect_h;
    if ((rc = ensureScratch(ctx, 0, n * 3 * sizeof(float))) || (rc = ensureScratch(ctx, 1, n * 3 * sizeof(float))) ||
        (rc = ensureScratch(ctx, 2, n * (sizeof(float) + 1))) || (rc = ensureScratch(ctx, 3, n * 2260 * sizeof(float))) ||
        (rc = ensureScratch(ctx, 4, n * 5 * sizeof(float))))
        return rc;
    float* dPos = (float*)ctx->scratch[0];
    float* dDir = (float*)ctx->scratch[1];
    float* dAngle = (float*)ctx->scratch[2];
    uint8_t* dActive = (uint8_t*)(dAngle + n);
    float* dInput = (float*)ctx->scratch[3];
    float* dInfo = (float*)ctx->scratch[4];
    TraceJob job;
    memset(&job, 0, sizeof(job));
    memcpy(job.eye, cam->eye, 12);
    memcpy(job.U, cam->U, 12);
    memcpy(job.V, cam->V, 12);
    memcpy(job.W, cam->W, 12);
    job.width = (int)frame_width;
    job.height = (int)frame_height;
    DevScene sc;
    fillDevScene(ctx, sc);
    if (ctx->opt["precision"] == DS_PRECISION_FAST)
        DS_CUDA(ctx, KernelSet<true>::networkInfo(sc, job, (int)rect_x, (int)rect_y, (int)rect_w, (int)rect_h, stream, dInfo, dPos, dDir, dAngle, dActive,
                                                  ctx->stats, ctx->stream));
    else
        DS_CUDA(ctx, KernelSet<false>::networkInfo(sc, job, (int)rect_x, (int)rect_y, (int)rect_w, (int)rect_h, stream, dInfo, dPos, dDir, dAngle, dActive,
                                                   ctx->stats, ctx->stream));
    LevelTable lv;
    DescriptorLayers layers;
    fillDescriptorTables(ctx, lv, layers);
    cudaTextureObject_t mipTex = 0;
    if ((rc = descriptorTexture(ctx, true, &mipTex))) return rc;
    DS_CUDA(ctx, launchDescriptors(sc, lv, layers, dPos, dDir, (uint32_t)n, nullptr, dInput, nullptr, ctx->stream, 226, dAngle, dActive, nullptr, mipTex));
    ctx->launches += 2;
    return DS_OK;
}

int ds_render_network_input(DsContext* ctx, const DsCamera* cam, uint32_t frame_width, uint32_t frame_height, uint32_t rect_x, uint32_t rect_y,
                            uint32_t rect_w, uint32_t rect_h, uint32_t stream, float* network_input_out, DsIntersectionInfo* info_out)
{
    DS_CHECK_CTX(ctx);
    int rc = requireScene(ctx, true);
    if (rc) return rc;
    if (!cam || !network_input_out || !info_out) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    if (frame_width == 0 || frame_height == 0 || rect_w == 0 || rect_h == 0 || (unsigned long long)rect_w * rect_h > (1u << 24))
        DS_FAIL(ctx, DS_ERR_INVALID, "bad frame / rectangle size");
    const size_t n = (size_t)rect_w * rect_h;
    if ((rc = networkInputDevice(ctx, cam, frame_width, frame_height, rect_x, rect_y, rect_w, rect_h, stream))) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(network_input_out, ctx->scratch[3], n * 2260 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaMemcpyAsync(info_out, ctx->scratch[4], n * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DS_OK;
}

/* ---- the radiance-predicting network (DeepestScatter_Train/Disney/DisneyModel.py) ---- */

size_t ds_disney_model_weight_count(void) { return MLP_WEIGHT_COUNT; }

int ds_disney_model_pack(const float* weights, size_t count, int bf16, void* stream_out, size_t stream_capacity, void* chunks_out, size_t chunks_capacity,
                         size_t* stream_bytes, size_t* chunk_count)
{
    if (!weights || count != MLP_WEIGHT_COUNT || !stream_bytes || !chunk_count) return DS_ERR_INVALID;
    DisneyModelHost h;
    packDisneyModel(weights, h);
    const std::vector<uint8_t>& stream = bf16 == 2 ? h.streamF16 : bf16 ? h.streamBf16 : h.stream;
    const std::vector<MlpChunk>& chunks = bf16 ? h.chunksBf16 : h.chunks;
    *stream_bytes = stream.size();
    *chunk_count = chunks.size();
    if (stream_out) {
        if (stream_capacity < stream.size()) return DS_ERR_INVALID;
        memcpy(stream_out, stream.data(), stream.size());
    }
    if (chunks_out) {
        if (chunks_capacity < chunks.size() * sizeof(MlpChunk)) return DS_ERR_INVALID;
        memcpy(chunks_out, chunks.data(), chunks.size() * sizeof(MlpChunk));
    }
    return DS_OK;
}

int ds_disney_model_load(DsContext* ctx, const float* weights, size_t count)
{
    DS_CHECK_CTX(ctx);
    if (!weights) DS_FAIL(ctx, DS_ERR_INVALID, "NULL weights");
    if (count != MLP_WEIGHT_COUNT)
        DS_FAIL(ctx, DS_ERR_INVALID, "DisneyModel state_dict has %zu floats, got %zu", (size_t)MLP_WEIGHT_COUNT, count);
    for (size_t i = 0; i < count; ++i)
        if (!std::isfinite(weights[i])) DS_FAIL(ctx, DS_ERR_INVALID, "weight %zu is not finite", i);
    DisneyModelHost h;
    packDisneyModel(weights, h);
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    freeDisneyModel(ctx);
    DisneyModelDev& m = ctx->model;
    DS_CUDA(ctx, cudaMalloc(&m.wT, h.wT.size() * sizeof(float)));
    DS_CUDA(ctx, cudaMalloc(&m.bias, h.bias.size() * sizeof(float)));
    DS_CUDA(ctx, cudaMalloc(&m.w4b4, h.w4b4.size() * sizeof(float)));
    DS_CUDA(ctx, cudaMalloc(&m.stream, h.stream.size()));
    DS_CUDA(ctx, cudaMalloc(&m.streamBf16, h.streamBf16.size()));
    DS_CUDA(ctx, cudaMalloc(&m.streamF16, h.streamF16.size()));
    DS_CUDA(ctx, cudaMalloc(&m.error, sizeof(uint32_t)));
    DS_CUDA(ctx, cudaMalloc(&m.prof, 16 * sizeof(unsigned long long)));
    DS_CUDA(ctx, cudaMemset(m.prof, 0, 16 * sizeof(unsigned long long)));
    DS_CUDA(ctx, cudaMemcpy(m.wT, h.wT.data(), h.wT.size() * sizeof(float), cudaMemcpyHostToDevice));
    DS_CUDA(ctx, cudaMemcpy(m.bias, h.bias.data(), h.bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    DS_CUDA(ctx, cudaMemcpy(m.w4b4, h.w4b4.data(), h.w4b4.size() * sizeof(float), cudaMemcpyHostToDevice));
    DS_CUDA(ctx, cudaMemcpy(m.stream, h.stream.data(), h.stream.size(), cudaMemcpyHostToDevice));
    DS_CUDA(ctx, cudaMemcpy(m.streamBf16, h.streamBf16.data(), h.streamBf16.size(), cudaMemcpyHostToDevice));
    DS_CUDA(ctx, cudaMemcpy(m.streamF16, h.streamF16.data(), h.streamF16.size(), cudaMemcpyHostToDevice));
    m.program = makeMlpProgram(h.chunks);
    m.programBf16 = makeMlpProgram(h.chunksBf16);
    if (!m.program || !m.programBf16) DS_FAIL(ctx, DS_ERR_INVALID, "model program has %zu chunks", h.chunks.size());
    DS_CUDA(ctx, cudaMemset(m.error, 0, sizeof(uint32_t)));
    m.nChunks = (int)h.chunks.size();
    m.loaded = true;
    return DS_OK;
}

/* operand type of the tensor-core model kernel: 1 = bfloat16 (option mlp_bf16), else 2 = IEEE half (option mlp_fp16, the default), else 0 = tf32 */
static int mlpOperands(DsContext* ctx) { return ctx->opt["mlp_bf16"] ? 1 : ctx->opt["mlp_fp16"] ? 2 : 0; }
static size_t networkTileBytes(DsContext* ctx, size_t rows) { return (rows + 127) / 128 * networkTileBytesOf(mlpOperands(ctx)); }

/* evaluates the loaded model on device rows.  FAST flavour: tcgen05 tf32 kernel, dIn = 128-row tiles (NETWORK_TILE_FLOATS); EXACT flavour: fp32
 * FMA kernel, dIn = DisneyNetworkInput rows [n][10][226] */
static int disneyForwardDevice(DsContext* ctx, const float* dIn, uint32_t nRows, float* dOut)
{
    if (!ctx->model.loaded) DS_FAIL(ctx, DS_ERR_STATE, "no model loaded (ds_disney_model_load)");
    const int prof = ctx->opt["profile_events"];
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) {
        DS_CUDA(ctx, cudaEventCreate(&e0));
        DS_CUDA(ctx, cudaEventCreate(&e1));
        DS_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    }
    if (ctx->opt["precision"] == DS_PRECISION_FAST)
        DS_CUDA(ctx, launchDisneyMlpTc(ctx->model, dIn, nRows, dOut, ctx->stream, prof >= 2 ? ctx->model.prof : nullptr, mlpOperands(ctx)));
    else
        DS_CUDA(ctx, launchDisneyMlpF32(ctx->model, dIn, nullptr, nRows, dOut, ctx->stream));
    ctx->launches += 1;
    if (prof) {
        float ms = 0.0f;
        DS_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        DS_CUDA(ctx, cudaEventSynchronize(e1));
        DS_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        ctx->opt["mlp_last_us"] = (int)(ms * 1000.0f + 0.5f);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return DS_OK;
}

static int disneyCheckError(DsContext* ctx)
{
    uint32_t err = 0;
    DS_CUDA(ctx, cudaMemcpyAsync(&err, ctx->model.error, sizeof(err), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (err) {
        cudaMemset(ctx->model.error, 0, sizeof(uint32_t));
        DS_FAIL(ctx, DS_ERR_CUDA, "k_disney_mlp_tc: barrier wait timed out in block %u", err - 1u);
    }
    return DS_OK;
}

int ds_disney_model_profile(DsContext* ctx, uint64_t* cycles16)
{
    DS_CHECK_CTX(ctx);
    if (!cycles16) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    if (!ctx->model.loaded) DS_FAIL(ctx, DS_ERR_STATE, "no model loaded (ds_disney_model_load)");
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DS_CUDA(ctx, cudaMemcpy(cycles16, ctx->model.prof, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DS_OK;
}

int ds_disney_model_forward(DsContext* ctx, const float* network_input, uint32_t n, float* predicted_out)
{
    DS_CHECK_CTX(ctx);
    if (!network_input || !predicted_out) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    if (!ctx->model.loaded) DS_FAIL(ctx, DS_ERR_STATE, "no model loaded (ds_disney_model_load)");
    if (n == 0) return DS_OK;
    int rc;
    if ((rc = ensureMlpScratch(ctx, 0, (size_t)n * 2260 * sizeof(float))) || (rc = ensureMlpScratch(ctx, 1, (size_t)n * sizeof(float)))) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(ctx->mlpScratch[0], network_input, (size_t)n * 2260 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    const float* dIn = (const float*)ctx->mlpScratch[0];
    if (ctx->opt["precision"] == DS_PRECISION_FAST) {
        /* rows -> the tiles the tensor-core kernel reads (the renderer's descriptor gather writes tiles directly) */
        if ((rc = ensureMlpScratch(ctx, 4, networkTileBytes(ctx, n)))) return rc;
        DS_CUDA(ctx, launchNetworkInputToTiles(dIn, n, ctx->mlpScratch[4], ctx->stream, mlpOperands(ctx)));
        ctx->launches += 1;
        dIn = (const float*)ctx->mlpScratch[4];
    }
    if ((rc = disneyForwardDevice(ctx, dIn, n, (float*)ctx->mlpScratch[1]))) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(predicted_out, ctx->mlpScratch[1], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return disneyCheckError(ctx);
}

/* DisneyRenderer::render (DG/Scene/Cameras/DisneyRenderer.cpp:58-110).  The reference walks the frame in 128 x 128 rectangles because its
 * network-input tensor holds one rectangle; per rectangle: the network-input launch, the model (skipped when nothing scattered, :91-99),
 * copyToFrameResult.  Pixels do not interact, so here the WHOLE frame goes through each stage in one launch -- every pixel with the seed and
 * stream of its own rectangle, so the result is the reference's pixel for pixel -- and only the pixels that scattered reach the descriptor
 * gather and the model, compacted, in batches sized for HBM (2.4 GB of network input per batch):
 *   k_network_info (frame) -> k_compact_active -> per batch: k_descriptors (gather) -> model -> k_blit_predicted (scatter) */
static int renderDisneyDevice(DsContext* ctx, const DsCamera* cam, uint32_t frame_width, uint32_t frame_height, uint32_t stream, float4** frame_out)
{
    int rc = requireScene(ctx, true);
    if (rc) return rc;
    if (!cam) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    if (frame_width == 0 || frame_height == 0 || (unsigned long long)frame_width * frame_height > (1ull << 26))
        DS_FAIL(ctx, DS_ERR_INVALID, "bad frame size");
    if (!ctx->model.loaded) DS_FAIL(ctx, DS_ERR_STATE, "no model loaded (ds_disney_model_load)");
    const int RECT = 128;            /* DisneyRenderer.cpp:10 */
    const uint32_t BATCH = 1u << 18; /* rows of network input resident at a time */
    const size_t pixels = (size_t)frame_width * frame_height;
    const size_t batchRows = std::min<size_t>(BATCH, pixels);
    /* FAST: the descriptor gather writes the tiles the tensor-core kernel reads; EXACT: DisneyNetworkInput rows for the fp32 kernel */
    const bool tiled = ctx->opt["precision"] == DS_PRECISION_FAST;
    if ((rc = ensureScratch(ctx, 0, pixels * 3 * sizeof(float))) || (rc = ensureScratch(ctx, 1, pixels * 3 * sizeof(float))) ||
        (rc = ensureScratch(ctx, 2, pixels * (sizeof(float) + 1))) ||
        (rc = ensureScratch(ctx, 3, tiled ? networkTileBytes(ctx, batchRows) : batchRows * 2260 * sizeof(float))) ||
        (rc = ensureScratch(ctx, 4, pixels * 5 * sizeof(float))) || (rc = ensureMlpScratch(ctx, 1, batchRows * sizeof(float))) ||
        (rc = ensureMlpScratch(ctx, 2, (pixels + 1) * sizeof(uint32_t))) || (rc = ensureMlpScratch(ctx, 3, pixels * sizeof(float4))))
        return rc;
    float* dPos = (float*)ctx->scratch[0];
    float* dDir = (float*)ctx->scratch[1];
    float* dAngle = (float*)ctx->scratch[2];
    uint8_t* dActive = (uint8_t*)(dAngle + pixels);
    float* dInput = (float*)ctx->scratch[3];
    float* dInfo = (float*)ctx->scratch[4];
    float* dPred = (float*)ctx->mlpScratch[1];
    uint32_t* dIdx = (uint32_t*)ctx->mlpScratch[2];
    uint32_t* dCount = dIdx + pixels;
    float4* dFrame = (float4*)ctx->mlpScratch[3];
    DS_CUDA(ctx, cudaMemsetAsync(dFrame, 0, pixels * sizeof(float4), ctx->stream));
    TraceJob job;
    memset(&job, 0, sizeof(job));
    memcpy(job.eye, cam->eye, 12);
    memcpy(job.U, cam->U, 12);
    memcpy(job.V, cam->V, 12);
    memcpy(job.W, cam->W, 12);
    job.width = (int)frame_width;
    job.height = (int)frame_height;
    DevScene sc;
    fillDevScene(ctx, sc);
    /* FAST: the empty-space leg of every camera ray, walked once per (camera, frame size, volume) by k_primary_prepass as for the path
     * tracer -- pixels that never reach an occupied cell have transmittance 1 and do not scatter; the others start at the cloud */
    const uint32_t* dEntry = nullptr;
    if (tiled && ctx->opt["skip_empty"] != 0 && ctx->opt["primary_cache"] != 0) {
        if ((rc = ensureMlpScratch(ctx, 5, pixels * sizeof(uint32_t))) || (rc = ensureMlpScratch(ctx, 6, (pixels + 8) * sizeof(uint32_t)))) return rc;
        uint32_t* entry = (uint32_t*)ctx->mlpScratch[5];
        if (!ctx->disneyPrimaryValid || ctx->disneyPrimaryW != frame_width || ctx->disneyPrimaryH != frame_height ||
            memcmp(&ctx->disneyPrimaryCam, cam, sizeof(DsCamera)) != 0) {
            TraceJob pj = job;
            pj.tilesX = ((int)frame_width + 7) / 8;
            pj.itemsPerSubframe = (unsigned long long)pj.tilesX * (((int)frame_height + 3) / 4) * 32ull;
            uint32_t* hitList = (uint32_t*)ctx->mlpScratch[6];
            unsigned long long* counts = (unsigned long long*)(hitList + ((pixels + 1) & ~(size_t)1)); /* 8-byte aligned tail of slot 6 */
            DS_CUDA(ctx, cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), ctx->stream));
            DS_CUDA(ctx, KernelSet<true>::primaryPrepass(sc, pj, entry, hitList, counts, ctx->stream));
            ctx->launches++;
            ctx->disneyPrimaryCam = *cam;
            ctx->disneyPrimaryW = frame_width;
            ctx->disneyPrimaryH = frame_height;
            ctx->disneyPrimaryValid = true;
        }
        dEntry = entry;
    }
    if (ctx->opt["precision"] == DS_PRECISION_FAST)
        DS_CUDA(ctx, KernelSet<true>::networkInfo(sc, job, 0, 0, (int)frame_width, (int)frame_height, stream, dInfo, dPos, dDir, dAngle, dActive,
                                                  ctx->stats, ctx->stream, RECT, dEntry));
    else
        DS_CUDA(ctx, KernelSet<false>::networkInfo(sc, job, 0, 0, (int)frame_width, (int)frame_height, stream, dInfo, dPos, dDir, dAngle, dActive,
                                                   ctx->stats, ctx->stream, RECT));
    DS_CUDA(ctx, launchCompactActive(dActive, (uint32_t)pixels, dIdx, dCount, ctx->stream));
    uint32_t nActive = 0;
    DS_CUDA(ctx, cudaMemcpyAsync(&nActive, dCount, sizeof(nActive), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += 2;
    if (ctx->opt["compact_reverse"]) DS_CUDA(ctx, launchReverse(dIdx, nActive, ctx->stream));
    LevelTable lv;
    DescriptorLayers layers;
    fillDescriptorTables(ctx, lv, layers);
    cudaTextureObject_t mipTex = 0;
    if ((rc = descriptorTexture(ctx, true, &mipTex))) return rc;
    for (uint32_t first = 0; first < nActive; first += BATCH) {
        const uint32_t n = std::min(BATCH, nActive - first);
        DS_CUDA(ctx, launchDescriptors(sc, lv, layers, dPos, dDir, n, nullptr, dInput, nullptr, ctx->stream, tiled ? -mlpOperands(ctx) : 226, dAngle, nullptr, dIdx + first,
                                       mipTex));
        if ((rc = disneyForwardDevice(ctx, dInput, n, dPred))) return rc;
        DS_CUDA(ctx, launchBlitPredicted(dPred, dInfo, dIdx + first, n, dFrame, ctx->stream));
        ctx->launches += 2;
    }
    *frame_out = dFrame;
    return DS_OK;
}

int ds_render_disney(DsContext* ctx, const DsCamera* cam, uint32_t frame_width, uint32_t frame_height, uint32_t stream, float* frame_result_out)
{
    DS_CHECK_CTX(ctx);
    if (!frame_result_out) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    float4* dFrame = nullptr;
    int rc = renderDisneyDevice(ctx, cam, frame_width, frame_height, stream, &dFrame);
    if (rc) return rc;
    DS_CUDA(ctx, cudaMemcpyAsync(frame_result_out, dFrame, (size_t)frame_width * frame_height * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    return disneyCheckError(ctx);
}

/* Camera::render with DisneyRenderer as the ARenderer (Camera.cpp:189-199): per subframe the neural frame, then updateFrameResult into the
 * progressive / variance buffers of the context's frame -- all on the device.  Subframe s uses stream s * 4096 (+ rectangle ordinal). */
int ds_render_disney_subframes(DsContext* ctx, const DsCamera* cam, uint32_t first_subframe, uint32_t n)
{
    DS_CHECK_CTX(ctx);
    if (!ctx->progressive) DS_FAIL(ctx, DS_ERR_STATE, "no frame (ds_frame_create)");
    if (first_subframe == 0) DS_FAIL(ctx, DS_ERR_INVALID, "subframe ids are 1-based (Camera.cpp:191)");
    const size_t px = (size_t)ctx->width * ctx->height;
    for (uint32_t k = 0; k < n; ++k) {
        float4* dFrame = nullptr;
        int rc = renderDisneyDevice(ctx, cam, (uint32_t)ctx->width, (uint32_t)ctx->height, (first_subframe + k) * 4096u, &dFrame);
        if (rc) return rc;
        DS_CUDA(ctx, launchUpdateFrame(dFrame, nullptr, ctx->progressive, ctx->variance, px, first_subframe + k, 1, ctx->stream));
        ctx->launches += 1;
    }
    return disneyCheckError(ctx);
}

int ds_blit_predicted(uint32_t frame_width, uint32_t frame_height, uint32_t rect_x, uint32_t rect_y, uint32_t rect_w, uint32_t rect_h,
                      const float* predicted, const DsIntersectionInfo* info, float* frame_result_inout)
{
    if (!predicted || !info || !frame_result_inout) return DS_ERR_INVALID;
    if ((unsigned long long)rect_x + rect_w > frame_width || (unsigned long long)rect_y + rect_h > frame_height) return DS_ERR_INVALID;
    for (uint32_t y = 0; y < rect_h; y++)
        for (uint32_t x = 0; x < rect_w; x++) {
            const size_t i = (size_t)y * rect_w + x;
            if (!info[i].has_scattered) continue;
            float* px = frame_result_inout + 4 * ((size_t)(y + rect_y) * frame_width + (x + rect_x));
            const float w = 1 - info[i].transmittance;
            /* (make_float4(predicted) + make_float4(radiance)) * (1 - transmittance): make_float4(float3) sets w = 0 */
            px[0] = (predicted[i] + info[i].radiance[0]) * w;
            px[1] = (predicted[i] + info[i].radiance[1]) * w;
            px[2] = (predicted[i] + info[i].radiance[2]) * w;
            px[3] = (predicted[i] + 0.0f) * w;
        }
    return DS_OK;
}

void ds_radiance_settings_default(DsRadianceSettings* s)
{
    if (!s) return;
    s->max_thread_count = 10 * 2048; /* RadianceCollector.cpp:17 */
    s->launches_per_update = 100;    /* :88 */
    s->max_updates = 0;
    s->relative_ci = 2e-2f;          /* :113 */
    s->absolute_ci = 1e-4f;          /* :114 */
    s->zero_radiance_min_experiments = 100000; /* :117 */
}

/* CU/PointRadianceTask.h:23-36 */
static float absoluteCI(const DsPointRadianceTask& t)
{
    const float N = (float)t.experiment_count;
    const float sigma = sqrtf(t.running_variance / N);
    return 1.96f * sigma / sqrtf(N);
}
static float relativeCI(const DsPointRadianceTask& t) { return absoluteCI(t) / (t.radiance + FLT_EPSILON); }

/* CU/PointRadianceTask.h:54-68 */
static void mergeTask(DsPointRadianceTask& a, const DsPointRadianceTask& other)
{
    const float newWeight = other.experiment_count * 1.0f / (a.experiment_count + other.experiment_count);
    a.radiance += (other.radiance - a.radiance) * newWeight;
    a.running_variance += other.running_variance;
    a.experiment_count += other.experiment_count;
}

int ds_point_radiance_run(DsContext* ctx, const float* positions, const float* directions, uint32_t n, const DsRadianceSettings* settings,
                          DsPointRadianceTask* tasks_out, uint8_t* converged_out, uint32_t* updates_out)
{
    DS_CHECK_CTX(ctx);
    int rc = requireScene(ctx, true);
    if (rc) return rc;
    DsRadianceSettings cfg;
    if (settings)
        cfg = *settings;
    else
        ds_radiance_settings_default(&cfg);
    if (n == 0) return DS_OK;
    if (!positions || !directions || !tasks_out || !converged_out) DS_FAIL(ctx, DS_ERR_INVALID, "NULL argument");
    if (cfg.max_thread_count < n) DS_FAIL(ctx, DS_ERR_INVALID, "taskRepeatCount would be 0 (RadianceCollector.cpp:179): max_thread_count < n");
    if (cfg.launches_per_update == 0) DS_FAIL(ctx, DS_ERR_INVALID, "launches_per_update must be > 0");
    if ((unsigned long long)cfg.max_thread_count * cfg.launches_per_update >= (1ull << 32))
        DS_FAIL(ctx, DS_ERR_INVALID, "max_thread_count x launches_per_update must stay below 2^32");

    /* FAST flavour, option "radiance_scheduler" = 1 (default): the device-resident collector (AdaptiveCollector,
     * ds_kernels.h) -- one launch for the whole batch, same convergence rule, no host round trips.  The reference
     * schedule below stays available (option = 0) and is what the EXACT flavour always runs (bit-exact with the oracle). */
    if (ctx->opt["precision"] == DS_PRECISION_FAST && ctx->opt["radiance_scheduler"] == 1 && ctx->borderEmpty && ctx->params.sample_step <= 0.01f &&
        ctx->opt["skip_empty"] && ctx->opt["variant"] == 0) {
        const uint32_t repeat = cfg.max_thread_count / n;
        const size_t perSample = 8 + 8 + 8 + 4 + 4;
        const size_t stateBytes = (size_t)n * perSample + 64;
        if ((rc = ensureScratch(ctx, 5, (size_t)n * sizeof(DsPointRadianceTask))) || (rc = ensureScratch(ctx, 6, stateBytes))) return rc;
        std::vector<DsPointRadianceTask> samples(n);
        for (uint32_t i = 0; i < n; i++) {
            DsPointRadianceTask t{};
            t.id = (int32_t)i;
            memcpy(t.position, positions + 3 * i, 12);
            memcpy(t.direction, directions + 3 * i, 12);
            samples[i] = t;
        }
        DsPointRadianceTask* dTasks = (DsPointRadianceTask*)ctx->scratch[5];
        uint8_t* state = (uint8_t*)ctx->scratch[6];
        DS_CUDA(ctx, cudaMemcpyAsync(dTasks, samples.data(), (size_t)n * sizeof(DsPointRadianceTask), cudaMemcpyHostToDevice, ctx->stream));
        DS_CUDA(ctx, cudaMemsetAsync(state, 0, stateBytes, ctx->stream));
        TraceJob job;
        memset(&job, 0, sizeof(job));
        job.kind = JOB_ADAPTIVE;
        job.mode = DS_MODE_SUN_MULTIPLE_SCATTER; /* Tasks.cpp:134 */
        job.tasks = dTasks;
        job.total = ~0ull >> 1;
        AdaptiveCollector& ad = job.ad;
        ad.sum = (double*)state;
        ad.sumSq = ad.sum + n;
        ad.count = (unsigned long long*)(ad.sumSq + n);
        ad.issued = (uint32_t*)(ad.count + n);
        ad.flag = ad.issued + n;
        ad.closed = ad.flag + n;
        ad.ticket = (unsigned long long*)(state + (((size_t)n * perSample + 8 + 7) & ~(size_t)7));
        ad.nSamples = n;
        ad.quota = (uint32_t)std::max(1, ctx->opt["radiance_quota"]);
        /* the reference's first convergence test sees repeat x launches_per_update experiments per sample */
        ad.minExperiments = repeat * cfg.launches_per_update;
        ad.maxExperiments = cfg.max_updates ? (uint32_t)std::min<unsigned long long>(0xffffffffull, (unsigned long long)cfg.max_updates * ad.minExperiments) : 0u;
        ad.zeroMin = cfg.zero_radiance_min_experiments;
        ad.relCI = cfg.relative_ci;
        ad.absCI = cfg.absolute_ci;
        ad.m2Scale = cfg.launches_per_update > 1 ? (float)(cfg.launches_per_update - 1) / (float)cfg.launches_per_update : 1.0f;
        rc = runTrace(ctx, job);
        if (rc) return rc;
        std::vector<uint8_t> host(stateBytes);
        DS_CUDA(ctx, cudaMemcpyAsync(host.data(), state, stateBytes, cudaMemcpyDeviceToHost, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const double* hSum = (const double*)host.data();
        const double* hSq = hSum + n;
        const unsigned long long* hCount = (const unsigned long long*)(hSq + n);
        const uint32_t* hFlag = (const uint32_t*)(hCount + n) + n;
        for (uint32_t i = 0; i < n; i++) {
            DsPointRadianceTask t = samples[i];
            const double N = (double)hCount[i];
            const double mean = N > 0 ? hSum[i] / N : 0.0;
            t.experiment_count = (uint32_t)std::min<unsigned long long>(hCount[i], 0xffffffffull);
            t.radiance = (float)mean;
            t.running_variance = (float)(std::max(hSq[i] - N * mean * mean, 0.0) * (double)ad.m2Scale);
            tasks_out[i] = t;
            converged_out[i] = hFlag[i] == 1u ? 1 : 0;
        }
        if (updates_out) *updates_out = 1;
        return DS_OK;
    }

    /* RadianceCollector::init (:27-47) */
    std::vector<DsPointRadianceTask> todo(n);
    for (uint32_t i = 0; i < n; i++) {
        DsPointRadianceTask t{};
        t.id = (int32_t)i;
        memcpy(t.position, positions + 3 * i, 12);
        memcpy(t.direction, directions + 3 * i, 12);
        todo[i] = t;
        converged_out[i] = 0;
        tasks_out[i] = t;
    }
    const size_t taskBytes = (size_t)cfg.max_thread_count * sizeof(DsPointRadianceTask);
    const size_t xBytes = (size_t)cfg.max_thread_count * cfg.launches_per_update * sizeof(float);
    if ((rc = ensureScratch(ctx, 5, taskBytes)) || (rc = ensureScratch(ctx, 6, xBytes))) return rc;
    DsPointRadianceTask* dTasks = (DsPointRadianceTask*)ctx->scratch[5];
    float* dX = (float*)ctx->scratch[6];
    std::vector<DsPointRadianceTask> threads;
    uint32_t frameId = 0, updates = 0;
    while (!todo.empty() && (cfg.max_updates == 0 || updates < cfg.max_updates)) {
        /* scheduleTasks (:176-192) */
        const uint32_t taskRepeatCount = cfg.max_thread_count / (uint32_t)todo.size();
        const uint32_t threadsCount = (uint32_t)todo.size() * taskRepeatCount;
        threads.assign(threadsCount, DsPointRadianceTask{});
        for (uint32_t i = 0; i < todo.size(); i++) {
            threads[(size_t)i * taskRepeatCount] = todo[i];
            for (uint32_t j = 1; j < taskRepeatCount; j++) {
                DsPointRadianceTask f{};
                f.id = todo[i].id;
                memcpy(f.position, todo[i].position, 12);
                memcpy(f.direction, todo[i].direction, 12);
                threads[(size_t)i * taskRepeatCount + j] = f;
            }
        }
        DS_CUDA(ctx, cudaMemcpyAsync(dTasks, threads.data(), (size_t)threadsCount * sizeof(DsPointRadianceTask), cudaMemcpyHostToDevice, ctx->stream));
        /* update (:88-93): launches_per_update launches of estimateEmission, fused into one queue */
        TraceJob job;
        memset(&job, 0, sizeof(job));
        job.kind = JOB_POINT;
        job.mode = DS_MODE_SUN_MULTIPLE_SCATTER; /* Tasks.cpp:134 */
        job.tasks = dTasks;
        job.launches = cfg.launches_per_update;
        job.frame0 = frameId;
        job.xOut = dX;
        job.total = (unsigned long long)threadsCount * cfg.launches_per_update;
        rc = runTrace(ctx, job);
        if (rc) return rc;
        DS_CUDA(ctx, launchTaskWelford(dTasks, dX, threadsCount, cfg.launches_per_update, ctx->stream));
        DS_CUDA(ctx, cudaMemcpyAsync(threads.data(), dTasks, (size_t)threadsCount * sizeof(DsPointRadianceTask), cudaMemcpyDeviceToHost, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        frameId += cfg.launches_per_update;
        updates++;
        /* merge repeats + convergence test (:100-130) */
        std::vector<DsPointRadianceTask> next;
        for (uint32_t i = 0; i < todo.size(); i++) {
            DsPointRadianceTask& representative = threads[(size_t)i * taskRepeatCount];
            for (uint32_t j = 1; j < taskRepeatCount; j++) mergeTask(representative, threads[(size_t)i * taskRepeatCount + j]);
            bool isConverged = relativeCI(representative) < cfg.relative_ci || absoluteCI(representative) < cfg.absolute_ci;
            if (representative.radiance < FLT_EPSILON) isConverged = representative.experiment_count > cfg.zero_radiance_min_experiments;
            tasks_out[representative.id] = representative;
            if (isConverged)
                converged_out[representative.id] = 1;
            else
                next.push_back(representative);
        }
        todo.swap(next);
    }
    if (updates_out) *updates_out = updates;
    return DS_OK;
}

} /* extern "C" */
