// Microbenchmark (development tool, not part of the library): the memory-side ceilings of the path-tracing kernels, which gather
// trilinear u8 taps at data-dependent addresses -- i.e. they live on L2 sector bandwidth and texture-unit rate, not on HBM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather l2_gather.cu && ./l2_gather > l2_peaks.json
// Measures (all with CUDA events, best of 5 after a warm-up):
//   l2_stream_gbs        coalesced 16-byte ld.global.cg loads over a buffer that fits L2 (48 MiB), re-read many times
//   l2_sector_gather_gbs random 32-byte-sector gathers (one 4-byte load per random sector) over the same L2-resident buffer:
//                        sectors/s * 32 B
//   dram_sector_gather_gbs the same gather over a 4 GiB buffer (far beyond L2): DRAM random-sector rate
//   tex3d_l2_gtaps       hardware trilinear tex3D (u8, normalized float) at random coordinates in a 320^3 array (L2-resident)
//   tex3d_l1_gtaps       the same in a 32^3 array (L1-resident): the texture units' filter rate
//   tex3d_dram_gtaps     the same in a 1024^3 array (1 GiB, DRAM-resident)
// Output: one JSON object on stdout.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ uint32_t lowbias32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(1024) k_stream(const uint4* __restrict__ buf, size_t n16, int passes, uint32_t* sink)
{
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i); /* .cg: cached in L2 only, so the L1 cannot serve re-reads */
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) sink[0] = acc;
}

// each thread issues `perThread` independent 4-byte loads, each from its own random 32-byte sector; 4 loads in flight per thread
__global__ void __launch_bounds__(1024) k_sector_gather(const uint32_t* __restrict__ buf, size_t sectors, int perThread, uint32_t seed, uint32_t* sink)
{
    uint32_t acc = 0;
    uint32_t s = lowbias32((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed);
    for (int i = 0; i < perThread; i += 4) {
        uint32_t a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            s = s * 1664525u + 1013904223u;
            const size_t sector = (size_t)(((unsigned long long)lowbias32(s) * sectors) >> 32);
            a[k] = __ldcg(buf + sector * 8 + (s & 7u)); /* .cg: L2 only */
        }
        acc += a[0] ^ a[1] ^ a[2] ^ a[3];
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void __launch_bounds__(1024) k_tex_gather(cudaTextureObject_t tex, int perThread, uint32_t seed, float* sink)
{
    float acc = 0;
    uint32_t s = lowbias32((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed);
    for (int i = 0; i < perThread; i += 4) {
        float a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            s = s * 1664525u + 1013904223u;
            const uint32_t h = lowbias32(s);
            const float u = (float)(h & 0x3ffu) * (1.0f / 1024.0f), v = (float)((h >> 10) & 0x3ffu) * (1.0f / 1024.0f),
                        w = (float)((h >> 20) & 0x3ffu) * (1.0f / 1024.0f);
            a[k] = tex3D<float>(tex, u, v, w);
        }
        acc += a[0] + a[1] + a[2] + a[3];
    }
    if (acc == -1.0f) sink[0] = acc;
}

// coherent variant: the lanes of a warp march along nearby rays (what the renderer does): consecutive taps 1/512 apart
__global__ void __launch_bounds__(1024) k_tex_march(cudaTextureObject_t tex, int perThread, uint32_t seed, float* sink)
{
    float acc = 0;
    const uint32_t h = lowbias32((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed);
    float u = (float)(h & 0x3ffu) * (1.0f / 1024.0f), v = (float)((h >> 10) & 0x3ffu) * (1.0f / 1024.0f), w = (float)((h >> 20) & 0x3ffu) * (1.0f / 1024.0f);
    const float du = 0.0011f, dv = 0.0013f, dw = 0.0009f;
    for (int i = 0; i < perThread; i += 4) {
        float a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u += du; v += dv; w += dw;
            if (u > 1.f) u -= 1.f;
            if (v > 1.f) v -= 1.f;
            if (w > 1.f) w -= 1.f;
            a[k] = tex3D<float>(tex, u, v, w);
        }
        acc += a[0] + a[1] + a[2] + a[3];
    }
    if (acc == -1.0f) sink[0] = acc;
}

template <class F> static float bestMs(F launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static cudaTextureObject_t makeTex(int n, cudaArray_t* arrOut)
{
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<unsigned char>();
    cudaArray_t arr;
    CK(cudaMalloc3DArray(&arr, &desc, make_cudaExtent(n, n, n)));
    std::vector<unsigned char> host((size_t)n * n * n);
    uint32_t s = 12345u;
    for (auto& b : host) {
        s = s * 1664525u + 1013904223u;
        b = (unsigned char)(s >> 24);
    }
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(host.data(), n, n, n);
    p.dstArray = arr;
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    *arrOut = arr;
    return tex;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int grid = sms * 2, block = 1024;
    uint32_t* sink;
    CK(cudaMalloc(&sink, 64));

    // ---- linear buffers
    const size_t l2Bytes = 48ull << 20, dramBytes = 4ull << 30;
    uint32_t* big;
    CK(cudaMalloc(&big, dramBytes));
    CK(cudaMemset(big, 1, dramBytes));
    const int passes = 64;
    const float msStream = bestMs([&] { k_stream<<<grid, block>>>((const uint4*)big, l2Bytes / 16, passes, sink); });
    const double streamGbs = (double)l2Bytes * passes / (msStream * 1e-3) / 1e9;
    const int perThread = 2048;
    const double gathers = (double)grid * block * perThread;
    const float msL2 = bestMs([&] { k_sector_gather<<<grid, block>>>(big, l2Bytes / 32, perThread, 1u, sink); });
    const float msDram = bestMs([&] { k_sector_gather<<<grid, block>>>(big, dramBytes / 32, perThread / 4, 2u, sink); });
    const double l2GatherGbs = gathers * 32.0 / (msL2 * 1e-3) / 1e9;
    const double dramGatherGbs = gathers / 4 * 32.0 / (msDram * 1e-3) / 1e9;
    CK(cudaFree(big));

    // ---- textures
    double texRandom[3], texMarch[3];
    const int sizes[3] = {32, 320, 1024};
    for (int t = 0; t < 3; t++) {
        cudaArray_t arr;
        cudaTextureObject_t tex = makeTex(sizes[t], &arr);
        const int per = t == 2 ? 512 : 2048;
        const float ms = bestMs([&] { k_tex_gather<<<grid, block>>>(tex, per, 3u, (float*)sink); });
        texRandom[t] = (double)grid * block * per / (ms * 1e-3) / 1e9;
        const float ms2 = bestMs([&] { k_tex_march<<<grid, block>>>(tex, per, 4u, (float*)sink); });
        texMarch[t] = (double)grid * block * per / (ms2 * 1e-3) / 1e9;
        CK(cudaDestroyTextureObject(tex));
        CK(cudaFreeArray(arr));
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"sm_clock_khz\": %d, "
           "\"l2_stream_gbs\": %.1f, \"l2_sector_gather_gbs\": %.1f, \"l2_sector_gather_gsectors\": %.2f, \"dram_sector_gather_gbs\": %.1f, "
           "\"tex3d_random_gtaps\": {\"l1_32\": %.2f, \"l2_320\": %.2f, \"dram_1024\": %.2f}, "
           "\"tex3d_march_gtaps\": {\"l1_32\": %.2f, \"l2_320\": %.2f, \"dram_1024\": %.2f}, "
           "\"how\": \"tools/microbench/l2_gather.cu: CUDA events, best of 5; gather = one 4-byte load per random 32-byte sector, 4 in flight per thread, %d x 1024 threads\"}\n",
           prop.name, sms, prop.l2CacheSize, prop.clockRate, streamGbs, l2GatherGbs, l2GatherGbs / 32.0, dramGatherGbs, texRandom[0], texRandom[1], texRandom[2],
           texMarch[0], texMarch[1], texMarch[2], grid);
    return 0;
}
