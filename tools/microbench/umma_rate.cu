// Microbenchmark (development tool, not part of the library): cycles per tcgen05.mma for the operand layouts and shapes the model kernel
// could use.  One CTA per SM, one thread issues `iters` back-to-back MMAs on garbage operands in shared memory, commits, waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KIND>  // 0 = tf32, 1 = bf16 (kind::f16)
__device__ __forceinline__ void mma(uint32_t d, uint32_t aLo, uint32_t bLo, uint32_t hi, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(aLo), "r"(bLo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(aLo), "r"(bLo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}

// layout 0: no swizzle (core matrices 8 x 16 B, SBO 128, LBO = rows/8*128); layout 2: 128-byte swizzle (rows of 128 B, SBO 1024)
template <int KIND>
__global__ void k_rate(int N, int layout, int iters, int kSteps, unsigned long long* out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmemSlot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u; /* small finite values */
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&tmemSlot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemSlot;
    if (threadIdx.x == 0) {
        const uint32_t aBase = smemAddr(smem), bBase = aBase + 64 * 1024;
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t aLo, bLo, hi, aStep, bStep;
        if (layout == 0) {
            const uint32_t aLbo = 128 / 8 * 128, bLbo = (uint32_t)N / 8 * 128;
            aLo = ((aBase & 0x3ffffu) >> 4) | ((aLbo >> 4) << 16);
            bLo = ((bBase & 0x3ffffu) >> 4) | ((bLbo >> 4) << 16);
            hi = (128u >> 4) | (1u << 14);
            aStep = 2 * aLbo >> 4;
            bStep = 2 * bLbo >> 4;
        } else {
            aLo = ((aBase & 0x3ffffu) >> 4) | (1u << 16);
            bLo = ((bBase & 0x3ffffu) >> 4) | (1u << 16);
            hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            aStep = bStep = 32 >> 4; /* 32 bytes further along the 128-byte row */
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i)
            for (int j = 0; j < kSteps; ++j) mma<KIND>(tmem, aLo + j * aStep, bLo + j * bStep, hi, idesc, (i | j) ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(&bar)) : "memory");
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smemAddr(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = (unsigned long long)(t1 - t0);
            out[1] = (unsigned long long)(t2 - t0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main()
{
    unsigned long long* d;
    cudaMalloc(&d, 16);
    const int smem = 160 * 1024;
    cudaFuncSetAttribute(k_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 500, kSteps = 4;
    for (int kind = 0; kind < 2; ++kind)
        for (int layout = 0; layout <= 2; layout += 2)
            for (int N : {208, 112, 256}) {
                for (int blocks : {1, 148}) {
                    if (kind == 0) k_rate<0><<<blocks, 128, smem>>>(N, layout, iters, kSteps, d);
                    else k_rate<1><<<blocks, 128, smem>>>(N, layout, iters, kSteps, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    unsigned long long h[2] = {0, 0};
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    const double per = (double)h[1] / (iters * kSteps);
                    const int kPer = kind == 0 ? 8 : 16;
                    printf("%s layout=%s M=128 N=%d K=%d blocks=%d: issue %.1f cyc/mma, complete %.1f cyc/mma, %.0f MAC/clk/SM (%s)\n", kind == 0 ? "tf32" : "bf16",
                           layout == 0 ? "none " : "sw128", N, kPer, blocks, (double)h[0] / (iters * kSteps), per, 128.0 * N * kPer / per, cudaGetErrorString(e));
                }
            }
    return 0;
}
