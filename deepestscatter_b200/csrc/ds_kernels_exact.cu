/*
 * ds_kernels_exact.cu -- translation unit compiled with -fmad=false: the bit-reproducible flavour of
 * the estimator kernels plus every kernel whose output the parity tests compare exactly (mip chain,
 * occupancy mask, Welford accumulation, descriptors, tone map, moments).
 */
#include "ds_kernels.cuh"

#include "../../include/ds_synth.h"

namespace dsk {

/* the FAST instantiations live in ds_kernels_fast.cu, compiled with other flags: never instantiate them here (two copies of one kernel
 * with different arithmetic would be an ODR violation -- which copy a launch gets is then decided per process) */
extern template struct KernelSet<true>;

template <>
cudaError_t KernelSet<false>::trace(const DevScene& sc, const TraceJob& job, const LaunchConfig& cfg, cudaStream_t st)
{
    return traceGeneric<false>(sc, job, cfg, st);
}

template <>
cudaError_t KernelSet<false>::primaryPrepass(const DevScene&, const TraceJob&, uint32_t*, uint32_t*, unsigned long long*, cudaStream_t)
{
    return cudaErrorNotSupported; /* the exact flavour marches every step like the reference */
}

template struct KernelSet<false>;

/* ---- synthetic grid (include/ds_synth.h) ---- */
__global__ void __launch_bounds__(256) k_synth(uint8_t* out, int n, int kind, uint32_t seed)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const int y = blockIdx.y, z = blockIdx.z;
    out[((size_t)z * n + y) * n + x] = ds_synth_voxel(kind, seed, n, x, y, z);
}

cudaError_t launchSynth(uint8_t* out, int n, int kind, uint32_t seed, cudaStream_t st)
{
    dim3 grid((n + 255) / 256, n, n);
    k_synth<<<grid, 256, 0, st>>>(out, n, kind, seed);
    return cudaGetLastError();
}

/* ---- DG/Util/Resources.cpp:137: narrow_cast<uint8_t>(value / maxDensity * 255), double arithmetic ---- */
__global__ void __launch_bounds__(256) k_quantize(const float* __restrict__ in, size_t count, double maxDensity, uint8_t* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = (uint8_t)((double)in[i] / maxDensity * 255);
}

cudaError_t launchQuantize(const float* in, size_t count, double maxDensity, uint8_t* out, cudaStream_t st)
{
    k_quantize<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(in, count, maxDensity, out);
    return cudaGetLastError();
}

/* ---- DG/Util/Resources.cpp:169-209: one mip level, 8 children (out of range = 0), integer / 8 ---- */
__global__ void __launch_bounds__(256) k_mip(const uint8_t* __restrict__ prev, int pnx, int pny, int pnz, uint8_t* __restrict__ cur, int cnx,
                                             int cny, int cnz)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cnx) return;
    const int y = blockIdx.y, z = blockIdx.z;
    unsigned sum = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int sx = 2 * x + (c & 1), sy = 2 * y + ((c >> 1) & 1), sz = 2 * z + (c >> 2);
        if (sx < pnx && sy < pny && sz < pnz) sum += prev[((size_t)sz * pny + sy) * pnx + sx];
    }
    cur[((size_t)z * cny + y) * cnx + x] = (uint8_t)(sum / 8u);
}

cudaError_t launchMip(const uint8_t* prev, int pnx, int pny, int pnz, uint8_t* cur, int cnx, int cny, int cnz, cudaStream_t st)
{
    dim3 grid((cnx + 255) / 256, cny, cnz);
    k_mip<<<grid, 256, 0, st>>>(prev, pnx, pny, pnz, cur, cnx, cny, cnz);
    return cudaGetLastError();
}

/* ---- occupancy bit mask: cell (cx,cy,cz) covers voxels [c*2^s, min(c*2^s + 2^s, N-1)] per axis ---- */
__global__ void __launch_bounds__(128) k_occupancy(const uint8_t* __restrict__ density, int nx, int ny, int nz, int shift, int ocx, int ocy,
                                                   int ocz, uint32_t* bits)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ocx * ocy * ocz) return;
    const int cx = cell % ocx, cy = (cell / ocx) % ocy, cz = cell / (ocx * ocy);
    const int c = 1 << shift;
    const int x0 = cx * c, x1 = min(x0 + c, nx - 1);
    const int y0 = cy * c, y1 = min(y0 + c, ny - 1);
    const int z0 = cz * c, z1 = min(z0 + c, nz - 1);
    bool any = false;
    for (int z = z0; z <= z1 && !any; z++)
        for (int y = y0; y <= y1 && !any; y++) {
            const uint8_t* row = density + ((size_t)z * ny + y) * nx;
            for (int x = x0; x <= x1; x++)
                if (row[x]) {
                    any = true;
                    break;
                }
        }
    if (any) atomicOr(bits + (cell >> 5), 1u << (cell & 31));
}

cudaError_t launchOccupancy(const uint8_t* density, int nx, int ny, int nz, int shift, int ocx, int ocy, int ocz, uint32_t* bits,
                            cudaStream_t st)
{
    const int cells = ocx * ocy * ocz;
    k_occupancy<<<(cells + 127) / 128, 128, 0, st>>>(density, nx, ny, nz, shift, ocx, ocy, ocz, bits);
    return cudaGetLastError();
}

/* ---- number of non-zero voxels on the six faces of the grid ---- */
__global__ void __launch_bounds__(256) k_border_count(const uint8_t* __restrict__ density, int nx, int ny, int nz, uint32_t* count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nxy = (long long)nx * ny, nyz = (long long)ny * nz, nxz = (long long)nx * nz;
    bool nonzero = false;
    if (i < nxy) {
        const int x = (int)(i % nx), y = (int)(i / nx);
        nonzero = density[((size_t)0 * ny + y) * nx + x] || density[((size_t)(nz - 1) * ny + y) * nx + x];
    } else if (i < nxy + nyz) {
        const long long j = i - nxy;
        const int y = (int)(j % ny), z = (int)(j / ny);
        nonzero = density[((size_t)z * ny + y) * nx + 0] || density[((size_t)z * ny + y) * nx + (nx - 1)];
    } else if (i < nxy + nyz + nxz) {
        const long long j = i - nxy - nyz;
        const int x = (int)(j % nx), z = (int)(j / nx);
        nonzero = density[((size_t)z * ny + 0) * nx + x] || density[((size_t)z * ny + (ny - 1)) * nx + x];
    }
    if (nonzero) atomicAdd(count, 1u);
}

cudaError_t launchBorderCount(const uint8_t* density, int nx, int ny, int nz, uint32_t* count, cudaStream_t st)
{
    const long long total = (long long)nx * ny + (long long)ny * nz + (long long)nx * nz;
    k_border_count<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(density, nx, ny, nz, count);
    return cudaGetLastError();
}

/* ---- Chebyshev distance transform over occupancy cells (iterated 26-neighbour relaxation) ---- */
__global__ void __launch_bounds__(128) k_cell_dist_init(const uint32_t* __restrict__ bits, int cells, uint8_t* dist)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    dist[cell] = ((bits[cell >> 5] >> (cell & 31)) & 1u) ? 0 : 255;
}

__global__ void __launch_bounds__(128) k_cell_dist_relax(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int ocx, int ocy, int ocz)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ocx * ocy * ocz) return;
    const int cx = cell % ocx, cy = (cell / ocx) % ocy, cz = cell / (ocx * ocy);
    int best = in[cell];
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int x = cx + dx, y = cy + dy, z = cz + dz;
                if (x < 0 || y < 0 || z < 0 || x >= ocx || y >= ocy || z >= ocz) continue;
                best = min(best, min(254, (int)in[(z * ocy + y) * ocx + x]) + 1);
            }
    out[cell] = (uint8_t)best;
}

cudaError_t launchCellDistance(const uint32_t* occBits, int ocx, int ocy, int ocz, uint8_t* dist, uint8_t* tmp, cudaStream_t st)
{
    const int cells = ocx * ocy * ocz;
    const int blocks = (cells + 127) / 128;
    k_cell_dist_init<<<blocks, 128, 0, st>>>(occBits, cells, dist);
    int iters = max(max(ocx, ocy), ocz);
    iters += iters & 1; /* even number of ping-pong passes: the result ends up in `dist` */
    uint8_t *a = dist, *b = tmp;
    for (int i = 0; i < iters; i++) {
        k_cell_dist_relax<<<blocks, 128, 0, st>>>(a, b, ocx, ocy, ocz);
        uint8_t* t = a;
        a = b;
        b = t;
    }
    return cudaGetLastError();
}

/* ---- CU/progressive.cu:17-27 applied for subframes first..first+n-1 in order, per pixel ---- */
__device__ __forceinline__ void welford(float& mean, float& m2, float x, float w)
{
    const float previousMean = mean;
    const float newMean = previousMean + (x - previousMean) * w;
    mean = newMean;
    m2 = m2 + (x - previousMean) * (x - newMean);
}

__global__ void __launch_bounds__(256) k_update_frame(const float4* __restrict__ staging, const uint32_t* __restrict__ entrySteps,
                                                      float4* __restrict__ progressive, float4* __restrict__ variance, size_t pixels,
                                                      uint32_t firstSubframe, uint32_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    float4 mean = progressive[i], m2 = variance[i];
    /* pixels whose camera ray never reaches the cloud are not traced: their sample is (0, 0, 0, 1) */
    const bool missing = entrySteps != nullptr && entrySteps[i] == ENTRY_MISS;
    for (uint32_t k = 0; k < n; k++) {
        const float4 x = missing ? make_float4(0.f, 0.f, 0.f, 1.f) : staging[(size_t)k * pixels + i];
        const float w = 1.0f / (float)(firstSubframe + k);
        welford(mean.x, m2.x, x.x, w);
        welford(mean.y, m2.y, x.y, w);
        welford(mean.z, m2.z, x.z, w);
        welford(mean.w, m2.w, x.w, w);
    }
    progressive[i] = mean;
    variance[i] = m2;
}

cudaError_t launchUpdateFrame(const float4* staging, const uint32_t* entrySteps, float4* progressive, float4* variance, size_t pixels,
                              uint32_t firstSubframe, uint32_t n, cudaStream_t st)
{
    k_update_frame<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(staging, entrySteps, progressive, variance, pixels, firstSubframe, n);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fill_missing(float4* __restrict__ staging, const uint32_t* __restrict__ entrySteps, size_t pixels)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pixels && entrySteps[i] == ENTRY_MISS) staging[i] = make_float4(0.f, 0.f, 0.f, 1.f);
}

cudaError_t launchFillMissing(float4* staging, const uint32_t* entrySteps, size_t pixels, cudaStream_t st)
{
    k_fill_missing<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(staging, entrySteps, pixels);
    return cudaGetLastError();
}

/* ---- CU/reinhard.cu:20-83 ---- */
__device__ __forceinline__ float luminance(float4 c)
{
    return c.x * 0.265068f + c.y * 0.67023428f + c.z * 0.06409157f + c.w * 0.0f;
}

/* firstPass: one thread per column, serial over rows (keeps the reference's summation order) */
__global__ void __launch_bounds__(128) k_reinhard_columns(const float4* __restrict__ progressive, int w, int h, float* columns)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    float sum = 0;
    for (int y = 0; y < h; y++) sum += luminance(progressive[(size_t)y * w + x]) + 0.00001f;
    columns[x] = sum;
}

/* secondPass */
__global__ void k_reinhard_average(const float* columns, int w, unsigned totalPixels, float* average)
{
    float result = 0;
    for (int i = 0; i < w; i++) result += columns[i];
    average[0] = result / (float)totalPixels;
}

/* applyReinhard.  lw == 0 gives ld / lw = 0/0 = NaN in the reference (reinhard.cu:69); optix::clamp is
 * fmaxf(a, fminf(f, b)) and fminf(NaN, 1) = 1, so such pixels (the empty background) come out WHITE there.  Kept: the
 * reference's own sources compiled on the host (oracle/_ref) show it, and tests/test_oracle_vs_ref.py pins it. */
__global__ void __launch_bounds__(256) k_reinhard_apply(const float4* __restrict__ progressive, size_t pixels, float exposure,
                                                        const float* average, uchar4* __restrict__ screen)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const float4 c = progressive[i];
    const float lw = luminance(c);
    float ld = lw * exposure / average[0];
    ld = ld / (1.f + ld);
    const float k = ld / lw;
    /* nvcc folds fmaxf(0, fminf(x, 1)) into FMUL.SAT, which maps NaN to 0; the IEEE reading of the reference's text maps it
     * to 1 (what oracle/_ref computes), so the NaN case is spelled out */
    const float vr = c.x * k, vg = c.y * k, vb = c.z * k;
    const float r = powf(vr != vr ? 1.f : fmaxf(0.f, fminf(vr, 1.f)), 1.f / 2.2f) * 255;
    const float g = powf(vg != vg ? 1.f : fmaxf(0.f, fminf(vg, 1.f)), 1.f / 2.2f) * 255;
    const float b = powf(vb != vb ? 1.f : fmaxf(0.f, fminf(vb, 1.f)), 1.f / 2.2f) * 255;
    screen[i] = make_uchar4((unsigned char)r, (unsigned char)g, (unsigned char)b, 255);
}

cudaError_t launchTonemap(const float4* progressive, int w, int h, float exposure, float* columns, float* average, uchar4* screen,
                          cudaStream_t st)
{
    k_reinhard_columns<<<(w + 127) / 128, 128, 0, st>>>(progressive, w, h, columns);
    k_reinhard_average<<<1, 1, 0, st>>>(columns, w, (unsigned)w * (unsigned)h, average);
    const size_t pixels = (size_t)w * h;
    k_reinhard_apply<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(progressive, pixels, exposure, average, screen);
    return cudaGetLastError();
}

/* ---- DG/Scene/Cameras/Camera.cpp:232-268 ---- */
__global__ void __launch_bounds__(256) k_unconverged(const float4* __restrict__ progressive, const float4* __restrict__ variance, size_t pixels,
                                                     uint32_t subframeId, uint32_t* count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < pixels) {
        const float N = (float)subframeId;
        const float sigma = sqrtf(variance[i].x / N);
        const float absoluteConfidence = 1.96f * sigma / sqrtf(N);
        const float relativeConfidence = absoluteConfidence / (progressive[i].x + 1.1920929e-07f);
        bad = !(relativeConfidence < 0.02f || absoluteConfidence < 1e-2f);
    }
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (uint32_t)__popc(m));
}

cudaError_t launchUnconverged(const float4* progressive, const float4* variance, size_t pixels, uint32_t subframeId, uint32_t* count,
                              cudaStream_t st)
{
    k_unconverged<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(progressive, variance, pixels, subframeId, count);
    return cudaGetLastError();
}

/* ---- multi-GPU merge: per channel {n*mean, M2 + n*mean^2} in double, summable across ranks ---- */
__global__ void __launch_bounds__(256) k_export_moments(const float4* __restrict__ progressive, const float4* __restrict__ variance,
                                                        size_t pixels, uint32_t n, double* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const float4 m = progressive[i], v = variance[i];
    const double dn = (double)n;
    const double mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        out[8 * i + c] = dn * mm[c];
        out[8 * i + 4 + c] = vv[c] + dn * mm[c] * mm[c];
    }
}

__global__ void __launch_bounds__(256) k_import_moments(const double* __restrict__ in, size_t pixels, uint32_t nTotal,
                                                        float4* __restrict__ progressive, float4* __restrict__ variance)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const double dn = (double)nTotal;
    float mean[4], m2[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const double mu = in[8 * i + c] / dn;
        const double s2 = in[8 * i + 4 + c] - dn * mu * mu;
        mean[c] = (float)mu;
        m2[c] = (float)(s2 > 0.0 ? s2 : 0.0);
    }
    progressive[i] = make_float4(mean[0], mean[1], mean[2], mean[3]);
    variance[i] = make_float4(m2[0], m2[1], m2[2], m2[3]);
}

cudaError_t launchExportMoments(const float4* progressive, const float4* variance, size_t pixels, uint32_t n, double* out, cudaStream_t st)
{
    k_export_moments<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(progressive, variance, pixels, n, out);
    return cudaGetLastError();
}

cudaError_t launchImportMoments(const double* in, size_t pixels, uint32_t nTotal, float4* progressive, float4* variance, cudaStream_t st)
{
    k_import_moments<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(in, pixels, nTotal, progressive, variance);
    return cudaGetLastError();
}

/* the descriptor gather: EXACT instantiation with the software LOD fetch lives in this translation unit (ds_kernels.cuh), the FAST one
 * (texture units, fast math) in ds_kernels_fast.cu */
cudaError_t launchDescriptors(const DevScene& sc, const LevelTable& lv, const DescriptorLayers& layers, const float* pos, const float* dir,
                              uint32_t n, uint8_t* outU8, float* outF32, int32_t* tapIndex, cudaStream_t st, int layerStride, const float* angle,
                              const uint8_t* active, const uint32_t* gather, cudaTextureObject_t mipTex)
{
    if (mipTex)
        return KernelSet<true>::descriptors(sc, lv, layers, pos, dir, n, outU8, outF32, tapIndex, st, layerStride, angle, active, gather, mipTex);
    return KernelSet<false>::descriptors(sc, lv, layers, pos, dir, n, outU8, outF32, tapIndex, st, layerStride, angle, active, gather, 0);
}

/* ---- CU/PointRadianceTask.h:38-49 applied to the `launches` experiments of each thread, in order ---- */
__global__ void __launch_bounds__(256) k_task_welford(DsPointRadianceTask* tasks, const float* __restrict__ x, uint32_t nThreads, uint32_t launches)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nThreads) return;
    uint32_t count = tasks[t].experiment_count;
    float radiance = tasks[t].radiance, var = tasks[t].running_variance;
    for (uint32_t l = 0; l < launches; l++) {
        const float newRadiance = x[(size_t)t * launches + l];
        count++;
        const float N = (float)count;
        const float newWeight = (float)(1.0 / (double)N);
        const float previousMean = radiance;
        const float newMean = radiance + (newRadiance - previousMean) * newWeight;
        radiance = newMean;
        var += (newRadiance - previousMean) * (newRadiance - newMean);
    }
    tasks[t].experiment_count = count;
    tasks[t].radiance = radiance;
    tasks[t].running_variance = var;
}

cudaError_t launchTaskWelford(DsPointRadianceTask* tasks, const float* x, uint32_t nThreads, uint32_t launches, cudaStream_t st)
{
    if (nThreads == 0) return cudaSuccess;
    k_task_welford<<<(nThreads + 255) / 256, 256, 0, st>>>(tasks, x, nThreads, launches);
    return cudaGetLastError();
}

} // namespace dsk
