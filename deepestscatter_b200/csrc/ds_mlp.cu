/*
 * ds_mlp.cu -- the radiance-predicting network of the neural renderer on sm_100a (see ds_mlp.h).
 *
 *   k_disney_mlp_tc   FAST flavour: tcgen05.mma kind::tf32 (or kind::f16 on bf16 operands), M = 128 rows per CTA, N = 208, fp32
 *                     accumulators and the residual in tensor memory, weight and descriptor chunks streamed by 1-D bulk TMA,
 *                     activations never leave the SM;
 *   k_disney_mlp_f32  EXACT flavour: plain fp32 FMA in a fixed summation order (register-tiled, shared-memory staged).
 *
 * Both are checked against oracle/ds_oracle_mlp.cpp, which is pinned to the reference's own DisneyModel.py through
 * tests/golden/disney_mlp.json.
 */
#include <cstring>

#include "ds_mlp.h"
#include "ds_kernels.h"

namespace dsk {

/* ------------------------------------------------------------------------------------------------ host packing */

namespace {

struct BlockPtrs {
    const float *f1zW, *f1zB, *f1oW, *f1oB, *f2W, *f2B;
};

/* tf32 = fp32 with 10 mantissa bits; round to nearest, ties away (what cvt.rna.tf32.f32 does) */
float roundTf32(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x1000u) & 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

constexpr uint32_t B_LBO = MLP_NPAD / 8 * 128; /* bytes between two K groups (16 bytes of K each) of the weights: all 26 row groups */

/* bf16: fp32 truncated to 7 mantissa bits, round to nearest even (__float2bfloat16_rn) */
uint16_t roundBf16(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float bf16ToFloat(uint16_t h)
{
    const uint32_t u = (uint32_t)h << 16;
    float x;
    memcpy(&x, &u, 4);
    return x;
}
/* IEEE half: round to nearest even, subnormals kept, saturating at the largest finite value (what cvt.rn.satfinite.f16.f32 does) */
uint16_t roundF16(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
    const uint32_t a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return (uint16_t)(sign | 0x7fffu); /* NaN */
    if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu); /* >= 65520 rounds beyond the largest half: saturate at 65504 */
    if (a < 0x33000001u) return sign;                        /* <= 2^-25: rounds to zero */
    if (a < 0x38800000u) {                                   /* subnormal half: units of 2^-24 */
        const int shift = 126 - (int)(a >> 23); /* 14..24 */
        const uint32_t mant = (a & 0x7fffffu) | 0x800000u;
        uint32_t h = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1u))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((a - 0x38000000u) >> 13);
    const uint32_t rem = a & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}
float f16ToFloat(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    float x;
    uint32_t u;
    if (e == 0) {
        x = (float)m * 5.9604644775390625e-08f; /* 2^-24 */
        memcpy(&u, &x, 4);
        u |= sign;
    } else if (e == 31) {
        u = sign | 0x7f800000u | (m << 13);
    } else {
        u = sign | ((e + 112u) << 23) | (m << 13);
    }
    memcpy(&x, &u, 4);
    return x;
}
/* operand types of the tensor-core kernel */
enum { OPS_TF32 = 0, OPS_BF16 = 1, OPS_F16 = 2 };
uint16_t roundHalf(float x, int ops) { return ops == OPS_F16 ? roundF16(x) : roundBf16(x); }
float halfToFloat(uint16_t h, int ops) { return ops == OPS_F16 ? f16ToFloat(h) : bf16ToFloat(h); }

/* The tensor-core kernel exists for two operand types.  G = K values per 16-byte K group: 4 (tf32, kept in 4-byte words) or 8 (bf16).  One MMA
 * step consumes two K groups (K = 8 or 16), a chunk is up to four steps (K = 32 or 64), so chunk and stage sizes in bytes are the same. */

/* one K chunk of W [200][ld] (torch Linear layout), k in [k0, k0 + 2 * G * steps), zero beyond row 200, in the UMMA canonical K-major no-swizzle
 * layout: 8-row x 16-byte core matrices, row groups 128 B apart, K groups B_LBO apart.  Columns kValid and kValid + 1 (when bias != NULL)
 * carry the bias split into a rounded value and the rounded remainder: the A operand holds 1.0 there, so the MMA adds the bias at nearly
 * fp32 precision */
void appendWeightChunk(std::vector<uint8_t>& stream, const float* W, int ld, int k0, int steps, int kValid, const float* bias, int ops)
{
    const bool bf16 = ops != OPS_TF32; /* 16-bit operands: bfloat16 or IEEE half */
    const int G = bf16 ? 8 : 4, E = bf16 ? 2 : 4;
    const size_t base = stream.size();
    const int kc = 2 * G * steps;
    stream.resize(base + (size_t)(kc / G) * B_LBO, 0);
    for (int kk = 0; kk < kc; ++kk)
        for (int n = 0; n < MLP_D; ++n) {
            const int k = k0 + kk;
            float v = 0.0f;
            if (k < kValid)
                v = W[(size_t)n * ld + k];
            else if (bias && k == kValid)
                v = bias[n];
            else if (bias && k == kValid + 1)
                v = bias[n] - (bf16 ? halfToFloat(roundHalf(bias[n], ops), ops) : roundTf32(bias[n]));
            const size_t off = base + (size_t)(kk / G) * B_LBO + (size_t)(n / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(kk % G) * E;
            if (bf16) {
                const uint16_t h = roundHalf(v, ops);
                memcpy(&stream[off], &h, 2);
            } else {
                const float r = roundTf32(v);
                memcpy(&stream[off], &r, 4);
            }
        }
}

/* the chunks of one GEMM operand of K values (+ 2 bias columns when bias != NULL; padded to a whole number of MMA steps): four steps at a
 * time, then the rest */
void appendGemmPart(std::vector<uint8_t>& stream, std::vector<MlpChunk>& chunks, int ops, const float* W, int ld, int K, const float* bias, uint8_t src,
                    uint8_t layer, uint8_t dst, uint8_t gemm, bool first, bool waitAct, bool last, uint8_t epilogue)
{
    const int G = ops != OPS_TF32 ? 8 : 4, stepK = 2 * G, chunkK = 4 * stepK;
    const int kPad = (K + (bias ? 2 : 0) + stepK - 1) / stepK * stepK;
    for (int k0 = 0; k0 < kPad; k0 += chunkK) {
        const int steps = (kPad - k0 >= chunkK ? chunkK : kPad - k0) / stepK;
        MlpChunk c{};
        c.wOffset = (uint32_t)stream.size();
        appendWeightChunk(stream, W, ld, k0, steps, K, bias, ops);
        c.wBytes = (uint32_t)stream.size() - c.wOffset;
        c.k8 = (uint16_t)steps;
        c.aKGroup = (uint16_t)(src == 0 ? k0 / G : k0);
        c.src = src;
        c.layer = layer;
        c.dst = dst;
        c.gemm = gemm;
        c.flags = 0;
        if (k0 == 0 && first) c.flags |= MLP_FIRST;
        if (k0 == 0 && waitAct) c.flags |= MLP_WAIT_ACT;
        if (k0 + chunkK >= kPad && last) {
            c.flags |= MLP_LAST;
            c.epilogue = epilogue;
        }
        chunks.push_back(c);
    }
}

} // namespace

void packDisneyModel(const float* w, DisneyModelHost& out)
{
    BlockPtrs blk[MLP_NB];
    for (int i = 0; i < MLP_NB; ++i) {
        blk[i].f1zW = w;
        blk[i].f1zB = blk[i].f1zW + MLP_D * MLP_ZD;
        blk[i].f1oW = blk[i].f1zB + MLP_D;
        blk[i].f1oB = blk[i].f1oW + MLP_D * MLP_D;
        blk[i].f2W = blk[i].f1oB + MLP_D;
        blk[i].f2B = blk[i].f2W + MLP_D * MLP_D;
        w = blk[i].f2B + MLP_D;
    }
    const float* fc0W = w;
    const float* fc0B = fc0W + MLP_D * MLP_D;
    const float* fc2W = fc0B + MLP_D;
    const float* fc2B = fc2W + MLP_D * MLP_D;
    const float* fc4W = fc2B + MLP_D;
    const float* fc4B = fc4W + MLP_D;

    /* fp32 kernel: W^T per GEMM, [K][200] */
    out.wT.clear();
    out.bias.assign((size_t)MLP_GEMMS * MLP_NPAD, 0.0f);
    auto appendT = [&](const float* W, int K) {
        const size_t base = out.wT.size();
        out.wT.resize(base + (size_t)K * MLP_D);
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < MLP_D; ++c) out.wT[base + (size_t)k * MLP_D + c] = W[(size_t)c * K + k];
    };
    for (int i = 0; i < MLP_NB; ++i) {
        appendT(blk[i].f1oW, MLP_D);
        appendT(blk[i].f1zW, MLP_ZD);
        appendT(blk[i].f2W, MLP_D);
        for (int c = 0; c < MLP_D; ++c) {
            out.bias[(size_t)(2 * i) * MLP_NPAD + c] = blk[i].f1oB[c] + blk[i].f1zB[c];
            out.bias[(size_t)(2 * i + 1) * MLP_NPAD + c] = blk[i].f2B[c];
        }
    }
    appendT(fc0W, MLP_D);
    appendT(fc2W, MLP_D);
    for (int c = 0; c < MLP_D; ++c) {
        out.bias[(size_t)20 * MLP_NPAD + c] = fc0B[c];
        out.bias[(size_t)21 * MLP_NPAD + c] = fc2B[c];
    }
    out.w4b4.assign(MLP_NPAD + 1, 0.0f); /* 200 weights, zero padding to 208, bias */
    for (int c = 0; c < MLP_D; ++c) out.w4b4[c] = fc4W[c];
    out.w4b4[MLP_NPAD] = fc4B[0];

    /* tensor-core kernel: the program, once per operand type.  Biases ride in the GEMMs: the descriptor layer is staged with z[226] = z[227] = 1
     * and the activation buffer holds 1 in columns 200 and 201 */
    for (int t = 0; t < 3; ++t) {
        const int bf16 = t; /* OPS_TF32, OPS_BF16, OPS_F16; the two 16-bit types share one chunk table */
        std::vector<MlpChunk> chunksF16;
        std::vector<uint8_t>& stream = t == OPS_F16 ? out.streamF16 : t == OPS_BF16 ? out.streamBf16 : out.stream;
        std::vector<MlpChunk>& chunks = t == OPS_F16 ? chunksF16 : t == OPS_BF16 ? out.chunksBf16 : out.chunks;
        stream.clear();
        chunks.clear();
        for (int i = 0; i < MLP_NB; ++i) {
            /* h = relu(z_i . f1z^T + o . f1o^T + b).  The descriptor part goes first: it does not depend on the previous block's epilogue, so
             * its MMAs run while the workers are still writing o; o = 0 in block 0 (DisneyModel.py:34), whose f1o part is skipped */
            appendGemmPart(stream, chunks, bf16, blk[i].f1zW, MLP_ZD, MLP_ZD, &out.bias[(size_t)(2 * i) * MLP_NPAD], 1, (uint8_t)i, 0, (uint8_t)(2 * i), true,
                           false, i == 0, MLP_EPI_H);
            if (i > 0) appendGemmPart(stream, chunks, bf16, blk[i].f1oW, MLP_D, MLP_D, nullptr, 0, 0, 0, (uint8_t)(2 * i), false, true, true, MLP_EPI_H);
            /* o = relu(h . f2^T + b + o): D2 still holds o, the MMAs accumulate on top of it */
            appendGemmPart(stream, chunks, bf16, blk[i].f2W, MLP_D, MLP_D, blk[i].f2B, 0, 0, 1, (uint8_t)(2 * i + 1), i == 0, true, true, MLP_EPI_O);
        }
        appendGemmPart(stream, chunks, bf16, fc0W, MLP_D, MLP_D, fc0B, 0, 0, 0, 20, true, true, true, MLP_EPI_H);
        /* into D2 (the residual is no longer needed): its first MMAs start while the epilogue of the layer before is still reading D1 */
        appendGemmPart(stream, chunks, bf16, fc2W, MLP_D, MLP_D, fc2B, 0, 0, 1, 21, true, true, true, MLP_EPI_OUT);
    }
}

/* ------------------------------------------------------------------------------------------------ fp32 kernel */

constexpr int F32_TM = 64;       /* rows per CTA */
constexpr int F32_KC = 32;       /* K rows of W^T staged at a time */
constexpr int F32_THREADS = 200; /* 8 row groups x 25 column groups, 8 x 8 outputs per thread */
constexpr size_t F32_SMEM = (size_t)(2 * MLP_D * F32_TM + MLP_ZD * F32_TM + F32_KC * MLP_D) * sizeof(float);

/* acc[i][j] += sum_k A[k][ty*8 + i] * W^T[k][tx*8 + j]; A = A0 (K0 rows) followed by A1 (K1 rows), both [k][64] in shared memory */
__device__ __forceinline__ void gemmF32(const float* A0, int K0, const float* A1, int K1, const float* __restrict__ Wg, float* wS, int ty, int tx,
                                        float (&acc)[8][8])
{
    const int K = K0 + K1;
    for (int k0 = 0; k0 < K; k0 += F32_KC) {
        const int kc = min(F32_KC, K - k0);
        __syncthreads(); /* the previous chunk has been consumed */
        const float4* src = reinterpret_cast<const float4*>(Wg + (size_t)k0 * MLP_D);
        for (int i = threadIdx.x; i < kc * (MLP_D / 4); i += F32_THREADS) reinterpret_cast<float4*>(wS)[i] = __ldg(src + i);
        __syncthreads();
        for (int kk = 0; kk < kc; ++kk) {
            const int k = k0 + kk;
            const float* arow = k < K0 ? A0 + k * F32_TM : A1 + (k - K0) * F32_TM;
            const float4 a0 = *reinterpret_cast<const float4*>(arow + ty * 8), a1 = *reinterpret_cast<const float4*>(arow + ty * 8 + 4);
            const float4 w0 = *reinterpret_cast<const float4*>(wS + kk * MLP_D + tx * 8), w1 = *reinterpret_cast<const float4*>(wS + kk * MLP_D + tx * 8 + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], wv[j], acc[i][j]);
        }
    }
}

/* dst[c][r] = relu(acc + bias[c] (+ res[c][r])) for the thread's 8 x 8 outputs; dst may alias res (same element, same thread) */
__device__ __forceinline__ void epilogueF32(const float (&acc)[8][8], const float* __restrict__ bias, const float* res, float* dst, int ty, int tx)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = tx * 8 + j;
        const float b = __ldg(bias + c);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = acc[i][j] + b;
            if (res) x += res[c * F32_TM + ty * 8 + i];
            v[i] = fmaxf(x, 0.0f);
        }
        *reinterpret_cast<float4*>(dst + c * F32_TM + ty * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + c * F32_TM + ty * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__global__ void __launch_bounds__(F32_THREADS, 1)
    k_disney_mlp_f32(const float* __restrict__ in, const uint32_t* __restrict__ rowIndex, uint32_t nRows, const float* __restrict__ wT,
                     const float* __restrict__ bias, const float* __restrict__ w4b4, float* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float* actO = reinterpret_cast<float*>(smemRaw); /* [200][64]: element (k, row) */
    float* actH = actO + MLP_D * F32_TM;
    float* actZ = actH + MLP_D * F32_TM; /* [226][64] */
    float* wS = actZ + MLP_ZD * F32_TM;  /* [32][200] */
    const int ty = threadIdx.x / 25, tx = threadIdx.x % 25;
    const uint32_t row0 = blockIdx.x * F32_TM;
    for (int i = threadIdx.x; i < MLP_D * F32_TM; i += F32_THREADS) actO[i] = 0.0f; /* DisneyModel.py:34 */

    const float* wg = wT;
    float acc[8][8];
    for (int blk = 0; blk < MLP_NB; ++blk) {
        __syncthreads(); /* actZ of the previous block has been consumed */
        /* descriptor layer blk of the tile's rows, transposed to [k][row] (lane <-> row: conflict-free stores) */
        for (int i = threadIdx.x; i < MLP_ZD * F32_TM; i += F32_THREADS) {
            const int k = i / F32_TM, r = i % F32_TM;
            float v = 0.0f;
            if (row0 + r < nRows) {
                const size_t src = rowIndex ? rowIndex[row0 + r] : row0 + r;
                v = __ldg(in + (src * MLP_NB + blk) * MLP_ZD + k);
            }
            actZ[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
        gemmF32(actO, MLP_D, actZ, MLP_ZD, wg, wS, ty, tx, acc); /* first syncs, so actZ is complete */
        wg += (size_t)(MLP_D + MLP_ZD) * MLP_D;
        epilogueF32(acc, bias + (size_t)(2 * blk) * MLP_NPAD, nullptr, actH, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
        gemmF32(actH, MLP_D, nullptr, 0, wg, wS, ty, tx, acc);
        wg += (size_t)MLP_D * MLP_D;
        epilogueF32(acc, bias + (size_t)(2 * blk + 1) * MLP_NPAD, actO, actO, ty, tx);
    }
    /* fullyConnected */
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    gemmF32(actO, MLP_D, nullptr, 0, wg, wS, ty, tx, acc);
    wg += (size_t)MLP_D * MLP_D;
    epilogueF32(acc, bias + (size_t)20 * MLP_NPAD, nullptr, actH, ty, tx);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    gemmF32(actH, MLP_D, nullptr, 0, wg, wS, ty, tx, acc);
    __syncthreads(); /* every thread is done reading actO's predecessor chain before it is overwritten */
    epilogueF32(acc, bias + (size_t)21 * MLP_NPAD, nullptr, actO, ty, tx);
    __syncthreads();
    if (threadIdx.x < F32_TM && row0 + threadIdx.x < nRows) {
        float y = __ldg(w4b4 + MLP_NPAD);
        for (int k = 0; k < MLP_D; ++k) y = fmaf(actO[k * F32_TM + threadIdx.x], __ldg(w4b4 + k), y);
        const size_t dst = rowIndex ? rowIndex[row0 + threadIdx.x] : row0 + threadIdx.x;
        out[dst] = y > 0.0f ? y : 0.01f * y; /* torch.nn.LeakyReLU default slope */
    }
}

cudaError_t launchDisneyMlpF32(const DisneyModelDev& m, const float* in, const uint32_t* rowIndex, uint32_t nRows, float* out, cudaStream_t st)
{
    if (nRows == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_disney_mlp_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F32_SMEM);
    if (e != cudaSuccess) return e;
    k_disney_mlp_f32<<<(nRows + F32_TM - 1) / F32_TM, F32_THREADS, F32_SMEM, st>>>(in, rowIndex, nRows, m.wT, m.bias, m.w4b4, out);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------ tcgen05 kernel */

constexpr int TC_M = 128;          /* rows per CTA = UMMA M = TMEM lanes */
constexpr int TC_WORKERS = 256;    /* warps 0-7 run the epilogues: threads t and t + 128 share row t (TMEM lane t) and split its columns (a warp
                                      reaches the TMEM lanes of quadrant warp % 4) */
constexpr int TC_ISSUER_WARP = TC_WORKERS / 32;     /* TMEM allocation; walks the program and issues the MMAs (one elected lane) */
constexpr int TC_PRODUCER_WARP = TC_ISSUER_WARP + 1; /* operand streams: bulk copies of weight and descriptor chunks (one lane) */
constexpr int TC_THREADS = TC_WORKERS + 64;
constexpr int TC_PIECES = 7;       /* epilogue pieces = K chunks of the next GEMM, one hand-off barrier each: 7 of 32 columns (the last one 16) for
                                      tf32 operands, 4 of 64 (16) for bf16; columns [0, 128) belong to the first half of the workers */
constexpr int TC_WSTAGES = 3, TC_ZSTAGES = 2;
constexpr int TC_MAX_CHUNKS = 232;
/* the program every warp follows, as a kernel parameter: it sits in the constant bank, so a chunk's fields are uniform-register operands for
 * the warp that issues the MMAs (no per-thread registers, no shared-memory copy) */
struct MlpProgram {
    int nChunks;
    MlpChunk chunks[TC_MAX_CHUNKS];
};
constexpr uint32_t TC_A_LBO = TC_M / 8 * 128;           /* 2048: bytes between 4-float K groups of the activations (all 16 row groups) */
constexpr uint32_t TC_SBO = 128;                        /* bytes between 8-row groups */
constexpr uint32_t TC_ACT_KGROUPS = MLP_NPAD / 4;       /* 200 activations, the two constant-one columns, padding to 208 (tf32; bf16 needs half) */
constexpr uint32_t TC_ACT_BYTES = TC_ACT_KGROUPS * TC_A_LBO; /* 106496 */
constexpr uint32_t TC_WSTAGE_BYTES = 8 * B_LBO;    /* 26624: a chunk is 8 K groups of 16 bytes: K = 32 (tf32) or 64 (bf16) */
constexpr uint32_t TC_ZSTAGE_BYTES = 8 * TC_A_LBO; /* 16384 */
constexpr uint32_t TC_OFF_W = TC_ACT_BYTES;
constexpr uint32_t TC_OFF_Z = TC_OFF_W + TC_WSTAGES * TC_WSTAGE_BYTES;
constexpr uint32_t TC_OFF_BAR = TC_OFF_Z + TC_ZSTAGES * TC_ZSTAGE_BYTES;
constexpr uint32_t TC_SMEM = TC_OFF_BAR + 256;
static_assert(sizeof(MlpChunk) == 20, "MlpChunk layout (include/ds_abi.h documents it)");
static_assert(TC_SMEM <= 232448, "shared memory budget");
static_assert(((TC_ACT_BYTES + TC_WSTAGES * TC_WSTAGE_BYTES + TC_ZSTAGES * TC_ZSTAGE_BYTES) >> 4) + 3 * (2 * 3328 >> 4) < 0x4000, "descriptor start field");
constexpr uint32_t TC_TMEM_COLS = 512, TC_D2_COL = 256;
/* instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
 * both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28 */
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(MLP_NPAD >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
/* A = B = BF16 (format 1), kind::f16 */
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MLP_NPAD >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
/* A = B = F16 (format 0), kind::f16 */
constexpr uint32_t TC_IDESC_F16 = (1u << 4) | ((uint32_t)(MLP_NPAD >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

/* barrier slots (8 bytes each) */
enum { BAR_WFULL = 0, BAR_WFREE = BAR_WFULL + TC_WSTAGES, BAR_ZFULL = BAR_WFREE + TC_WSTAGES, BAR_ZFREE = BAR_ZFULL + TC_ZSTAGES,
       BAR_GEMM = BAR_ZFREE + TC_ZSTAGES, BAR_ACT /* one per epilogue piece */, BAR_COUNT = BAR_ACT + TC_PIECES };
static_assert(BAR_COUNT * 8 + 8 <= 256, "barrier area");

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbarArrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbarTry(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
/* bounded wait: a protocol error must end the kernel with an error code, never hang the GPU */
template <bool PROFILE>
__device__ __forceinline__ bool mbarWait(uint32_t bar, uint32_t parity, volatile uint32_t* abortFlag, long long& waited)
{
    if (!PROFILE && mbarTry(bar, parity)) return true;
    const long long t0 = clock64(); /* try_wait itself may block for a while before it answers */
    for (;;) {
        if (mbarTry(bar, parity)) {
            if (PROFILE) waited += clock64() - t0;
            return true;
        }
        if (*abortFlag) return false;
        if (clock64() - t0 > 2000000000ll) { /* ~1 s */
            *abortFlag = 1u;
            return false;
        }
    }
}

/* shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): low word = start >> 4 at bits 0-13 and leading byte offset
 * >> 4 at bits 16-29; high word = stride byte offset >> 4 at bits 0-13 (32-45 of the descriptor), version 1 at bit 14 (46); layout type 0
 * (no swizzle) at bits 29-31 (61-63) */
__device__ __forceinline__ uint32_t ummaDescLo(uint32_t saddr, uint32_t lbo) { return ((saddr & 0x3ffffu) >> 4) | ((lbo >> 4) << 16); }
constexpr uint32_t TC_DESC_HI = (TC_SBO >> 4) | (1u << 14);

template <int OPS>
__device__ __forceinline__ void ummaIssue(uint32_t tmemD, uint32_t descALo, uint32_t descBLo, uint32_t accumulate)
{
    if (OPS != OPS_TF32)
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmemD), "r"(descALo), "r"(descBLo), "r"(TC_DESC_HI), "r"(OPS == OPS_F16 ? TC_IDESC_F16 : TC_IDESC_BF16), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmemD), "r"(descALo), "r"(descBLo), "r"(TC_DESC_HI), "r"(TC_IDESC), "r"(accumulate)
            : "memory");
}

/* elect.sync: one lane of the (converged) warp */
__device__ __forceinline__ bool electOne()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void ummaCommit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmemLoad(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmemLoad(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmemStore(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmemStore(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

__device__ __forceinline__ float toTf32(float x)
{
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

/* two fp32 -> two 16-bit operands in one word (bfloat16, or IEEE half saturating at the largest finite value) */
template <int OPS>
__device__ __forceinline__ uint32_t packBf16(float lo, float hi)
{
    uint32_t u;
    if (OPS == OPS_F16)
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi), "f"(lo));
    else
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi), "f"(lo)); /* first source -> upper half */
    return u;
}

/* N accumulator columns of the thread's row, starting at col0 (a multiple of 16).  The bias is already in the accumulator (it rides in the
 * GEMM), so this is relu + the write-back: tf32 as 4-column float4 K groups, bf16 as 8-column K groups */
template <int N, int OPS>
__device__ __forceinline__ void epilogueCols(uint32_t tacc, int col0, int epilogue, unsigned char* actRow, const float* __restrict__ w4, float& y)
{
    uint32_t v[N];
    tmemLoad(tacc + (uint32_t)col0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = __float_as_uint(fmaxf(__uint_as_float(v[q]), 0.0f));
    if (epilogue == MLP_EPI_OUT) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w4 + col0 + 4 * q)); /* zero beyond column 200 */
            y = fmaf(__uint_as_float(v[4 * q]), ww.x, y);
            y = fmaf(__uint_as_float(v[4 * q + 1]), ww.y, y);
            y = fmaf(__uint_as_float(v[4 * q + 2]), ww.z, y);
            y = fmaf(__uint_as_float(v[4 * q + 3]), ww.w, y);
        }
        return;
    }
    if (OPS != OPS_TF32) {
#pragma unroll
        for (int q = 0; q < N / 8; ++q) {
            const int col = col0 + 8 * q;
            if (col < MLP_D)
                *reinterpret_cast<uint4*>(actRow + (uint32_t)(col / 8) * TC_A_LBO) =
                    make_uint4(packBf16<OPS>(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1])), packBf16<OPS>(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                               packBf16<OPS>(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])), packBf16<OPS>(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
        }
    } else {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const int col = col0 + 4 * q;
            if (col < MLP_D)
                *reinterpret_cast<float4*>(actRow + (uint32_t)(col / 4) * TC_A_LBO) =
                    make_float4(toTf32(__uint_as_float(v[4 * q])), toTf32(__uint_as_float(v[4 * q + 1])), toTf32(__uint_as_float(v[4 * q + 2])),
                                toTf32(__uint_as_float(v[4 * q + 3])));
        }
    }
    /* the post-activation value is the residual of the next block: the next MMAs accumulate on top of it */
    if (epilogue == MLP_EPI_O) tmemStore(tacc + (uint32_t)col0, v);
}

template <bool PROFILE, int OPS>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_disney_mlp_tc(const __grid_constant__ MlpProgram prog, const void* __restrict__ tiles, uint32_t nRows, const uint8_t* __restrict__ stream,
                    const float* __restrict__ w4b4, float* __restrict__ out, uint32_t* __restrict__ errorOut, unsigned long long* __restrict__ prof)
{
    constexpr bool BF16 = OPS != OPS_TF32; /* 16-bit operands (bfloat16 or IEEE half): 8 values per K group */
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* actS = smem;
    const long long tStart = clock64();
    const MlpChunk* chunks = prog.chunks;
    const int nChunks = prog.nChunks;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
    volatile uint32_t* abortFlag = tmemSlot + 1;
    const uint32_t barBase = smemAddr(bars);
    const uint32_t actAddr = smemAddr(actS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_WSTAGES; ++s) {
            mbarInit(barBase + 8 * (BAR_WFULL + s), 1);
            mbarInit(barBase + 8 * (BAR_WFREE + s), 1);
        }
        for (int s = 0; s < TC_ZSTAGES; ++s) {
            mbarInit(barBase + 8 * (BAR_ZFULL + s), 1);
            mbarInit(barBase + 8 * (BAR_ZFREE + s), 1);
        }
        mbarInit(barBase + 8 * BAR_GEMM, 1);
        for (int p = 0; p < TC_PIECES; ++p) mbarInit(barBase + 8 * (BAR_ACT + p), TC_WORKERS / 2);
        *abortFlag = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_ISSUER_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemSlot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmemBase = *tmemSlot;

    if (warp == TC_PRODUCER_WARP) {
        /* ===== operand streams: one thread keeps every free weight stage and descriptor stage filled with bulk copies ===== */
        if (lane == 0) {
            long long wFree = 0, zFree = 0;
            bool ok = true;
            /* one descriptor layer of the tile: 58 K groups of 4 floats (tf32) or 30 K groups of 8 bf16, x 128 rows x 16 bytes */
            constexpr uint32_t G = BF16 ? 8u : 4u;
            constexpr uint32_t LAYER_BYTES = (BF16 ? 30u : 58u) * TC_A_LBO;
            const char* tile = reinterpret_cast<const char*>(tiles) + (size_t)blockIdx.x * (MLP_NB * LAYER_BYTES);
            uint32_t zFills = 0;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tile), "r"(LAYER_BYTES) : "memory");
            for (int c = 0; c < nChunks && ok; ++c) {
                const MlpChunk ch = chunks[c];
                const int ws = c % TC_WSTAGES;
                if (c >= TC_WSTAGES) ok = mbarWait<PROFILE>(barBase + 8 * (BAR_WFREE + ws), (uint32_t)(c / TC_WSTAGES - 1) & 1u, abortFlag, wFree);
                if (!ok) break;
                const uint32_t bar = barBase + 8 * (BAR_WFULL + ws);
                mbarExpectTx(bar, ch.wBytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 actAddr + TC_OFF_W + (uint32_t)ws * TC_WSTAGE_BYTES),
                             "l"(stream + ch.wOffset), "r"(ch.wBytes), "r"(bar)
                             : "memory");
                if (ch.src == 1) {
                    /* the chunk's K groups are contiguous in the tile: [layer][K group][row][4] */
                    const int zs = (int)(zFills % TC_ZSTAGES);
                    if (zFills >= TC_ZSTAGES) ok = mbarWait<PROFILE>(barBase + 8 * (BAR_ZFREE + zs), (zFills / TC_ZSTAGES - 1) & 1u, abortFlag, zFree);
                    if (!ok) break;
                    const uint32_t zBytes = (uint32_t)ch.k8 * 2u * TC_A_LBO;
                    const uint32_t zbar = barBase + 8 * (BAR_ZFULL + zs);
                    mbarExpectTx(zbar, zBytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     actAddr + TC_OFF_Z + (uint32_t)zs * TC_ZSTAGE_BYTES),
                                 "l"(tile + (size_t)ch.layer * LAYER_BYTES + (size_t)(ch.aKGroup / G) * TC_A_LBO), "r"(zBytes), "r"(zbar)
                                 : "memory");
                    zFills++;
                    /* the next layer of the tile starts its way into L2 a whole block ahead */
                    if (ch.aKGroup == 0 && ch.layer + 1 < MLP_NB)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tile + (size_t)(ch.layer + 1) * LAYER_BYTES), "r"(LAYER_BYTES) : "memory");
                }
            }
            if (!ok) *abortFlag = 1u;
            if (PROFILE && prof && blockIdx.x == 0) {
                prof[1] = (unsigned long long)wFree; /* producer: waiting for a free weight stage, for a free descriptor stage */
                prof[9] = (unsigned long long)zFree;
            }
        }
    } else if (warp == TC_ISSUER_WARP) {
        /* ===== MMA issuer: the whole warp walks the program (uniform control flow, operands in uniform registers), one elected lane issues ===== */
        {
            uint32_t actWaits = 0, zUses = 0;
            bool ok = true;
            long long wAct = 0, wZ = 0, wW = 0; /* cycles spent waiting, by cause (PROFILE) */
            for (int c = 0; c < nChunks; ++c) {
                const MlpChunk& ch = chunks[c];
                const int ws = c % TC_WSTAGES;
                /* a chunk that reads the activation buffer needs only ITS 32 columns from the epilogue before it: the GEMM starts while the
                 * workers are still writing the later pieces */
                if (ch.flags & MLP_WAIT_ACT) actWaits++;
                if (ch.src == 0) ok = mbarWait<PROFILE>(barBase + 8 * (BAR_ACT + ch.aKGroup / 8), (actWaits - 1u) & 1u, abortFlag, wAct);
                const int zs = (int)(zUses % TC_ZSTAGES);
                if (ok && ch.src == 1) ok = mbarWait<PROFILE>(barBase + 8 * (BAR_ZFULL + zs), (zUses / TC_ZSTAGES) & 1u, abortFlag, wZ);
                if (ok) ok = mbarWait<PROFILE>(barBase + 8 * (BAR_WFULL + ws), (uint32_t)(c / TC_WSTAGES) & 1u, abortFlag, wW);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t aLo = ummaDescLo(ch.src == 1 ? actAddr + TC_OFF_Z + (uint32_t)zs * TC_ZSTAGE_BYTES : actAddr + (uint32_t)ch.aKGroup * TC_A_LBO, TC_A_LBO);
                const uint32_t bLo = ummaDescLo(actAddr + TC_OFF_W + (uint32_t)ws * TC_WSTAGE_BYTES, B_LBO);
                const uint32_t d = tmemBase + (ch.dst ? TC_D2_COL : 0u);
                const int k8 = ch.k8;
                if (electOne()) {
                    /* one MMA per 8 k values: two 4-float K groups of each operand, i.e. 2 * LBO bytes further per step */
                    ummaIssue<OPS>(d, aLo, bLo, (ch.flags & MLP_FIRST) ? 0u : 1u);
                    if (k8 > 1) ummaIssue<OPS>(d, aLo + (2u * TC_A_LBO >> 4), bLo + (2u * B_LBO >> 4), 1u);
                    if (k8 > 2) ummaIssue<OPS>(d, aLo + 2u * (2u * TC_A_LBO >> 4), bLo + 2u * (2u * B_LBO >> 4), 1u);
                    if (k8 > 3) ummaIssue<OPS>(d, aLo + 3u * (2u * TC_A_LBO >> 4), bLo + 3u * (2u * B_LBO >> 4), 1u);
                    ummaCommit(barBase + 8 * (BAR_WFREE + ws));
                    if (ch.src == 1) ummaCommit(barBase + 8 * (BAR_ZFREE + zs));
                    if (ch.flags & MLP_LAST) ummaCommit(barBase + 8 * BAR_GEMM);
                }
                __syncwarp();
                if (ch.src == 1) zUses++;
            }
            if (!ok) *abortFlag = 1u;
            if (PROFILE && prof && blockIdx.x == 0 && lane == 0) {
                prof[0] = (unsigned long long)(clock64() - tStart); /* issuer: total, then waits for the previous epilogue, a staged */
                prof[2] = (unsigned long long)wAct;                 /* descriptor chunk, a landed weight chunk */
                prof[3] = (unsigned long long)wZ;
                prof[4] = (unsigned long long)wW;
            }
        }
    } else {
        /* ===== workers: the epilogues ===== */
        const int t = threadIdx.x & (TC_M - 1), half = threadIdx.x / TC_M;
        const uint32_t row = blockIdx.x * TC_M + t;
        const bool valid = row < nRows;
        const uint32_t rowOff = (uint32_t)(t / 8) * TC_SBO + (uint32_t)(t % 8) * 16u; /* the row's 16 bytes inside a K group */
        const uint32_t tmemRow = tmemBase + ((uint32_t)((warp & 3) * 32) << 16);
        /* the constant-one columns 200, 201 that carry the biases through the GEMMs; 202..207 are padding */
        if (half == 0) {
            if (BF16) {
                *reinterpret_cast<uint4*>(actS + (uint32_t)(MLP_D / 8) * TC_A_LBO + rowOff) =
                    make_uint4(OPS == OPS_F16 ? 0x3c003c00u : 0x3f803f80u, 0u, 0u, 0u); /* 1, 1, 0 x 6 in the operand type */
            } else {
                *reinterpret_cast<float4*>(actS + (uint32_t)(MLP_D / 4) * TC_A_LBO + rowOff) = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
                *reinterpret_cast<float4*>(actS + (uint32_t)(MLP_D / 4 + 1) * TC_A_LBO + rowOff) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        uint32_t gemmWaits = 0;
        bool ok = true;
        long long wGemm = 0, tEpi = 0;
        for (int c = 0; c < nChunks && ok; ++c) {
            const MlpChunk ch = chunks[c];
            if (ch.flags & MLP_LAST) {
                const bool mine = mbarWait<PROFILE>(barBase + 8 * BAR_GEMM, gemmWaits & 1u, abortFlag, wGemm);
                const long long te0 = PROFILE ? clock64() : 0;
                gemmWaits++;
                ok = __all_sync(0xffffffffu, mine);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmemRow + (ch.dst ? TC_D2_COL : 0u);
                const int epilogue = ch.epilogue;
                float y = 0.0f;
                /* columns [0, 128) belong to the first half of the workers, [128, 208) to the second; a piece (= K chunk of the next GEMM: 32
                 * columns for tf32 operands, 64 for bf16) is handed over as soon as it is complete */
                constexpr int PIECE_COLS = BF16 ? 64 : 32;
                const int colEnd = half ? MLP_NPAD : 128;
#pragma unroll 1
                for (int col0 = half ? 128 : 0; col0 < colEnd; col0 += 32) {
                    if (col0 + 32 <= MLP_NPAD)
                        epilogueCols<32, OPS>(tacc, col0, epilogue, actS + rowOff, w4b4, y);
                    else
                        epilogueCols<16, OPS>(tacc, col0, epilogue, actS + rowOff, w4b4, y);
                    const int done = min(col0 + 32, MLP_NPAD);
                    if (epilogue != MLP_EPI_OUT && (done % PIECE_COLS == 0 || done == MLP_NPAD)) {
                        /* hand the piece over: TMEM store done, shared-memory stores visible to the tensor core's async proxy */
                        if (epilogue == MLP_EPI_O) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbarArrive(barBase + 8 * (BAR_ACT + (done - 1) / PIECE_COLS));
                    }
                }
                if (epilogue == MLP_EPI_OUT) {
                    /* the two halves of a row meet in shared memory (the descriptor stages are idle by now) */
                    float* partial = reinterpret_cast<float*>(smem + TC_OFF_Z);
                    if (half == 1) partial[t] = y;
                    asm volatile("bar.sync 1, %0;" ::"n"(TC_WORKERS) : "memory");
                    if (half == 0) {
                        y += partial[t] + __ldg(w4b4 + MLP_NPAD);
                        if (valid) out[row] = y > 0.0f ? y : 0.01f * y; /* torch.nn.LeakyReLU default slope */
                    }
                }
                if (PROFILE) tEpi += clock64() - te0;
            }
        }
        if (!ok) *abortFlag = 1u;
        if (PROFILE && prof && blockIdx.x == 0 && threadIdx.x == 0) {
            prof[8] = (unsigned long long)(clock64() - tStart); /* worker 0: total, waiting for a GEMM, inside the epilogues */
            prof[10] = (unsigned long long)wGemm;
            prof[11] = (unsigned long long)tEpi;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && *abortFlag) atomicExch(errorOut, 1u + blockIdx.x);
    if (warp == TC_ISSUER_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(TC_TMEM_COLS) : "memory");
    }
}

template <bool PROFILE, int OPS>
static cudaError_t launchTc(const MlpProgram& prog, const void* tiles, uint32_t nRows, const uint8_t* stream, const float* w4b4, float* out, uint32_t* error,
                            unsigned long long* prof, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_disney_mlp_tc<PROFILE, OPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
    if (e != cudaSuccess) return e;
    k_disney_mlp_tc<PROFILE, OPS><<<(nRows + TC_M - 1) / TC_M, TC_THREADS, TC_SMEM, st>>>(prog, tiles, nRows, stream, w4b4, out, error, prof);
    return cudaGetLastError();
}

cudaError_t launchDisneyMlpTc(const DisneyModelDev& m, const void* tiles, uint32_t nRows, float* out, cudaStream_t st, unsigned long long* prof, int ops)
{
    if (nRows == 0) return cudaSuccess;
    const MlpProgram* prog = ops != OPS_TF32 ? m.programBf16 : m.program;
    const uint8_t* stream = ops == OPS_F16 ? m.streamF16 : ops == OPS_BF16 ? m.streamBf16 : m.stream;
    if (!prog || !stream) return cudaErrorInvalidValue;
    if (ops == OPS_F16) return prof ? launchTc<true, OPS_F16>(*prog, tiles, nRows, stream, m.w4b4, out, m.error, prof, st)
                                    : launchTc<false, OPS_F16>(*prog, tiles, nRows, stream, m.w4b4, out, m.error, nullptr, st);
    if (ops == OPS_BF16) return prof ? launchTc<true, OPS_BF16>(*prog, tiles, nRows, stream, m.w4b4, out, m.error, prof, st)
                                     : launchTc<false, OPS_BF16>(*prog, tiles, nRows, stream, m.w4b4, out, m.error, nullptr, st);
    return prof ? launchTc<true, OPS_TF32>(*prog, tiles, nRows, stream, m.w4b4, out, m.error, prof, st)
                : launchTc<false, OPS_TF32>(*prog, tiles, nRows, stream, m.w4b4, out, m.error, nullptr, st);
}

/* the chunk table as the kernel-parameter block (host memory, owned by the model) */
MlpProgram* makeMlpProgram(const std::vector<MlpChunk>& chunks)
{
    if ((int)chunks.size() > TC_MAX_CHUNKS) return nullptr;
    MlpProgram* p = new MlpProgram();
    memset(p, 0, sizeof(*p));
    p->nChunks = (int)chunks.size();
    memcpy(p->chunks, chunks.data(), chunks.size() * sizeof(MlpChunk));
    return p;
}
void freeMlpProgram(MlpProgram* p) { delete p; }

/* DisneyNetworkInput rows [n][10][226] -> the tiles the tensor-core kernel consumes (rounded to the operand type; k 226, 227 = 1; the rest of
 * the layer's K padding and the padding rows zero): one thread per (row, layer, K group) */
template <int OPS>
__global__ void __launch_bounds__(256) k_network_input_to_tiles(const float* __restrict__ in, uint32_t nRows, uint32_t nPadded, void* __restrict__ tiles)
{
    constexpr bool BF16 = OPS != OPS_TF32;
    constexpr uint32_t G = BF16 ? 8u : 4u, KG = BF16 ? 30u : 58u; /* K values per group, groups per layer */
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)nPadded * (MLP_NB * KG)) return;
    const uint32_t row = (uint32_t)(g / (MLP_NB * KG));
    const uint32_t lk = (uint32_t)(g - (size_t)row * (MLP_NB * KG)); /* layer * KG + K group */
    const uint32_t layer = lk / KG, kg = lk % KG;
    float v[G];
#pragma unroll
    for (uint32_t e = 0; e < G; ++e) {
        const uint32_t k = kg * G + e;
        v[e] = 0.0f;
        if (row < nRows) {
            if (k < (uint32_t)MLP_ZD)
                v[e] = __ldg(in + ((size_t)row * MLP_NB + layer) * MLP_ZD + k);
            else if (k < (uint32_t)MLP_ZD + 2)
                v[e] = 1.0f;
        }
    }
    unsigned char* dst = reinterpret_cast<unsigned char*>(tiles) + ((size_t)(row >> 7) * (MLP_NB * KG) + lk) * TC_A_LBO + (size_t)(row & 127u) * 16;
    if (BF16)
        *reinterpret_cast<uint4*>(dst) = make_uint4(packBf16<OPS>(v[0], v[1]), packBf16<OPS>(v[2], v[3]), packBf16<OPS>(v[G - 4], v[G - 3]), packBf16<OPS>(v[G - 2], v[G - 1]));
    else
        *reinterpret_cast<float4*>(dst) = make_float4(toTf32(v[0]), toTf32(v[1]), toTf32(v[2]), toTf32(v[3]));
}

cudaError_t launchNetworkInputToTiles(const float* in, uint32_t nRows, void* tiles, cudaStream_t st, int ops)
{
    if (nRows == 0) return cudaSuccess;
    const uint32_t nPadded = (nRows + 127u) / 128u * 128u;
    const size_t total = (size_t)nPadded * MLP_NB * (ops != OPS_TF32 ? 30 : 58);
    if (ops == OPS_F16)
        k_network_input_to_tiles<OPS_F16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, nRows, nPadded, tiles);
    else if (ops == OPS_BF16)
        k_network_input_to_tiles<OPS_BF16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, nRows, nPadded, tiles);
    else
        k_network_input_to_tiles<OPS_TF32><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, nRows, nPadded, tiles);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------ renderRect glue */

/* compaction of the rows with active[i] != 0: idx[0..*count) (count zeroed by the caller).  One warp-aggregated atomic per warp:
 * the order of idx depends on scheduling, the set does not (and the model's rows are independent of each other) */
__global__ void __launch_bounds__(256) k_compact_active(const uint8_t* __restrict__ active, uint32_t n, uint32_t* __restrict__ idx,
                                                        uint32_t* __restrict__ count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool a = i < n && active[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, a);
    if (m == 0u) return;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (a) idx[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = i;
}

cudaError_t launchCompactActive(const uint8_t* active, uint32_t n, uint32_t* idx, uint32_t* count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    k_compact_active<<<(n + 255) / 256, 256, 0, st>>>(active, n, idx, count);
    return cudaGetLastError();
}

/* test hook (option compact_reverse): the same set of rows in the opposite order -- results must not depend on where a row lands */
__global__ void k_reverse_u32(uint32_t* a, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n / 2) {
        const uint32_t x = a[i], y = a[n - 1 - i];
        a[i] = y;
        a[n - 1 - i] = x;
    }
}
cudaError_t launchReverse(uint32_t* a, uint32_t n, cudaStream_t st)
{
    if (n < 2) return cudaSuccess;
    k_reverse_u32<<<(n / 2 + 255) / 256, 256, 0, st>>>(a, n);
    return cudaGetLastError();
}

/* CU/disneyCamera.cu:38-46: frameResult[pixel] = (make_float4(predicted) + make_float4(radiance)) * (1 - transmittance) for the n compacted
 * rows; row i is frame pixel idx[i], info is indexed by the frame pixel */
__global__ void k_blit_predicted(const float* __restrict__ predicted, const float* __restrict__ info, const uint32_t* __restrict__ idx, uint32_t n,
                                 float4* __restrict__ frameResult)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pixel = idx[i];
    const float* f = info + (size_t)pixel * 5; /* DsIntersectionInfo: radiance rgb, transmittance, hasScattered */
    const float w = 1.0f - f[3], p = predicted[i];
    frameResult[pixel] = make_float4((p + f[0]) * w, (p + f[1]) * w, (p + f[2]) * w, (p + 0.0f) * w);
}

cudaError_t launchBlitPredicted(const float* predicted, const float* info, const uint32_t* idx, uint32_t n, float4* frameResult, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    k_blit_predicted<<<(n + 255) / 256, 256, 0, st>>>(predicted, info, idx, n, frameResult);
    return cudaGetLastError();
}

} // namespace dsk
