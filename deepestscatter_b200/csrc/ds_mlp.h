/*
 * ds_mlp.h -- the radiance-predicting network of the neural renderer (DeepestScatter_Train/Disney/DisneyModel.py,
 * DisneyBlock.py; evaluated by DisneyRenderer::renderRect, DG/Scene/Cameras/DisneyRenderer.cpp:104) as sm_100a kernels.
 * Internal to the library.
 *
 * The network is a chain of 22 GEMMs over the rows (pixels) of a batch, all of width 200:
 *   block i (x10):  h = relu([o | z_i] . [f1o.W | f1z.W]^T + f1o.b + f1z.b)     K = 200 + 226      (DisneyBlock.py:25-26)
 *                   o = relu(h . f2.W^T + f2.b + o)                              K = 200            (:28-30)
 *   fullyConnected: relu(Linear), relu(Linear), leaky_relu(Linear 200 -> 1)                          (DisneyModel.py:52-59)
 * Rows never interact, so a CTA carries a tile of rows through the whole chain with the activations on chip.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace dsk {

constexpr int MLP_D = 200;      /* DisneyModel.BLOCK_DIMENSION */
constexpr int MLP_NB = 10;      /* DisneyModel.BLOCK_COUNT */
constexpr int MLP_ZD = 226;     /* DESCRIPTOR_LAYER_WITH_ANGLE_DIMENSION */
constexpr int MLP_NPAD = 208;   /* output width padded to a multiple of 16 (UMMA N for M = 128) */
constexpr int MLP_GEMMS = 22;   /* 2 per block + fullyConnected.0 + fullyConnected.2 */
constexpr size_t MLP_WEIGHT_COUNT = (size_t)MLP_NB * (MLP_D * MLP_ZD + MLP_D + 2 * (MLP_D * MLP_D + MLP_D)) + 2 * (MLP_D * MLP_D + MLP_D) + MLP_D + 1;

/* one step of the tensor-core kernel's program: a K-chunk of one GEMM */
struct MlpChunk {
    uint32_t wOffset;  /* byte offset of the chunk's weights in the packed stream */
    uint32_t wBytes;   /* (k8 * 2) * MLP_NPAD * 16 */
    uint16_t k8;       /* number of MMA steps in the chunk (1..4): K = 8 each for tf32 operands, 16 for bf16 */
    uint16_t aKGroup;  /* src 0: first 16-byte K group of the activation buffer; src 1: first k of the descriptor layer */
    uint8_t src;       /* 0 = activation buffer, 1 = descriptor layer (z) */
    uint8_t layer;     /* src 1: descriptor layer index */
    uint8_t dst;       /* accumulator: 0 = D1 (h), 1 = D2 (o, carries the residual) */
    uint8_t flags;     /* MLP_FIRST | MLP_LAST | MLP_WAIT_ACT */
    uint8_t epilogue;  /* MLP_LAST: which epilogue follows */
    uint8_t gemm;      /* bias row */
    uint8_t pad[2];
};
enum { MLP_FIRST = 1 /* first MMA overwrites the accumulator */, MLP_LAST = 2 /* last chunk of its GEMM */,
       MLP_WAIT_ACT = 4 /* first chunk that reads what the previous epilogue wrote */ };
enum { MLP_EPI_H = 1 /* relu -> activation buffer */, MLP_EPI_O = 2 /* relu -> activation buffer and back into D2 */,
       MLP_EPI_OUT = 3 /* relu, dot with fullyConnected.4, leaky relu -> out */ };

struct DisneyModelDev {
    float* wT = nullptr;       /* fp32 kernel: per GEMM W^T [K][200] (K = 426 for the first GEMM of a block: o rows, then z rows) */
    float* bias = nullptr;     /* [22][208]; the two biases of a block's first GEMM are pre-summed */
    float* w4b4 = nullptr;     /* fullyConnected.4: 200 weights, zero padding to 208, bias at [208] */
    uint8_t* stream = nullptr; /* tensor-core kernel: weights in UMMA canonical K-major layout, in consumption order (tf32 operands) */
    uint8_t* streamBf16 = nullptr; /* the same for bf16 operands (option mlp_bf16) */
    uint8_t* streamF16 = nullptr;  /* ... and for IEEE half operands (option mlp_fp16); same chunk table as bf16 */
    struct MlpProgram* programBf16 = nullptr;
    struct MlpProgram* program = nullptr; /* host: the chunk table, passed to the kernel as its (grid-constant) parameter block */
    int nChunks = 0;
    uint32_t* error = nullptr; /* device word: non-zero if a barrier wait of the tensor-core kernel timed out */
    unsigned long long* prof = nullptr; /* 16 words: cycle accounting of block 0 of the last tensor-core launch (profile_events) */
    bool loaded = false;
};

/* host-side packing of the flat state_dict array (include/ds_abi.h: ds_disney_model_load) */
struct DisneyModelHost {
    std::vector<float> wT, bias, w4b4;
    std::vector<uint8_t> stream, streamBf16, streamF16; /* tf32 / bf16 / IEEE half operands */
    std::vector<MlpChunk> chunks, chunksBf16;
};
void packDisneyModel(const float* weights, DisneyModelHost& out);
struct MlpProgram* makeMlpProgram(const std::vector<MlpChunk>& chunks); /* NULL if the table is too long */
void freeMlpProgram(struct MlpProgram* p);

/* fp32 kernel: in = [nRowsTotal][10][226] floats on the device; rowIndex (may be NULL): the nRows input rows to evaluate (gather);
 * out[rowIndex[i]] (or out[i]) receives the prediction */
cudaError_t launchDisneyMlpF32(const DisneyModelDev& m, const float* in, const uint32_t* rowIndex, uint32_t nRows, float* out, cudaStream_t st);
/* tensor-core kernel: the rows come as 128-row tiles in the layout its MMAs read straight from shared memory (ds_kernels.h
 * NETWORK_TILE_FLOATS: [layer][K group][row][4], so that the K chunk of a layer is one contiguous block a bulk copy can fetch); out[i] for
 * i < nRows.  prof (may be NULL): 16 device words; block 0 leaves its cycle accounting there */
cudaError_t launchDisneyMlpTc(const DisneyModelDev& m, const void* tiles, uint32_t nRows, float* out, cudaStream_t st, unsigned long long* prof = nullptr,
                              int ops = 0 /* operand type: 0 = tf32, 1 = bfloat16, 2 = IEEE half */);
/* [nRows][10][226] -> ceil(nRows / 128) tiles (networkTileBytes per tile) */
cudaError_t launchNetworkInputToTiles(const float* in, uint32_t nRows, void* tiles, cudaStream_t st, int ops = 0);
/* bytes of one 128-row tile of network input: tf32 operands [10][58][128][4 floats], 16-bit operands [10][30][128][8] */
inline size_t networkTileBytesOf(int ops) { return (size_t)10 * (ops ? 30 : 58) * 128 * 16; }
/* indices of the rows with active[i] != 0: idx[0..*count), unordered */
cudaError_t launchCompactActive(const uint8_t* active, uint32_t n, uint32_t* idx, uint32_t* count, cudaStream_t st);
cudaError_t launchReverse(uint32_t* a, uint32_t n, cudaStream_t st); /* test hook: reverse the compacted order */
/* copyToFrameResult (CU/disneyCamera.cu:38-46) on the device for n compacted rows: row i is frame pixel idx[i] */
cudaError_t launchBlitPredicted(const float* predicted, const float* info, const uint32_t* idx, uint32_t n, float4* frameResult, cudaStream_t st);

} // namespace dsk
