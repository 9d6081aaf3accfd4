/*
 * ds_oracle_mlp.cpp -- TEST INFRASTRUCTURE.  CPU restatement of the reference's radiance-predicting network
 * (DeepestScatter_Train/Disney/DisneyModel.py, DisneyBlock.py), the model DisneyRenderer::renderRect evaluates on the
 * network inputs of a rectangle (DG/Scene/Cameras/DisneyRenderer.cpp:104).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg load this; the product never does.
 *
 * PARITY PINNED: tests/golden/disney_mlp.json holds outputs of the reference's own DisneyModel.py (torch, CPU), made by
 * tools/make_golden_disney_mlp.py; tests/test_disney_mlp.py checks this restatement against them.
 *
 * Weights: the flat state_dict layout documented in include/ds_abi.h (ds_disney_model_load).  Accumulation is in double
 * so that the oracle sits between any two float32 summation orders.
 */
#include <cstddef>
#include <cstdint>
#include <vector>

namespace {

constexpr int D = 200;  /* DisneyModel.BLOCK_DIMENSION (DisneyModel.py:6) */
constexpr int NB = 10;  /* DisneyModel.BLOCK_COUNT (:7) */
constexpr int ZD = 226; /* DESCRIPTOR_LAYER_WITH_ANGLE_DIMENSION (:8-9) */

/* y = W x + b, torch.nn.Linear: W is [out][in] row-major */
void linear(const float* W, const float* b, const double* x, int in, int out, double* y)
{
    for (int o = 0; o < out; ++o) {
        double acc = (double)b[o];
        const float* w = W + (size_t)o * in;
        for (int i = 0; i < in; ++i) acc += (double)w[i] * x[i];
        y[o] = acc;
    }
}

} // namespace

extern "C" {

size_t orc_disney_weight_count() { return (size_t)NB * (D * ZD + D + 2 * (D * D + D)) + 2 * (D * D + D) + D + 1; }

/* input: [n][10][226]; out: [n]; hidden (optional): [n][200], the activations entering fullyConnected */
void orc_disney_forward(const float* weights, const float* input, int n, float* out, float* hidden)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
        const float* w = weights;
        std::vector<double> o(D, 0.0), z(ZD), a(D), b(D), t(D);
        /* DisneyModel.__blocksForward (DisneyModel.py:31-37): out = zeros; out = block(out, zLayers[:, i]) */
        for (int i = 0; i < NB; ++i) {
            const float* f1zW = w;
            const float* f1zB = f1zW + D * ZD;
            const float* f1oW = f1zB + D;
            const float* f1oB = f1oW + D * D;
            const float* f2W = f1oB + D;
            const float* f2B = f2W + D * D;
            w = f2B + D;
            for (int k = 0; k < ZD; ++k) z[k] = (double)input[((size_t)r * NB + i) * ZD + k];
            /* DisneyBlock.forward (DisneyBlock.py:19-33): out = relu(f1o(o) + f1z(z)); out = relu(f2(out) + o) */
            linear(f1oW, f1oB, o.data(), D, D, a.data());
            linear(f1zW, f1zB, z.data(), ZD, D, b.data());
            for (int k = 0; k < D; ++k) {
                const double v = a[k] + b[k];
                a[k] = v > 0.0 ? v : 0.0;
            }
            linear(f2W, f2B, a.data(), D, D, t.data());
            for (int k = 0; k < D; ++k) {
                const double v = t[k] + o[k];
                o[k] = v > 0.0 ? v : 0.0;
            }
        }
        if (hidden)
            for (int k = 0; k < D; ++k) hidden[(size_t)r * D + k] = (float)o[k];
        /* fullyConnected (DisneyModel.py:52-59): Linear, ReLU, Linear, ReLU, Linear(200 -> 1), LeakyReLU(0.01) */
        const float* W0 = w;
        const float* B0 = W0 + D * D;
        const float* W2 = B0 + D;
        const float* B2 = W2 + D * D;
        const float* W4 = B2 + D;
        const float* B4 = W4 + D;
        linear(W0, B0, o.data(), D, D, a.data());
        for (int k = 0; k < D; ++k) a[k] = a[k] > 0.0 ? a[k] : 0.0;
        linear(W2, B2, a.data(), D, D, t.data());
        for (int k = 0; k < D; ++k) t[k] = t[k] > 0.0 ? t[k] : 0.0;
        double y;
        linear(W4, B4, t.data(), D, 1, &y);
        out[r] = (float)(y > 0.0 ? y : 0.01 * y);
    }
}

} /* extern "C" */
