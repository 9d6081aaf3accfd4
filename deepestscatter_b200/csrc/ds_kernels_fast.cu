/*
 * ds_kernels_fast.cu -- translation unit compiled with FMA contraction: the throughput flavour of the
 * estimator kernels (hardware 3-D texture filtering, MUFU exp/log/sincos), i.e. the arithmetic the
 * reference itself runs (rtTex3D + --use_fast_math).  Validated against the oracle statistically.
 */
#include "ds_kernels.cuh"

namespace dsk {

template struct KernelSet<true>;

} // namespace dsk
