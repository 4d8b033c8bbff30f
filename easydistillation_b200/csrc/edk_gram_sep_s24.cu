// Kernels of the separable contraction for the mode structure max |px|, |py| = 2, max px^2 + py^2 = 4
// (edk_gram_sep.cuh; one translation unit per structure so that they compile in parallel).
#include "edk_gram_sep.cuh"

namespace edk {
#ifndef EDK_EMU_NO_LAUNCHERS
cudaError_t launch_gram_sep_s24(const SepParams& P, const SepTma& T, int pairs, int bytes, unsigned items, cudaStream_t s) {
    return launch_gram_sep_q<2, 4>(P, T, pairs, bytes, items, s);
}
cudaError_t launch_gram_sepx_s24(const SepParams& P, const SepTmaX& T, const SepWeights& W, int pairs, int shape, int bytes, unsigned items,
                                 cudaStream_t s) {
    return launch_gram_sepx_q<2, 4>(P, T, W, pairs, shape, bytes, items, s);
}
#endif
}  // namespace edk
