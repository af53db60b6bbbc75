// sm_100a tensor-core (tcgen05 / TMEM) path of the ADN-SSD mixer, bf16 I/O.  Declarations only.
#pragma once
#include "adn_common.cuh"

namespace adn {
bool sm100_supported(const MixerDims& d);
bool sm100_rowconv(const MixerDims& d);      // subset served by the conv-as-GEMM row kernels
void sm100_workspace_bytes(const MixerDims& d, size_t* fwd, size_t* bwd);
size_t sm100_saved_extra_bytes(const MixerDims& d);
int sm100_forward(const MixerDims& d, const AdnWeights& w, const bf16* u, bf16* out, void* saved, void* ws,
                  cudaStream_t st);
int sm100_backward(const MixerDims& d, const AdnWeights& w, const bf16* u, const void* saved, const bf16* dout,
                   bf16* du, const AdnWeightGrads& g, void* ws, cudaStream_t st);
}  // namespace adn
