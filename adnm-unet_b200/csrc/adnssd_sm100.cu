#include "adnssd_sm100.cuh"
namespace adn {
bool sm100_supported(const MixerDims&) { return false; }
void sm100_workspace_bytes(const MixerDims&, size_t* f, size_t* b) { *f = 0; *b = 0; }
int sm100_forward(const MixerDims&, const AdnWeights&, const bf16*, bf16*, void*, void*, cudaStream_t) {
  set_error("sm100 path not built"); return ADN_ERR_ARCH; }
int sm100_backward(const MixerDims&, const AdnWeights&, const bf16*, const void*, const bf16*, bf16*,
                   const AdnWeightGrads&, void*, cudaStream_t) { set_error("sm100 path not built"); return ADN_ERR_ARCH; }
}
