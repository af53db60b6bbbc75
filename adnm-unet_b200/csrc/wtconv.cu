#include "adn_common.cuh"
extern "C" {
int wtconv_workspace_bytes(const WtShape*, size_t*, size_t*, size_t*) { adn::set_error("wtconv: not built yet"); return ADN_ERR_ARCH; }
int wtconv_forward(const WtShape*, const WtWeights*, const void*, void*, void*, void*, void*) { adn::set_error("wtconv: not built yet"); return ADN_ERR_ARCH; }
int wtconv_backward(const WtShape*, const WtWeights*, const void*, const void*, const void*, void*, const WtWeightGrads*, void*, void*) { adn::set_error("wtconv: not built yet"); return ADN_ERR_ARCH; }
}
