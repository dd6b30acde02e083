// Tensor-core (tcgen05 / TMEM) path of the CSPNet decoder -- placeholder until the kernels land.
#include "cb2_internal.cuh"

namespace cb2 {

int tc_forward_layers(const cb2_model *, const cb2_batch *, const cb2_forward_io *, ForwardWs &, cudaStream_t) {
  return fail(CB2_ERR_UNSUPPORTED, "tensor-core path not built yet");
}

}  // namespace cb2
