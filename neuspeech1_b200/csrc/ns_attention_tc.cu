// Tensor-core attention (placeholder dispatch until the tcgen05 flash kernels land: reports "unsupported" so that the
// caller takes the SIMT kernels -- still a CUDA path, never a CPU one).
#include "ns_common.cuh"

namespace ns {
int attention_fwd_tc(const ns_attn_shape&, const void*, const void*, const void*, void*, float*, cudaStream_t) {
  return NS_ERR_UNSUPPORTED;
}
int attention_bwd_tc(const ns_attn_shape&, const void*, const void*, const void*, const void*, const void*, const float*,
                     float*, void*, void*, void*, cudaStream_t) {
  return NS_ERR_UNSUPPORTED;
}
}  // namespace ns
