// Tensor-mode fused causal attention (tcgen05 / TMEM / TMA).  Placeholder until the
// kernel lands: bf16 attention is served by the CUDA-core kernel in attn_simt.cu.
#include "common.cuh"

namespace dgpt {
bool attn_tc_supported(const dgpt_attn_args*) { return false; }
int launch_attn_fwd_tc(const dgpt_attn_args*, cudaStream_t) {
  set_error("attn_fwd(tensor): not available");
  return DGPT_E_ARG;
}
int launch_attn_bwd_tc(const dgpt_attn_args*, cudaStream_t) {
  set_error("attn_bwd(tensor): not available");
  return DGPT_E_ARG;
}
}  // namespace dgpt
