// add_layer_norm instantiations for __nv_bfloat16 tensors (see layernorm.cu)
#include "layernorm_impl.cuh"

namespace oeh {
int launch_ln_bf16(const void* a, const void* b, const void* gamma, const void* beta, void* y, int64_t rows, int64_t cols, float eps,
                  const float* q_a, const float* q_res, const float* q_out, int n_bits, void* sum_out, cudaStream_t st,
                  const float* gate_w, const float* gate_b, float* gate_out, int T_len, int H) {
  return launch<__nv_bfloat16>(a, b, gamma, beta, y, rows, cols, eps, q_a, q_res, q_out, n_bits, sum_out, st, gate_w, gate_b, gate_out, T_len, H);
}
}  // namespace oeh
