// C entry points of the LayerNorm / embedding kernels (templates: layernorm_impl.cuh).
#include "layernorm_impl.cuh"

// add_layer_norm for one tensor dtype each: layernorm_f32.cu / layernorm_f16.cu / layernorm_bf16.cu (three translation units, so
// the (dtype x row width x gate x taps) instantiations compile in parallel: one file took 110 s of a clean build)
namespace oeh {
#define OEH_LN_ARGS                                                                                                          \
  const void *a, const void *b, const void *gamma, const void *beta, void *y, int64_t rows, int64_t cols, float eps,          \
      const float *q_a, const float *q_res, const float *q_out, int n_bits, void *sum_out, cudaStream_t st, const float *gate_w, \
      const float *gate_b, float *gate_out, int T_len, int H
int launch_ln_f32(OEH_LN_ARGS);
int launch_ln_f16(OEH_LN_ARGS);
int launch_ln_bf16(OEH_LN_ARGS);
}  // namespace oeh

using namespace oeh;

extern "C" int oeh_add_layer_norm_gate(const void* a, const void* b, const void* weight, const void* bias, void* y,
                                       int64_t rows, int64_t cols, double eps, int32_t dtype, const float* qparams_a,
                                       const float* qparams_res, const float* qparams_out, int32_t n_bits, void* sum_out,
                                       const float* gate_w, const float* gate_b, float* gate_out, int32_t T, int32_t H,
                                       oeh_stream_t stream);

extern "C" int oeh_add_layer_norm(const void* a, const void* b, const void* weight, const void* bias, void* y, int64_t rows,
                                  int64_t cols, double eps, int32_t dtype, const float* qparams_a,
                                  const float* qparams_res, const float* qparams_out, int32_t n_bits, void* sum_out,
                                  oeh_stream_t stream) {
  return oeh_add_layer_norm_gate(a, b, weight, bias, y, rows, cols, eps, dtype, qparams_a, qparams_res, qparams_out, n_bits,
                                 sum_out, nullptr, nullptr, nullptr, 1, 0, stream);
}

extern "C" int oeh_add_layer_norm_gate(const void* a, const void* b, const void* weight, const void* bias, void* y,
                                       int64_t rows, int64_t cols, double eps, int32_t dtype, const float* qparams_a,
                                       const float* qparams_res, const float* qparams_out, int32_t n_bits, void* sum_out,
                                       const float* gate_w, const float* gate_b, float* gate_out, int32_t T, int32_t H,
                                       oeh_stream_t stream) {
  if (gate_out) {
    OEH_CHECK_ARG(gate_w && gate_b, "oeh_add_layer_norm_gate: null gate weights");
    OEH_CHECK_ARG((reinterpret_cast<uintptr_t>(gate_w) & 15) == 0, "oeh_add_layer_norm_gate: gate weights must be 16-byte aligned");
    OEH_CHECK_ARG(T > 0 && H > 0 && (int64_t)H * 64 == cols && rows % T == 0,
                  "oeh_add_layer_norm_gate: the gate needs cols == 64 H and rows == B T (cols=%lld H=%d rows=%lld T=%d)",
                  (long long)cols, H, (long long)rows, T);
  }
  OEH_CHECK_ARG(rows >= 0 && cols > 0, "oeh_add_layer_norm: bad shape (%lld, %lld)", (long long)rows, (long long)cols);
  if (rows == 0) return OEH_OK;
  OEH_CHECK_ARG(a && weight && y, "oeh_add_layer_norm: null pointer");
  OEH_CHECK_ARG(rows <= 0x7fffffffLL * 8, "oeh_add_layer_norm: too many rows");
  if (qparams_a || qparams_res || qparams_out) OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_add_layer_norm: n_bits=%d outside [1,16]", n_bits);
  if (!(qparams_a || qparams_res || qparams_out)) n_bits = 8;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32:
      return launch_ln_f32(a, b, weight, bias, y, rows, cols, (float)eps, qparams_a, qparams_res, qparams_out, n_bits, sum_out, st,
                           gate_w, gate_b, gate_out, T, H);
    case OEH_F16:
      return launch_ln_f16(a, b, weight, bias, y, rows, cols, (float)eps, qparams_a, qparams_res, qparams_out, n_bits, sum_out, st,
                            gate_w, gate_b, gate_out, T, H);
    case OEH_BF16:
      return launch_ln_bf16(a, b, weight, bias, y, rows, cols, (float)eps, qparams_a, qparams_res, qparams_out, n_bits,
                                   sum_out, st, gate_w, gate_b, gate_out, T, H);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_add_layer_norm: unknown dtype %d", dtype);
}

extern "C" int oeh_embed_sum(const int64_t* ids, const int64_t* pos_ids, const void* word, const void* pos_table, void* y,
                             int64_t rows, int64_t cols, int32_t T, int64_t vocab, int64_t n_pos, int64_t pos_offset,
                             int32_t dtype, oeh_stream_t stream) {
  OEH_CHECK_ARG(rows >= 0 && cols > 0 && T > 0 && vocab > 0 && n_pos > 0, "oeh_embed_sum: bad shape");
  if (rows == 0) return OEH_OK;
  OEH_CHECK_ARG(ids && word && pos_table && y, "oeh_embed_sum: null pointer");
  const int vec = dtype == OEH_F32 ? 4 : 8;
  if (cols % vec != 0) return set_error(OEH_ERR_UNSUPPORTED, "oeh_embed_sum: cols=%lld must be a multiple of %d", (long long)cols, vec);
  if ((reinterpret_cast<uintptr_t>(word) | reinterpret_cast<uintptr_t>(pos_table) | reinterpret_cast<uintptr_t>(y)) & 15)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_embed_sum: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  switch (dtype) {
    case OEH_F32:
      embed_sum_kernel<float><<<grid, kWarpsPerBlock * 32, 0, st>>>(ids, pos_ids, (const float*)word, (const float*)pos_table,
                                                                    (float*)y, rows, (int)cols, T, vocab, n_pos, pos_offset);
      break;
    case OEH_F16:
      embed_sum_kernel<__half><<<grid, kWarpsPerBlock * 32, 0, st>>>(ids, pos_ids, (const __half*)word, (const __half*)pos_table,
                                                                     (__half*)y, rows, (int)cols, T, vocab, n_pos, pos_offset);
      break;
    case OEH_BF16:
      embed_sum_kernel<__nv_bfloat16><<<grid, kWarpsPerBlock * 32, 0, st>>>(ids, pos_ids, (const __nv_bfloat16*)word,
                                                                            (const __nv_bfloat16*)pos_table, (__nv_bfloat16*)y, rows,
                                                                            (int)cols, T, vocab, n_pos, pos_offset);
      break;
    default:
      return set_error(OEH_ERR_BAD_ARG, "oeh_embed_sum: unknown dtype %d", dtype);
  }
  return check_launch("embed_sum_kernel");
}

extern "C" int oeh_embed_layer_norm(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const void* word,
                                    const void* type_table, const void* pos_table, const void* weight, const void* bias, void* y,
                                    int64_t rows, int64_t cols, int32_t T, int64_t vocab, double eps, int32_t dtype,
                                    oeh_stream_t stream) {
  OEH_CHECK_ARG(rows >= 0 && cols > 0 && T > 0 && vocab > 0, "oeh_embed_layer_norm: bad shape");
  if (rows == 0) return OEH_OK;
  OEH_CHECK_ARG(ids && word && weight && y, "oeh_embed_layer_norm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32:
      return launch_embed<float>(ids, type_ids, pos_ids, word, type_table, pos_table, weight, bias, y, rows, cols, T, vocab, (float)eps, st);
    case OEH_F16:
      return launch_embed<__half>(ids, type_ids, pos_ids, word, type_table, pos_table, weight, bias, y, rows, cols, T, vocab, (float)eps, st);
    case OEH_BF16:
      return launch_embed<__nv_bfloat16>(ids, type_ids, pos_ids, word, type_table, pos_table, weight, bias, y, rows, cols, T, vocab,
                                         (float)eps, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_embed_layer_norm: unknown dtype %d", dtype);
}
