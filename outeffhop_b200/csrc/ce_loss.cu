// Mean cross-entropy of one evaluation batch straight from the half-precision logits:
//   loss = mean_{i : label_i != ignore_index} ( logsumexp_j x[i,j] - x[i,label_i] )
// = what HF's BertForMaskedLM / OPTForCausalLM compute with CrossEntropyLoss() on `prediction_scores.view(-1, V)`
// (the per-batch `outputs.loss` of validate_mlm_config.py:489-496 and validate_clm.py:555-561), evaluated in fp32.
//
// HBM-bound: the logits of a BERT-base batch are 4096 x 30522 fp16 = 250 MB, larger than L2, so each row is read
// exactly ONCE (online max / sum per thread, merged per block), never cast to an fp32 copy, and rows whose label is
// ignore_index -- 85 % of an MLM batch -- are not read at all (they do not enter the mean).  The row stride is
// explicit so the caller may keep the vocabulary dimension padded to a multiple of 8 elements (a 30522-wide fp16
// GEMM output has 4-byte aligned rows and falls off cuBLAS' fast path; 30528 does not).
// Two launches, deterministic: stage 1 one block per row, stage 2 one block sums the rows in a fixed order (float64).
#include "oeh_common.cuh"

namespace oeh {
namespace {

constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// running (m, l): sum_j e^(t_j) = l * e^m.  m stays in natural units and (t - m) is formed BEFORE the log2e multiply:
// for 16-bit logits the difference is exact, so the result does not lose digits when |m| is large (a folded
// t*log2e - m*log2e would carry the rounding of m*log2e, ~|m| * 6e-8, into every term with the same sign).
struct ML {
  float m, l;
};
template <int n>
__device__ __forceinline__ void ml_add_block(ML& a, const float (&t)[n]) {
  float mx = t[0];
#pragma unroll
  for (int j = 1; j < n; ++j) mx = fmaxf(mx, t[j]);
  if (mx > a.m) {
    a.l *= ex2f((a.m - mx) * kLog2e); // a.m = -inf on the first block: 2^-inf = 0
    a.m = mx;
  }
  if (a.m == -INFINITY) return;       // every value so far is -inf
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < n; ++j) s += ex2f((t[j] - a.m) * kLog2e);
  a.l += s;
}
__device__ __forceinline__ ML ml_merge(ML a, ML b) {
  const float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return ML{m, 0.f};
  return ML{m, a.l * ex2f((a.m - m) * kLog2e) + b.l * ex2f((b.m - m) * kLog2e)};
}

// Label of logits row r.  shift_T > 0: causal-LM scoring of blocks of shift_T tokens -- row (b, t) is scored against
// labels[b, t + 1] and the last row of every block is ignored (HF's shift_logits / shift_labels, validate_clm.py:555-561)
// without a shifted copy of the labels (a fill + a strided copy per step in ATen).
__device__ __forceinline__ int64_t label_of(const int64_t* __restrict__ labels, int64_t r, int64_t rows, int64_t shift_T,
                                            int64_t ignore_index) {
  if (shift_T <= 0) return labels[r];
  return (r % shift_T == shift_T - 1 || r + 1 >= rows) ? ignore_index : labels[r + 1];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
ce_rows_kernel(const T* __restrict__ x, int64_t vocab, int64_t row_stride, const int64_t* __restrict__ labels,
               int64_t ignore_index, int aligned, float* __restrict__ row_loss, int64_t shift_T) {
  const int64_t r = blockIdx.x;
  const int64_t label = label_of(labels, r, (int64_t)gridDim.x, shift_T, ignore_index);
  if (label == ignore_index) {        // not part of the mean: the row is never read
    if (threadIdx.x == 0) row_loss[r] = 0.f;
    return;
  }
  const T* row = x + r * row_stride;
  ML acc{-INFINITY, 0.f};
  constexpr int N = Vec16<T>::N;
  if (aligned) {
    const int64_t nvec = vocab / N;
    const Vec16<T>* rv = reinterpret_cast<const Vec16<T>*>(row);
    int64_t i = threadIdx.x;
    for (; i + kThreads < nvec; i += 2 * kThreads) {               // two independent 16-byte loads in flight
      const Vec16<T> a = rv[i], b = rv[i + kThreads];
      float t[2 * N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        t[j] = DT<T>::to_f(a.v[j]);
        t[N + j] = DT<T>::to_f(b.v[j]);
      }
      ml_add_block<2 * N>(acc, t);
    }
    for (; i < nvec; i += kThreads) {
      const Vec16<T> a = rv[i];
      float t[N];
#pragma unroll
      for (int j = 0; j < N; ++j) t[j] = DT<T>::to_f(a.v[j]);
      ml_add_block<N>(acc, t);
    }
    for (int64_t j = nvec * N + threadIdx.x; j < vocab; j += kThreads) {
      const float t[1] = {DT<T>::to_f(row[j])};
      ml_add_block<1>(acc, t);
    }
  } else {
    for (int64_t j = threadIdx.x; j < vocab; j += kThreads) {
      const float t[1] = {DT<T>::to_f(row[j])};
      ml_add_block<1>(acc, t);
    }
  }
  // block merge
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ML b{__shfl_xor_sync(0xffffffffu, acc.m, o), __shfl_xor_sync(0xffffffffu, acc.l, o)};
    acc = ml_merge(acc, b);
  }
  __shared__ ML part[kThreads / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    ML a = part[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) a = ml_merge(a, part[w]);
    float loss;
    if (label < 0 || label >= vocab) {
      loss = __int_as_float(0x7fc00000);                           // out-of-range target (ATen asserts): NaN
    } else {
      // logsumexp - x[label] = (m - x[label]) + ln l
      loss = (a.m - DT<T>::to_f(row[label])) + log2f(a.l) * kLn2;
    }
    row_loss[r] = loss;
  }
}

__global__ void __launch_bounds__(1024)
ce_mean_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ labels, int64_t rows,
               int64_t ignore_index, float* __restrict__ out_loss, float* __restrict__ out_count, int64_t shift_T) {
  double s = 0.0;
  long long c = 0;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) {
    if (label_of(labels, r, rows, shift_T, ignore_index) != ignore_index) {
      s += (double)row_loss[r];
      ++c;
    }
  }
  __shared__ double ss[1024];
  __shared__ long long sc[1024];
  ss[threadIdx.x] = s;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      sc[threadIdx.x] += sc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out_loss[0] = (float)(ss[0] / (double)sc[0]);                  // no labelled token: 0/0 = NaN, like ATen
    if (out_count) out_count[0] = (float)sc[0];
  }
}

template <typename T>
int launch(const void* x, int64_t rows, int64_t vocab, int64_t row_stride, const int64_t* labels, int64_t ignore_index,
           float* out_loss, float* out_count, float* row_loss, cudaStream_t st, int64_t shift_T) {
  const int aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (row_stride * (int64_t)sizeof(T)) % 16 == 0;
  ce_rows_kernel<T><<<(unsigned)rows, kThreads, 0, st>>>((const T*)x, vocab, row_stride, labels, ignore_index, aligned,
                                                         row_loss, shift_T);
  int rc = check_launch("ce_rows_kernel");
  if (rc) return rc;
  ce_mean_kernel<<<1, 1024, 0, st>>>(row_loss, labels, rows, ignore_index, out_loss, out_count, shift_T);
  return check_launch("ce_mean_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" size_t oeh_ce_loss_workspace_bytes(int64_t rows) { return rows > 0 ? (size_t)rows * sizeof(float) : 0; }

extern "C" int oeh_ce_loss_shifted(const void* logits, int64_t rows, int64_t vocab, int64_t row_stride, int32_t dtype,
                                   const int64_t* labels, int64_t ignore_index, int64_t shift_T, float* out_loss,
                                   float* out_count, void* workspace, size_t workspace_bytes, oeh_stream_t stream);

extern "C" int oeh_ce_loss(const void* logits, int64_t rows, int64_t vocab, int64_t row_stride, int32_t dtype,
                           const int64_t* labels, int64_t ignore_index, float* out_loss, float* out_count,
                           void* workspace, size_t workspace_bytes, oeh_stream_t stream) {
  return oeh_ce_loss_shifted(logits, rows, vocab, row_stride, dtype, labels, ignore_index, 0, out_loss, out_count, workspace,
                             workspace_bytes, stream);
}

extern "C" int oeh_ce_loss_shifted(const void* logits, int64_t rows, int64_t vocab, int64_t row_stride, int32_t dtype,
                                   const int64_t* labels, int64_t ignore_index, int64_t shift_T, float* out_loss,
                                   float* out_count, void* workspace, size_t workspace_bytes, oeh_stream_t stream) {
  OEH_CHECK_ARG(rows > 0 && vocab > 0, "oeh_ce_loss: empty batch");
  OEH_CHECK_ARG(shift_T >= 0 && (shift_T == 0 || rows % shift_T == 0),
                "oeh_ce_loss_shifted: rows=%lld is not a whole number of blocks of %lld tokens", (long long)rows, (long long)shift_T);
  OEH_CHECK_ARG(rows <= 0x7fffffffLL, "oeh_ce_loss: at most 2^31-1 rows per call");
  OEH_CHECK_ARG(row_stride >= vocab, "oeh_ce_loss: row_stride %lld < vocab %lld", (long long)row_stride, (long long)vocab);
  OEH_CHECK_ARG(logits && labels && out_loss, "oeh_ce_loss: null pointer");
  const size_t need = oeh_ce_loss_workspace_bytes(rows);
  if (!workspace || workspace_bytes < need)
    return set_error(OEH_ERR_WORKSPACE, "oeh_ce_loss: workspace of %zu bytes needed", need);
  OEH_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 3) == 0, "oeh_ce_loss: workspace must be 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* row_loss = (float*)workspace;
  switch (dtype) {
    case OEH_F32: return launch<float>(logits, rows, vocab, row_stride, labels, ignore_index, out_loss, out_count, row_loss, st, shift_T);
    case OEH_F16: return launch<__half>(logits, rows, vocab, row_stride, labels, ignore_index, out_loss, out_count, row_loss, st, shift_T);
    case OEH_BF16:
      return launch<__nv_bfloat16>(logits, rows, vocab, row_stride, labels, ignore_index, out_loss, out_count, row_loss, st, shift_T);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_ce_loss: unknown dtype %d", dtype);
}
