/* oeh.h -- C ABI of liboeh.so: the B200 (sm_100a) implementation of OutEffHop's
 * attention + activation fake-quant hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / a cudaStream_t
 * (as void*), returns an int status (0 = ok) and never throws or aborts.  All work is
 * enqueued asynchronously on the given stream, allocates nothing, performs no host
 * synchronisation and is CUDA-graph capturable; quantizer ranges are read from DEVICE
 * memory so a range estimated on the stream can be consumed on the same stream.
 *
 * Each declaration names the reference interface it replaces (paths relative to
 * /root/reference/OutEffHop).  The reference is pure PyTorch, so "the FFI" a maintainer
 * binds is ctypes from the reference's Python modules -- see INTEGRATION.md.
 */
#ifndef OEH_H_
#define OEH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OEH_VERSION 100 /* major*100 + minor */

typedef void* oeh_stream_t; /* cudaStream_t */

enum oeh_status {
  OEH_OK = 0,
  OEH_ERR_BAD_ARG = 1,     /* -> ValueError in the Python shim            */
  OEH_ERR_UNSUPPORTED = 2, /* -> NotImplementedError (shape/dtype/branch) */
  OEH_ERR_CUDA = 3,        /* -> RuntimeError, see oeh_last_error()       */
  OEH_ERR_WORKSPACE = 4,   /* workspace missing or too small              */
  OEH_ERR_NOT_INITIALIZED = 5 /* -> QuantizerNotInitializedError          */
};

enum oeh_dtype { OEH_F32 = 0, OEH_F16 = 1, OEH_BF16 = 2 };

/* SOFTMAX_MAPPING families (transformers_language/models/softmax.py:22-64). */
enum oeh_softmax_kind {
  OEH_SOFTMAX_VANILLA = 0,  /* F.softmax                                      softmax.py:23      */
  OEH_SOFTMAX_1 = 1,        /* exp(x_i)/(1+sum_j exp(x_j))        vutils/softmax_1.py:4-28      */
  OEH_SOFTMAX_CLIPPED = 2,  /* clip((eta-gamma)*softmax+gamma,0,1)            softmax.py:10-13   */
  OEH_SOFTMAX_CLIPPED1 = 3  /* clip((eta-gamma)*softmax_1+gamma,0,1)          softmax.py:16-19   */
};

/* Device-resident state of one activation quantizer + its running min/max estimator:
 * 8 consecutive floats.  Mirrors the reference buffers
 *   RunningMinMaxEstimator.current_xmin/current_xmax   (quantization/range_estimators.py:16-17)
 *   AsymmetricUniformQuantizer._delta/_zero_float      (quantizers/uniform_quantizers.py:32-33)
 * plus the derived scale (:72-76) and zero_point (:79-82) the kernels consume. */
enum oeh_qstate_slot {
  OEH_QS_CUR_MIN = 0,
  OEH_QS_CUR_MAX = 1,
  OEH_QS_DELTA = 2,
  OEH_QS_ZERO_FLOAT = 3,
  OEH_QS_SCALE = 4,      /* &state[OEH_QS_SCALE] is the `qparams` pointer ([scale, zero_point]) */
  OEH_QS_ZERO_POINT = 5,
  OEH_QS_INITIALIZED = 6, /* 0.0f until the first range update */
  OEH_QS_RESERVED = 7,
  OEH_QS_FLOATS = 8
};

int oeh_version(void);
/* Thread-local description of the last non-zero status returned on this thread. */
const char* oeh_last_error(void);
/* 1 if the current device is compute capability 10.x (the only supported target). */
int oeh_device_supported(void);

/* ---------------------------------------------------------------------------------------------
 * Fused attention forward: QK^T -> *scale -> [qdq] -> +mask -> softmax_fn -> [qdq] -> PV
 *                           -> [qdq before gate] -> *sigmoid(gate_logits)*gate_scale -> [qdq]
 * Replaces the body of BertSelfAttentionWithExtras.forward (models/bert_attention.py:222-337),
 * OPTAttentionWithExtras.forward (models/opt_attention.py:204-322) and, with the qdq pointers
 * set, of QuantizedBertSelfAttentionWithExtras.forward (models/quantized_bert.py:317-434) and
 * QuantizedOPTAttentionWithExtras.forward (models/quantized_opt.py:151-257) in fix_ranges state.
 * ------------------------------------------------------------------------------------------- */
typedef struct OehAttnArgs {
  /* Q (B,H,Tq,d), K/V (B,H,Tk,d), out (B,H,Tq,d) as strided views: element strides for
   * (batch, head, token); the head dimension d is contiguous.  BERT passes the permuted views of
   * the Linear outputs (bert_attention.py:164-167) and a (B,Tq,H*d) output (the permute of
   * :335-337 folded into the store); OPT passes head-major copies (opt_attention.py:146-147). */
  const void* q;
  const void* k;
  const void* v;
  void* out;
  int64_t q_strides[3];
  int64_t k_strides[3];
  int64_t v_strides[3];
  int64_t out_strides[3];
  int32_t B, H, Tq, Tk, d;
  int32_t dtype;     /* oeh_dtype of q,k,v,out: OEH_F16 or OEH_BF16 (fp32 accumulate), or OEH_F32: the tensors are
                      * fp32 (the reference's eval precision, accelerate_configs/1gpu_no_mp.yaml:14), Q / K / V are rounded to
                      * fp16 OPERANDS (saturating) in the workspace, accumulation / softmax / taps stay fp32 and the
                      * context is widened back: the fp16 path's accuracy (2e-3 max-abs), not an fp32 GEMM's */
  float scale;       /* 1/sqrt(d) for BERT (:265); 1.0 for OPT when q is pre-scaled (:167) */
  int32_t kind;      /* oeh_softmax_kind */
  double gamma, eta; /* clipped variants; (eta-gamma) is rounded to fp32 like the Python scalar */
  int32_t causal;    /* 1: key j attends iff j <= i + (Tk-Tq) (HF causal mask), no mask tensor read */
  /* optional additive mask, element strides for (batch, head, query); key dim contiguous; a
   * stride of 0 broadcasts.  HF-4.31: (B,1,1,Tk) for BERT, (B,1,Tq,Tk) for OPT. */
  const void* mask;
  int32_t mask_dtype; /* oeh_dtype */
  int64_t mask_strides[3];
  int32_t clamp_min;  /* 1: max(s+mask, finfo(dtype).min) (opt_attention.py:221-223) */
  /* optional per-token gate LOGITS (B,H,Tq) fp32 contiguous; ctx *= sigmoid(logit)*gate_scale
   * (bert_attention.py:294-331; opt_attention.py:276-313). */
  const float* gate_logits;
  float gate_scale;
  /* optional fixed-range fake-quant taps: device pointers to [scale, zero_point]
   * (quantized_bert.py:363,374,434; quantized_opt.py:154,182,210). */
  const float* qdq_scores;
  const float* qdq_probs;
  const float* qdq_ctx;
  int32_t n_bits;
  int32_t ctx_qdq_before_gate; /* OPT taps PV before the gate (:210), BERT after it (:434) */
  void* workspace;             /* oeh_attn_workspace_bytes() bytes (non-zero for OEH_F32 only), 256-B aligned, may be NULL if 0 */
  size_t workspace_bytes;
  oeh_stream_t stream;
} OehAttnArgs;

size_t oeh_attn_workspace_bytes(const OehAttnArgs* args);
int oeh_attn_fwd(const OehAttnArgs* args);

/* ---------------------------------------------------------------------------------------------
 * Row softmax on a materialised score tensor (the unfused / range-estimation path):
 *   y = softmax_fn(clamp(x*in_scale + mask))   over the last dimension.
 * Replaces softmax_1 / clipped_softmax / clipped_softmax1 / F.softmax as called at
 * bert_attention.py:276 and opt_attention.py:226-232 (vutils/softmax_1.py:4-28; softmax.py:10-19).
 * x,y: (B*H*Tq, Tk) contiguous.  mask as in OehAttnArgs (rows decompose as ((b*H)+h)*Tq+q).
 * ------------------------------------------------------------------------------------------- */
int oeh_softmax_rows(const void* x, void* y, int32_t B, int32_t H, int32_t Tq, int32_t Tk, int32_t dtype,
                     float in_scale, int32_t kind, double gamma, double eta, const void* mask,
                     int32_t mask_dtype, const int64_t* mask_strides, int32_t causal, int32_t clamp_min,
                     oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Per-token gate logits on the attention INPUT split by head (bert_attention.py:303-320;
 * opt_attention.py:285-302): out[b,h,t] = fc_h(x[b,t,h*d:(h+1)*d]).
 * hidden == 0: fc_h = Linear(d,1): w1 (H,d), b1 (H); w2,b2 ignored.
 * hidden  > 0: fc_h = Linear(d,hidden) -> ReLU -> Linear(hidden,1) (--attn_gate_mlp / _mlp2,
 *              bert_attention.py:126-143): w1 (H,hidden,d), b1 (H,hidden), w2 (H,hidden), b2 (H).
 * x: (B,T,H*d) with row stride x_row_stride elements; weights fp32; out (B,H,T) fp32.
 * ------------------------------------------------------------------------------------------- */
int oeh_gate_logits(const void* x, int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t d,
                    int64_t x_row_stride, const float* w1, const float* b1, const float* w2, const float* b2,
                    int32_t hidden, float* out, oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Asymmetric uniform fake-quant (AsymmetricUniformQuantizer.forward / to_integer_forward,
 * quantization/quantizers/uniform_quantizers.py:92-148):
 *   x_int = clamp(round_half_even(x/scale) + zp, 0, 2^n-1);  y = scale*(x_int - zp)
 * qparams: DEVICE pointer to [scale, zero_point].  emit_codes=1 writes x_int instead of y.
 * fp32 tensors are bit-exact with the reference; fp16/bf16 tensors follow ATen-CUDA semantics
 * (every scalar converted to the tensor dtype, one rounding per op).
 * ------------------------------------------------------------------------------------------- */
int oeh_fake_quant_asym(const void* x, void* y, int64_t n, int32_t dtype, const float* qparams,
                        int32_t n_bits, int32_t emit_codes, oeh_stream_t stream);

/* Global min/max of a tensor (torch.min / torch.max at range_estimators.py:96-97; NaN propagates).
 * out2: DEVICE [min, max].  Workspace from oeh_minmax_workspace_bytes(n). */
size_t oeh_minmax_workspace_bytes(int64_t n);
int oeh_minmax(const void* x, int64_t n, int32_t dtype, float* out2, void* workspace, size_t workspace_bytes,
               oeh_stream_t stream);

/* RunningMinMaxEstimator.forward EMA (range_estimators.py:99-104) followed by
 * AsymmetricUniformQuantizer.set_quant_range (uniform_quantizers.py:204-224), entirely on the
 * device: state (OEH_QS_FLOATS floats) is updated from new_minmax (DEVICE [min,max], e.g. the
 * output of oeh_minmax after an optional cross-GPU MIN/MAX all-reduce). */
int oeh_range_update(float* state, const float* new_minmax, double momentum, int32_t n_bits, double eps,
                     oeh_stream_t stream);

/* set_quant_range from host scalars (QuantizationManager(x_min=,x_max=), quantization_manager.py:54-56):
 * out4 = [_delta, _zero_float, scale, zero_point], same fp32 op order as the device path. */
int oeh_set_quant_range_host(float x_min, float x_max, int32_t n_bits, double eps, float* out4);

/* QuantizationManager.forward in estimate_ranges state (quantization_manager.py:94-102) for one
 * GPU: minmax -> EMA -> set range -> quantize the same tensor, 3 launches, no host sync. */
int oeh_estimate_and_quantize(const void* x, void* y, int64_t n, int32_t dtype, float* state, double momentum,
                              int32_t n_bits, double eps, void* workspace, size_t workspace_bytes,
                              oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Percentile range estimation: np.percentile(x, (q_lo, q_hi)), default 'linear' method, over
 * the WHOLE tensor, as RunningMinMaxEstimator / CurrentMinMaxEstimator evaluate it on a host copy
 * (quantization/range_estimators.py:89-94, :62-69; --percentile at transformers_language/args.py:399).
 * Exact: the four neighbouring order statistics are found by radix select on the device, the interpolation follows
 * numpy's _lerp (difference rounded to the tensor dtype, the rest in float64).  out2: DEVICE [p(q_lo), p(q_hi)] as
 * float64 (the reference's torch.tensor(np.float64)).  fp32 / fp16 only (numpy has no bfloat16).  Any NaN -> NaN.
 *
 * oeh_percentile runs everything for one process.  The step-wise entry points exist so that a caller can all-reduce
 * (SUM, int32) the histogram -- the first 4*2048 uint32 of the workspace -- across ranks between _hist and _scan:
 * the result is then the percentile of the union of all ranks' tensors (n_total = sum of n in oeh_percentile_plan).
 *   plan(host) -> begin -> for pass in 0..passes-1: hist, [all-reduce], scan -> finish
 * ------------------------------------------------------------------------------------------- */
size_t oeh_percentile_workspace_bytes(void);
int oeh_percentile_passes(int32_t dtype);   /* 3 for fp32, 2 for fp16, 0 = unsupported dtype */
int oeh_percentile_plan(int64_t n_total, double q_lo, double q_hi, int64_t ranks_out[4], double gamma_out[2]);
int oeh_percentile_begin(void* workspace, size_t workspace_bytes, const int64_t ranks[4], oeh_stream_t stream);
int oeh_percentile_hist(const void* x, int64_t n, int32_t dtype, int32_t pass, void* workspace, oeh_stream_t stream);
int oeh_percentile_scan(int32_t dtype, int32_t pass, void* workspace, oeh_stream_t stream);
int oeh_percentile_finish(int32_t dtype, const double gamma[2], double* out2, void* workspace, oeh_stream_t stream);
int oeh_percentile(const void* x, int64_t n, int32_t dtype, double q_lo, double q_hi, double* out2, void* workspace,
                   size_t workspace_bytes, oeh_stream_t stream);

/* RunningMinMaxEstimator EMA + set_quant_range for the float64 ranges of the percentile branch
 * (range_estimators.py:99-104 on float64 0-dim tensors; uniform_quantizers.py:199-224).  state64: DEVICE
 * [cur_min, cur_max, _delta, _zero_float, initialised] as float64; qstate: the OEH_QS_FLOATS fp32 state the quantize
 * kernels read (scale / zero_point rounded to fp32, as ATen does when it applies a 0-dim float64 to an fp32 tensor). */
int oeh_range_update_f64(double* state64, float* qstate, const double* new_minmax, double momentum, int32_t n_bits,
                         double eps, oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Residual add fused with the activation fake-quant tap:
 *   y = qdq(a + b)    -- QuantizedBertSelfOutput / QuantizedBertOutput.forward
 *                        (transformers_language/models/quantized_bert.py:566-567, :602-603),
 *                        QuantizedOPTDecoderLayer residual sums (quantized_opt.py:342-345, :367-369).
 * The sum is rounded once to the tensor dtype (as ATen's add), then quantised exactly like oeh_fake_quant_asym:
 * bit-exact with the reference for fp32.  Fixed ranges only (qparams = DEVICE [scale, zero_point]); range
 * estimation needs min/max of the sum first and uses the unfused kernels.  y may alias a or b.
 * ------------------------------------------------------------------------------------------- */
int oeh_add_fake_quant(const void* a, const void* b, void* y, int64_t n, int32_t dtype, const float* qparams,
                       int32_t n_bits, oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Outlier metrics of an activation tensor viewed as (rows, cols): per row
 *   inf_norm[r] = max_j |x[r,j]|                      (x.norm(dim=1, p=inf), validate_mlm_config.py:513-518)
 *   kurt[r]     = mean((x-mu)^4) / (std^4 + eps)      (kurtosis(), transformers_language/utils.py:9-20;
 *                                                      std unbiased, eps = 1e-6)
 * in ONE pass over x (power sums accumulated in float64), replacing per-sample .item() host syncs.
 * out_inf / out_kurt: DEVICE fp32 (rows).  meter (optional, DEVICE, 3 float64): running AverageMeter state
 * [sum of inf-norms, sum of kurtoses, samples] += this call's rows, so a whole eval needs one read-back.
 * ------------------------------------------------------------------------------------------- */
size_t oeh_outlier_stats_workspace_bytes(int64_t rows, int64_t cols);
int oeh_outlier_stats(const void* x, int64_t rows, int64_t cols, int32_t dtype, double eps, float* out_inf,
                      float* out_kurt, double* meter, void* workspace, size_t workspace_bytes, oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Mean cross-entropy of one evaluation batch from the model's logits: the per-batch
 * `outputs.loss` the eval loops aggregate (validate_mlm_config.py:489-496, validate_clm.py:555-561), i.e. HF's
 * CrossEntropyLoss()(prediction_scores.view(-1, V), labels.view(-1)) evaluated in fp32:
 *   loss = mean over rows with label != ignore_index of ( logsumexp_j x[i,j] - x[i,label_i] ).
 * logits: (rows, vocab) of `dtype`, row stride `row_stride` elements (>= vocab: the vocabulary dimension may be
 * padded so rows stay 16-byte aligned); labels: DEVICE int64 (rows).  out_loss: DEVICE fp32 scalar (NaN when no row
 * is labelled, like ATen); out_count (optional): DEVICE fp32 number of labelled rows.  The per-row losses are left
 * in the workspace (rows fp32; 0 for ignored rows).  One pass over the labelled rows, no fp32 copy of the logits.
 * ------------------------------------------------------------------------------------------- */
size_t oeh_ce_loss_workspace_bytes(int64_t rows);
/* oeh_ce_loss with HF's causal-LM shift (validate_clm.py:555-561: shift_logits = logits[..., :-1, :], shift_labels =
 * labels[..., 1:]) applied to the LABEL INDEX instead of to copies of the tensors: with shift_T > 0 the rows are blocks of
 * shift_T tokens, row (b, t) is scored against labels[b, t + 1] and the last row of every block is ignored -- the same
 * B * (T - 1) terms.  shift_T == 0 is oeh_ce_loss. */
int oeh_ce_loss_shifted(const void* logits, int64_t rows, int64_t vocab, int64_t row_stride, int32_t dtype,
                        const int64_t* labels, int64_t ignore_index, int64_t shift_T, float* out_loss, float* out_count,
                        void* workspace, size_t workspace_bytes, oeh_stream_t stream);
int oeh_ce_loss(const void* logits, int64_t rows, int64_t vocab, int64_t row_stride, int32_t dtype,
                const int64_t* labels, int64_t ignore_index, float* out_loss, float* out_count, void* workspace,
                size_t workspace_bytes, oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Residual add + LayerNorm in one pass, optionally with the two fake-quant taps of the INT8 residual blocks
 * (SURVEY 8f rank 2):   y = [qdq_out] LayerNorm( [qdq_res] ([qdq_a] a + b) )
 *   FP   (qparams_* NULL): BertSelfOutput / BertOutput.forward, the HF modules wrapped at
 *                          transformers_language/models/quantized_bert.py:540-605.
 *   INT8 (fixed ranges)  : QuantizedBertSelfOutput / QuantizedBertOutput.forward (quantized_bert.py:566-568, :602-604):
 *                          [qparams_a: the output tap of the QuantLinear `dense` that produced a,] res_act_quantizer on
 *                          the sum, then QuantLayerNorm (quantization/autoquant_utils.py:62-70) with its output tap.
 * a, b (b may be NULL: plain LayerNorm), y: (rows, cols) contiguous of `dtype`; weight / bias (bias may be NULL): (cols)
 * of the same dtype.  The sum is rounded once to the tensor dtype (ATen add), statistics / affine map in fp32 (biased
 * variance, rsqrt(var + eps)), one rounding of y; taps exactly as oeh_fake_quant_asym.  cols must be a multiple of
 * 16/sizeof(dtype) and at most 256 such vectors (2048 fp16 / 1024 fp32 columns); all pointers 16-byte aligned.
 * sum_out (optional, (rows, cols)): also receives the (tapped) sum -- pre-LayerNorm blocks (OPTDecoderLayer, HF
 * modeling_opt; quantized_opt.py:338-357) need it as the next residual.
 * ------------------------------------------------------------------------------------------- */
int oeh_add_layer_norm(const void* a, const void* b, const void* weight, const void* bias, void* y, int64_t rows,
                       int64_t cols, double eps, int32_t dtype, const float* qparams_a, const float* qparams_res,
                       const float* qparams_out, int32_t n_bits, void* sum_out, oeh_stream_t stream);

/* oeh_add_layer_norm that ALSO emits the per-token attention-gate logits of the module that consumes y (Linear(64,1) gate
 * predictors, bert_attention.py:303-320 / opt_attention.py:285-302: out[b,h,t] = w_h . y[b,t,64h:64h+64] + b_h):
 * gate_w (H,64) fp32 contiguous and 16-byte aligned, gate_b (H) fp32, gate_out (B,H,T) fp32, rows == B*T, cols == 64*H.  Saves the separate
 * oeh_gate_logits pass over the hidden states.  gate_out NULL = oeh_add_layer_norm. */
int oeh_add_layer_norm_gate(const void* a, const void* b, const void* weight, const void* bias, void* y, int64_t rows,
                            int64_t cols, double eps, int32_t dtype, const float* qparams_a, const float* qparams_res,
                            const float* qparams_out, int32_t n_bits, void* sum_out, const float* gate_w,
                            const float* gate_b, float* gate_out, int32_t T, int32_t H, oeh_stream_t stream);


/* y = qdq_sum(a + qdq_b(b)): oeh_add_fake_quant with the output tap of the QuantLinear that produced b folded in
 * (QuantizedOPTDecoderLayer: fc2's tap, the residual add and ffn_res_act_quantizer, quantized_opt.py:362-369).
 * qparams_b may be NULL (= oeh_add_fake_quant).  Fixed ranges only. */
int oeh_tap_add_fake_quant(const void* a, const void* b, void* y, int64_t n, int32_t dtype, const float* qparams_b,
                           const float* qparams_sum, int32_t n_bits, oeh_stream_t stream);

/* Fake-quant of n_seg (<= 4) column segments of seg_cols columns, each with its own DEVICE [scale, zero_point]
 * (qparams: HOST array of n_seg device pointers); x, y: (rows, n_seg*seg_cols) with row strides in elements.  One
 * launch taps the output of a fused query|key|value projection whose three QuantLinear modules keep their own ranges
 * (quantized_bert.py:317-330, quantized_opt.py:113-140).  Same arithmetic as oeh_fake_quant_asym. */
int oeh_fake_quant_asym_segments(const void* x, void* y, int64_t rows, int32_t seg_cols, int32_t n_seg,
                                 int64_t x_row_stride, int64_t y_row_stride, int32_t dtype, const float* const* qparams,
                                 int32_t n_bits, oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * erf GELU of the feed-forward block: y = x * 0.5 * (1 + erf(x / sqrt 2)), the activation HF's BertIntermediate applies
 * between the two FFN GEMMs (transformers 4.31 modeling_bert.py, ACT2FN["gelu"]; the reference wraps it as
 * QuantizedBertIntermediate / activation "gelu" at quantized_bert.py:609-617).  fp16 / bf16 tensors, fp32 arithmetic,
 * one rounding to the tensor dtype; x == y (in place) is allowed.  Evaluated through erfc (relative accuracy ~1e-6 on
 * both tails): correctly rounded for all but 5 of the 63488 finite fp16 inputs (1 ulp) and for every bf16 input.  fp32 tensors: OEH_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------------------------- */
int oeh_gelu_erf(const void* x, void* y, int64_t n, int32_t dtype, oeh_stream_t stream);

/* The OPT decoder's input (HF modeling_opt, the shell around the reference's OPTAttentionWithExtras; quantized_opt.py:432-457
 * builds the same sum): y[r] = word[ids[r]] + pos_table[pos_ids[r] + pos_offset] (OPTLearnedPositionalEmbedding adds its offset
 * of 2 inside forward), rounded once to the tensor dtype like the ATen add.  pos_ids NULL = r % T.  ids / positions are
 * clamped into their tables.  rows = B*T; cols a multiple of 16 / sizeof(dtype); tables and y 16-byte aligned. */
int oeh_embed_sum(const int64_t* ids, const int64_t* pos_ids, const void* word, const void* pos_table, void* y, int64_t rows,
                  int64_t cols, int32_t T, int64_t vocab, int64_t n_pos, int64_t pos_offset, int32_t dtype, oeh_stream_t stream);

/* HF BertEmbeddings.forward (transformers modeling_bert; the embedding shell the reference's models keep, quantized at
 * quantized_bert.py:146-218) in one launch: y[r] = LayerNorm((word[ids[r]] + type[type_ids[r]]) + pos[pos_ids[r]]), sums in
 * the tensor dtype in that order (each rounded once, as the two ATen adds do).  type_ids NULL = type 0, pos_ids NULL =
 * r % T; type_table / pos_table may be NULL.  ids are clamped to [0, vocab).  rows = B*T. */
int oeh_embed_layer_norm(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const void* word,
                         const void* type_table, const void* pos_table, const void* weight, const void* bias, void* y,
                         int64_t rows, int64_t cols, int32_t T, int64_t vocab, double eps, int32_t dtype, oeh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Linear layer with fused epilogue:  y = tap( act( x . w^T + bias ) ),  x (M,K), w (N,K) = nn.Linear.weight, y (M,N),
 * row strides ldx / ldw / ldy in elements, fp16 / bf16 operands, fp32 accumulation (tcgen05, TMEM accumulators).
 * Replaces, in one launch:
 *   - QuantLinear.forward = F.linear -> [fused activation] -> activation_quantizer (quantization/hijacker.py:78-134,
 *     autoquant_utils.py:18-91; call sites quantized_bert.py:278,300-303,541-606, quantized_opt.py:113-140) with FIXED ranges:
 *     qparams = HOST array of n_seg (<= 4) device pointers to [scale, zero_point]; n_seg > 1 = equal column segments with their
 *     own ranges (the fused query|key|value projection, quantized_bert.py:317-330); n_seg = 0 = no tap;
 *   - HF BertIntermediate: dense -> erf GELU (act = OEH_ACT_GELU_ERF; same evaluation as oeh_gelu_erf), OPT fc1 -> ReLU.
 * Order of roundings follows the unfused chain: the accumulator + bias is rounded to the tensor dtype (what the GEMM would
 * have stored), then the activation (rounded again), then the tap.  The tap is tolerance-parity with
 * `library GEMM -> oeh_fake_quant_asym` (accumulation order moves values across bin edges), like the fused attention taps.
 * tile_n: 0 = choose; 128 / 256 = single-CTA tiles of 128 rows x that many columns; 2128 / 2256 = CTA pairs
 * (tcgen05 cta_group::2) on 256 x 128 / 256 x 256 tiles.  Asynchronous, allocation-free, graph-capturable.
 * ------------------------------------------------------------------------------------------- */
enum oeh_activation { OEH_ACT_NONE = 0, OEH_ACT_RELU = 1, OEH_ACT_GELU_ERF = 2 };

typedef struct OehLinearArgs {
  const void* x;
  const void* w;
  const void* bias;            /* (N) of the tensor dtype, 16-byte aligned, or NULL */
  void* y;
  int64_t ldx, ldw, ldy;
  int32_t M, N, K;
  int32_t dtype;               /* OEH_F16 or OEH_BF16 */
  int32_t act;                 /* oeh_activation */
  int32_t n_seg;               /* 0..4 tap segments */
  int32_t n_bits;
  int32_t tile_n;
  const float* qparams[4];     /* device [scale, zero_point] per segment */
  oeh_stream_t stream;
} OehLinearArgs;

int oeh_linear_fwd(const OehLinearArgs* args);

/* ---------------------------------------------------------------------------------------------
 * Linear -> (+ residual) -> LayerNorm in one launch:
 *     v = round(x . w^T + bias);   s = [qdq_res]([qdq_a](v) + residual);   y = [qdq_out](LayerNorm(s) * ln_weight + ln_bias)
 * x (M,K), w (N,K) = nn.Linear.weight (row strides ldx / ldw in elements), residual / y / sum_out (M,N) contiguous, fp16 / bf16,
 * fp32 accumulation and statistics.  Replaces
 *   - HF BertSelfOutput / BertOutput.forward: LayerNorm(dense(hidden_states) + input_tensor) (transformers modeling_bert; the
 *     modules the reference wraps at quantized_bert.py:541-606), and with q_a (the QuantLinear's output tap, hijacker.py:119-134),
 *     q_res (res_act_quantizer, quantized_bert.py:566-567, :602-603) and q_out (QuantLayerNorm's tap, autoquant_utils.py:62-70)
 *     their quantized forms with FIXED ranges (device [scale, zero_point] pointers; NULL = no tap);
 *   - the residual joints of the OPT decoder layer (quantized_opt.py:338-370; HF OPTDecoderLayer.forward): out_proj / fc2 ->
 *     + residual -> the next LayerNorm, with sum_out = the (tapped) sum, which the pre-LayerNorm block carries on as its residual.
 * residual NULL = LayerNorm(dense(x)) (BertPredictionHeadTransform without its activation does not apply: use oeh_linear_fwd).
 * A row block is computed by a thread-block cluster of N / 192 or N / 256 CTAs that exchange the row statistics through
 * distributed shared memory: N must be 192 c or 256 c with c <= 8 (768, 1024, 1536, 2048 ...; oeh_linear_add_layer_norm_supported).
 * Same element arithmetic as `GEMM -> oeh_add_layer_norm` (sum and taps in the tensor dtype, two-pass statistics in fp32); the
 * fp32 summation orders differ, so the two agree to the rounding of the tensor dtype.  Asynchronous, allocation-free,
 * graph-capturable.  Measured on B200 it does NOT beat the two launches yet (profiles/r02_gemm_ln_fused.txt): the shipped model
 * paths keep `GEMM -> oeh_add_layer_norm`; this entry point is for callers that want the single launch.
 * ------------------------------------------------------------------------------------------- */
typedef struct OehLinearLnArgs {
  const void* x;
  const void* w;
  const void* bias;            /* (N) of the tensor dtype, 16-byte aligned, or NULL */
  const void* residual;        /* (M,N) or NULL */
  const void* ln_weight;       /* (N), 16-byte aligned */
  const void* ln_bias;         /* (N) or NULL */
  void* y;                     /* (M,N) */
  void* sum_out;               /* (M,N) or NULL */
  int64_t ldx, ldw;
  int32_t M, N, K;
  int32_t dtype;               /* OEH_F16 or OEH_BF16 */
  int32_t n_bits;
  double eps;
  const float* q_a;            /* device [scale, zero_point] or NULL */
  const float* q_res;
  const float* q_out;
  oeh_stream_t stream;
} OehLinearLnArgs;

int oeh_linear_add_layer_norm(const OehLinearLnArgs* args);
int oeh_linear_add_layer_norm_supported(int32_t N);   /* 1 if N can be normalised by one cluster */

#ifdef __cplusplus
}
#endif
#endif /* OEH_H_ */
