"""CPU oracle for the OutEffHop attention + activation fake-quant hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``outeffhop_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs do, and there only as the checker / baseline.

This is a plain numpy restatement of the reference's algorithm (the reference is
pure PyTorch eager code, so "the algorithm" is a chain of fp32 elementwise ops,
reductions and two batched matmuls).  Every function cites the reference
file:line (relative to /root/reference/OutEffHop) that it follows.

Pinning: ``tests/golden/make_golden.py`` imports the *live* reference from
/root/reference in the build container and writes small fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against
them (bit-exact for the quantizer / range estimator, <=1e-6 for the softmax
family whose ``exp``/sum order differs between numpy and ATen) and against the
known-answer table of SURVEY.md section 4.

All arithmetic is done in the dtype of the input (np.float32 unless stated);
numpy, like ATen, evaluates float16 ops in fp32 and rounds per op.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------
# softmax family
# ----------------------------------------------------------------------------


def softmax_vanilla(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """torch.nn.functional.softmax as used by SOFTMAX_MAPPING["vanilla"]
    (transformers_language/models/softmax.py:23)."""
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True, dtype=x.dtype)


def softmax_n_shifted_zeros(x: np.ndarray, n: int, axis: int = -1) -> np.ndarray:
    """vutils/softmax_1.py:4-21.

    m = max_j x; e = exp(x - m); out = e / (sum e + n * exp(-m)).
    """
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        input_maxes = np.max(x, axis=axis, keepdims=True)            # :11
        shifted_inputs = x - input_maxes                              # :13
        numerator = np.exp(shifted_inputs)                            # :15
        original_denominator = np.sum(numerator, axis=axis, keepdims=True, dtype=x.dtype)  # :16
        shifted_zeros = input_maxes * x.dtype.type(-1)                # :18
        denominator = original_denominator + np.exp(shifted_zeros) * x.dtype.type(n)  # :20
        return numerator / denominator                                # :21


def softmax_1(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """vutils/softmax_1.py:24-28."""
    return softmax_n_shifted_zeros(x, 1, axis=axis)


def _stretch_clip(sm_out: np.ndarray, eta: float, gamma: float) -> np.ndarray:
    # softmax.py:12-13 / :18-19.  (eta - gamma) is evaluated in Python double and
    # converted to the tensor dtype by the multiply; gamma likewise by the add.
    t = sm_out.dtype.type
    stretched = sm_out * t(eta - gamma) + t(gamma)
    return np.clip(stretched, t(0), t(1))


def clipped_softmax(x: np.ndarray, axis: int = -1, eta: float = 1.1, gamma: float = -0.1) -> np.ndarray:
    """transformers_language/models/softmax.py:10-13."""
    return _stretch_clip(softmax_vanilla(x, axis=axis), eta, gamma)


def clipped_softmax1(x: np.ndarray, axis: int = -1, eta: float = 1.1, gamma: float = -0.1) -> np.ndarray:
    """transformers_language/models/softmax.py:16-19."""
    return _stretch_clip(softmax_1(x, axis=axis), eta, gamma)


# kind ids shared with include/oeh.h
KIND_VANILLA, KIND_SOFTMAX1, KIND_CLIPPED, KIND_CLIPPED1 = 0, 1, 2, 3


def _build_softmax_table():
    """The literal 40-key table of softmax.py:22-64 as (kind, gamma, eta).

    Two quirks are reproduced on purpose: "clipped(-.005:1.005)" has gamma=-0.003
    (softmax.py:57) and "clippedsoftmax1(-.025:1)" has eta=1.1 (softmax.py:61).
    "entmax" (softmax.py:25) is listed but is not part of the accelerated path.
    """
    t = {"vanilla": (KIND_VANILLA, 0.0, 1.0), "softmax1": (KIND_SOFTMAX1, 0.0, 1.0), "entmax": None}
    for eta in ("1.0003", "1.001", "1.002", "1.003", "1.004", "1.01", "1.02", "1.03", "1.1"):
        t[f"clipped(0:{eta})"] = (KIND_CLIPPED, 0.0, float(eta))                 # :27-35
    for g in (".1", ".00001", ".00003", ".0001", ".0003", ".0005", ".001", ".002", ".0025", ".003",
              ".004", ".005", ".01", ".015", ".02", ".025", ".03", ".04"):
        t[f"clipped(-{g}:1)"] = (KIND_CLIPPED, -float("0" + g), 1.0)             # :36-53
    t["clipped(-.001:1.001)"] = (KIND_CLIPPED, -0.001, 1.001)                    # :54
    t["clipped(-.002:1.002)"] = (KIND_CLIPPED, -0.002, 1.002)                    # :55
    t["clipped(-.003:1.003)"] = (KIND_CLIPPED, -0.003, 1.003)                    # :56
    t["clipped(-.005:1.005)"] = (KIND_CLIPPED, -0.003, 1.005)                    # :57 (quirk)
    t["clipped(-.01:1.01)"] = (KIND_CLIPPED, -0.01, 1.01)                        # :58
    t["clipped(-.03:1.03)"] = (KIND_CLIPPED, -0.03, 1.03)                        # :59
    t["clipped(-.1:1.1)"] = (KIND_CLIPPED, -0.1, 1.1)                            # :60
    t["clippedsoftmax1(-.025:1)"] = (KIND_CLIPPED1, -0.025, 1.1)                 # :61 (quirk)
    t["clippedsoftmax1(-.00001:1)"] = (KIND_CLIPPED1, -0.00001, 1.0)             # :62
    t["clippedsoftmax1(-.0001:1)"] = (KIND_CLIPPED1, -0.0001, 1.0)               # :63
    return t


SOFTMAX_TABLE = _build_softmax_table()


def softmax_by_kind(x: np.ndarray, kind: int, gamma: float = 0.0, eta: float = 1.0, axis: int = -1) -> np.ndarray:
    if kind == KIND_VANILLA:
        return softmax_vanilla(x, axis)
    if kind == KIND_SOFTMAX1:
        return softmax_1(x, axis)
    if kind == KIND_CLIPPED:
        return clipped_softmax(x, axis, eta=eta, gamma=gamma)
    if kind == KIND_CLIPPED1:
        return clipped_softmax1(x, axis, eta=eta, gamma=gamma)
    raise ValueError(f"unknown softmax kind {kind}")


def softmax_mapping(key: str):
    """SOFTMAX_MAPPING[key] (softmax.py:22-64) as a callable (data, dim=-1)."""
    ent = SOFTMAX_TABLE[key]
    if ent is None:
        raise NotImplementedError("entmax is outside the accelerated hot path")
    kind, gamma, eta = ent
    return lambda data, dim=-1: softmax_by_kind(data, kind, gamma, eta, axis=dim)


def logit(p: float, eps: float = 1e-16) -> float:
    """bert_attention.py:16-18."""
    p = float(np.clip(p, eps, 1 - eps))
    return -math.log(1 / p - 1)


# ----------------------------------------------------------------------------
# asymmetric uniform quantizer + running min/max
# ----------------------------------------------------------------------------


def asym_set_quant_range(x_min, x_max, n_bits: int = 8, eps: float = 1e-8) -> Tuple[np.float32, np.float32]:
    """AsymmetricUniformQuantizer.set_quant_range + _tensorize_min_max
    (quantization/quantizers/uniform_quantizers.py:173-224).  Returns
    (_delta, _zero_float) as fp32 scalars."""
    x_min = F32(x_min)
    x_max = F32(x_max)
    x_min = np.minimum(x_min, F32(0.0))                # :199
    x_max = np.maximum(x_max, F32(1.0) * F32(eps))     # :200
    int_max = F32(2.0 ** n_bits - 1)                   # :67-69
    delta = F32((x_max - x_min) / int_max)             # :218
    with np.errstate(divide="ignore", invalid="ignore"):
        zero_float = F32(-x_min / delta)               # :219
    return delta, zero_float


def asym_scale_zero_point(delta, zero_float, n_bits: int = 8, eps: float = 1e-8) -> Tuple[np.float32, np.float32]:
    """scale (uniform_quantizers.py:72-76, linear domain) and zero_point (:79-82)."""
    scale = np.maximum(F32(delta), F32(eps))                                   # clamp(min=eps)
    zp = np.clip(np.rint(F32(zero_float)), F32(0.0), F32(2.0 ** n_bits - 1))   # round-half-even
    return F32(scale), F32(zp)


def asym_to_integer(x: np.ndarray, delta, zero_float, n_bits: int = 8, eps: float = 1e-8,
                    lowp_mode: str = "cuda") -> np.ndarray:
    """to_integer_forward (uniform_quantizers.py:92-117) without grad scaling.
    The integer codes stay in a float tensor of x's dtype, exactly as the reference.

    fp32 tensors: every op is a plain fp32 op (bit-exact contract).
    fp16/bf16 tensors: ATen rounds after every op and converts the 0-dim fp32 scale / zero_point
    to the tensor dtype first -- except that ATen's *CPU* kernels keep a 0-dim SECOND operand of
    div/mul in fp32 (probed with torch 2.11: ``x_h / s`` uses fp32 s, ``s * x_h`` and add/sub use
    fp16 s).  ``lowp_mode="cpu"`` reproduces that quirk (it is what the CPU-generated goldens
    contain); ``lowp_mode="cuda"`` is what ATen's CUDA kernels do (dynamic cast of every operand to
    the common dtype) and is the contract of the fp16/bf16 CUDA quantizer.
    """
    scale, zp = asym_scale_zero_point(delta, zero_float, n_bits, eps)
    t = x.dtype.type
    zp_t = t(zp)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        if x.dtype == np.float32 or lowp_mode == "cuda":
            q = x / t(scale)
        else:
            q = (x.astype(np.float32) / F32(scale)).astype(x.dtype)
        x_int = np.rint(q) + zp_t                                              # :114
        x_int = np.clip(x_int, t(0.0), t(2.0 ** n_bits - 1))                   # :115
    return x_int


def asym_quantize(x: np.ndarray, delta, zero_float, n_bits: int = 8, eps: float = 1e-8,
                  lowp_mode: str = "cuda") -> np.ndarray:
    """AsymmetricUniformQuantizer.forward (uniform_quantizers.py:119-148): quantize-dequantize."""
    scale, zp = asym_scale_zero_point(delta, zero_float, n_bits, eps)
    t = x.dtype.type
    x_int = asym_to_integer(x, delta, zero_float, n_bits, eps, lowp_mode)
    return t(scale) * (x_int - t(zp))                                          # :146


class RunningMinMax:
    """RunningMinMaxEstimator, no-percentile / per-tensor branch
    (quantization/range_estimators.py:77-106)."""

    def __init__(self, momentum: float = 0.9):
        self.momentum = momentum
        self.current_xmin: Optional[np.float32] = None
        self.current_xmax: Optional[np.float32] = None

    def update_from_minmax(self, x_min, x_max):
        x_min, x_max = F32(x_min), F32(x_max)
        if self.current_xmin is None:                                          # :99-101
            self.current_xmin, self.current_xmax = x_min, x_max
        else:                                                                  # :103-104
            one_minus = F32(1 - self.momentum)   # python double -> fp32 at the multiply
            mom = F32(self.momentum)
            self.current_xmin = F32(F32(one_minus * x_min) + F32(mom * self.current_xmin))
            self.current_xmax = F32(F32(one_minus * x_max) + F32(mom * self.current_xmax))
        return self.current_xmin, self.current_xmax

    def __call__(self, x: np.ndarray):
        return self.update_from_minmax(np.min(x), np.max(x))                   # :96-97

    def reset(self):
        self.current_xmin = self.current_xmax = None


class QuantizationManager:
    """quantization/quantization_manager.py:11-108 restricted to
    asymmetric_uniform + running_minmax, states estimate_ranges / fix_ranges."""

    def __init__(self, n_bits: int = 8, momentum: float = 0.9, eps: float = 1e-8):
        self.n_bits, self.eps = n_bits, eps
        self.estimator = RunningMinMax(momentum)
        self.delta = None
        self.zero_float = None
        self.estimating = True

    def fix_ranges(self):
        if self.delta is None:
            raise RuntimeError("Quantizer has  not been initialized yet")      # quantizer_utils.py:31-37
        self.estimating = False

    def __call__(self, x: np.ndarray) -> np.ndarray:
        if self.estimating:                                                    # :94-100
            lo, hi = self.estimator(x)
            self.delta, self.zero_float = asym_set_quant_range(lo, hi, self.n_bits, self.eps)
        return asym_quantize(x, self.delta, self.zero_float, self.n_bits, self.eps)  # :102


# ----------------------------------------------------------------------------
# attention core and module forwards
# ----------------------------------------------------------------------------


def gate_logits_per_token(x_heads: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Linear(d,1) per head (bert_attention.py:311-320; opt_attention.py:293-302).
    x_heads (B,H,T,d), w (H,d), b (H,) -> logits (B,H,T)."""
    return np.einsum("bhtd,hd->bht", x_heads, w).astype(x_heads.dtype) + b[None, :, None]


def gate_logits_mlp(x_heads: np.ndarray, w1: np.ndarray, b1: np.ndarray, w2: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """Sequential(Linear(d,m), ReLU, Linear(m,1)) per head (bert_attention.py:126-143).
    w1 (H,m,d), b1 (H,m), w2 (H,m), b2 (H,)."""
    hid = np.einsum("bhtd,hmd->bhtm", x_heads, w1).astype(x_heads.dtype) + b1[None, :, None, :]
    hid = np.maximum(hid, x_heads.dtype.type(0))
    return np.einsum("bhtm,hm->bht", hid, w2).astype(x_heads.dtype) + b2[None, :, None]


def sigmoid(x: np.ndarray) -> np.ndarray:
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(x.dtype)


def attention_core(q, k, v, *, scale: float, kind: int, gamma: float = 0.0, eta: float = 1.0,
                   mask: Optional[np.ndarray] = None, causal: bool = False, clamp_min: Optional[float] = None,
                   gate: Optional[np.ndarray] = None, gate_scale: float = 1.0,
                   qdq_scores=None, qdq_probs=None, qdq_ctx=None, ctx_qdq_before_gate: bool = False,
                   return_probs: bool = False):
    """QK^T -> scale -> [qdq] -> +mask -> softmax_fn -> [qdq] -> PV -> gate -> [qdq].

    q,k,v: (B,H,T,d) fp32.  Follows bert_attention.py:222,265,270-276,292,294-331 and
    opt_attention.py:204,215-232,263,276-313; quantizer taps as quantized_bert.py:363,374,434
    (ctx after gate) and quantized_opt.py:154,182,210 (ctx before gate).
    ``mask`` is additive and broadcastable to (B,H,Tq,Tk); ``causal`` adds the HF-4.31
    lower-triangular finfo.min mask; ``clamp_min`` is OPT's max(w, finfo.min) (:221-223).
    ``gate`` is the (B,H,T) sigmoid output; qdq_* are callables x -> x_q.
    Returns ctx (B,H,T,d) (and probs when asked).
    """
    q = np.asarray(q)
    dt = q.dtype
    s = np.matmul(q, np.swapaxes(k, -1, -2)) * dt.type(scale)
    if qdq_scores is not None:
        s = qdq_scores(s)
    if causal:
        tq, tk = s.shape[-2], s.shape[-1]
        tri = np.triu(np.ones((tq, tk), dtype=bool), k=1 + (tk - tq))
        s = s + np.where(tri, np.finfo(dt).min, 0).astype(dt)
    if mask is not None:
        s = s + mask.astype(dt)
    if clamp_min is not None:
        s = np.maximum(s, dt.type(clamp_min))
    p = softmax_by_kind(s, kind, gamma, eta, axis=-1)
    if qdq_probs is not None:
        p = qdq_probs(p)
    ctx = np.matmul(p, v)
    if qdq_ctx is not None and ctx_qdq_before_gate:
        ctx = qdq_ctx(ctx)
    if gate is not None:
        ctx = ctx * (gate[..., None] * dt.type(gate_scale))
    if qdq_ctx is not None and not ctx_qdq_before_gate:
        ctx = qdq_ctx(ctx)
    return (ctx, p) if return_probs else ctx


def _linear(x, w, b):
    y = x @ w.T
    return y + b if b is not None else y


def _split_heads(x, h):
    b, t, c = x.shape
    return x.reshape(b, t, h, c // h).transpose(0, 2, 1, 3)


def bert_self_attention_forward(hidden, wq, bq, wk, bk, wv, bv, num_heads: int, *, kind: int,
                                gamma: float = 0.0, eta: float = 1.0, attention_mask=None,
                                gate_w=None, gate_b=None, gate_scale: float = 1.0) -> np.ndarray:
    """BertSelfAttentionWithExtras.forward (bert_attention.py:169-343), self-attention,
    absolute positions, eval mode, conditional_per_token Linear(d,1) gate.  Returns (B,T,H*d)."""
    qh = _split_heads(_linear(hidden, wq, bq), num_heads)          # :183,:164-167
    kh = _split_heads(_linear(hidden, wk, bk), num_heads)          # :205
    vh = _split_heads(_linear(hidden, wv, bv), num_heads)          # :206
    d = qh.shape[-1]
    gate = None
    if gate_w is not None:                                         # :294-331 (x = hidden_states!)
        gate = sigmoid(gate_logits_per_token(_split_heads(hidden, num_heads), gate_w, gate_b))
    # :265 divides by sqrt(d); expressed as a multiply by the fp32-rounded reciprocal only when exact
    s_scale = 1.0 / math.sqrt(d)
    ctx = attention_core(qh, kh, vh, scale=s_scale, kind=kind, gamma=gamma, eta=eta,
                         mask=attention_mask, gate=gate, gate_scale=gate_scale)
    b, h, t, _ = ctx.shape
    return ctx.transpose(0, 2, 1, 3).reshape(b, t, h * d)          # :335-337


def opt_attention_forward(hidden, wq, bq, wk, bk, wv, bv, wo, bo, num_heads: int, *, kind: int,
                          gamma: float = 0.0, eta: float = 1.0, attention_mask=None, causal: bool = False,
                          gate_w=None, gate_b=None, gate_scale: float = 1.0) -> np.ndarray:
    """OPTAttentionWithExtras.forward (opt_attention.py:149-326), self-attention, eval mode."""
    d = hidden.shape[-1] // num_heads
    scaling = hidden.dtype.type(d ** -0.5)
    qh = _split_heads(_linear(hidden, wq, bq) * scaling, num_heads)     # :167 (scale BEFORE the matmul)
    kh = _split_heads(_linear(hidden, wk, bk), num_heads)
    vh = _split_heads(_linear(hidden, wv, bv), num_heads)
    gate = None
    if gate_w is not None:
        gate = sigmoid(gate_logits_per_token(_split_heads(hidden, num_heads), gate_w, gate_b))
    ctx = attention_core(qh, kh, vh, scale=1.0, kind=kind, gamma=gamma, eta=eta, mask=attention_mask,
                         causal=causal, clamp_min=float(np.finfo(hidden.dtype).min),
                         gate=gate, gate_scale=gate_scale)
    b, h, t, _ = ctx.shape
    out = ctx.transpose(0, 2, 1, 3).reshape(b, t, h * d)               # :318-322
    return _linear(out, wo, bo)                                         # :324


# ---------------------------------------------------------------------------------------------
# percentile ranges, residual-sum tap, outlier metrics
# ---------------------------------------------------------------------------------------------

def percentile_linear(x: np.ndarray, q_lo: float, q_hi: float) -> Tuple[np.float64, np.float64]:
    """np.percentile(x, (q_lo, q_hi)) restated (numpy `_quantile`, method 'linear', called at
    quantization/range_estimators.py:91 with (100-p, p) and at :65-67 with (p, 100-p)): full sort instead of
    numpy's partition, then numpy's `_lerp` with the difference rounded to the array dtype.  Pinned against
    np.percentile itself in tests/test_oracle_golden.py (numpy is the reference's own implementation here)."""
    flat = np.sort(np.asarray(x).reshape(-1), kind="stable")      # NaN sorts last, like numpy's partition
    n = flat.size
    out = []
    for q in (q_lo, q_hi):
        virt = (n - 1) * (np.float64(q) / 100.0)
        prev = int(np.floor(virt))
        nxt = prev + 1
        prev_for_gamma = float(prev)
        if virt >= n - 1:                                          # _get_indexes: index -1
            prev = nxt = n - 1
            prev_for_gamma = -1.0
        if virt < 0:
            prev = nxt = 0
        a, b = flat[prev], flat[nxt]
        t = np.float64(virt - prev_for_gamma)
        with np.errstate(invalid="ignore", over="ignore"):
            d = np.float64(flat.dtype.type(b - a))                 # subtract(b, a) in the array dtype
            r = np.float64(b) - d * (1 - t) if t >= 0.5 else np.float64(a) + d * t
        if np.isnan(flat[-1]):
            r = np.float64("nan")
        out.append(np.float64(r))
    return out[0], out[1]


class RunningMinMaxPercentile:
    """RunningMinMaxEstimator, percentile branch (range_estimators.py:89-94, 99-104): float64 ranges."""

    def __init__(self, percentile: float, momentum: float = 0.9):
        self.percentile, self.momentum = percentile, momentum
        self.current_xmin = self.current_xmax = None

    def update(self, x: np.ndarray):
        lo, hi = percentile_linear(x, 100 - self.percentile, self.percentile)
        if self.current_xmin is None:
            self.current_xmin, self.current_xmax = lo, hi
        else:
            self.current_xmin = (1 - self.momentum) * lo + self.momentum * self.current_xmin
            self.current_xmax = (1 - self.momentum) * hi + self.momentum * self.current_xmax
        return self.current_xmin, self.current_xmax


def asym_set_quant_range_f64(x_min, x_max, n_bits: int = 8, eps: float = 1e-8):
    """set_quant_range on float64 0-dim tensors (uniform_quantizers.py:199-200, 218-219) -> (_delta, _zero_float,
    scale, zero_point) as float64; the quantize step then uses float32(scale), float32(zero_point) (ATen converts a
    0-dim operand to the dimensioned tensor's dtype)."""
    x_min = np.minimum(np.float64(x_min), 0.0)
    x_max = np.maximum(np.float64(x_max), np.float64(eps))
    int_max = 2.0 ** n_bits - 1
    delta = (x_max - x_min) / int_max
    zero_float = -x_min / delta
    scale = np.maximum(delta, eps)
    zp = np.clip(np.round(zero_float), 0.0, int_max)
    return delta, zero_float, scale, zp


def add_fake_quant(a: np.ndarray, b: np.ndarray, delta, zero_float, n_bits: int = 8, eps: float = 1e-8) -> np.ndarray:
    """res_act_quantizer(hidden_states + input_tensor) (quantized_bert.py:566-567, 602-603; quantized_opt.py:342-345)."""
    return asym_quantize(a + b, delta, zero_float, n_bits, eps)


def outlier_stats(x: np.ndarray, eps: float = 1e-6):
    """Per-sample inf-norm (validate_mlm_config.py:513-518) and kurtosis (transformers_language/utils.py:9-20) of x
    viewed as (B, -1), evaluated in float64 (the reference works in the tensor dtype; compared with a tolerance)."""
    x2 = np.asarray(x, dtype=np.float64).reshape(x.shape[0], -1)
    inf = np.abs(x2).max(axis=1)
    mu = x2.mean(axis=1, keepdims=True)
    s = x2.std(axis=1, ddof=1)                                     # torch.std: unbiased
    mu4 = ((x2 - mu) ** 4.0).mean(axis=1)
    return inf, mu4 / (s ** 4.0 + eps)


def ce_loss_mean(logits: np.ndarray, labels: np.ndarray, ignore_index: int = -100):
    """Per-batch eval loss: HF's CrossEntropyLoss()(prediction_scores.view(-1, V), labels.view(-1))
    behind `outputs.loss` at validate_mlm_config.py:489-496 / validate_clm.py:555-561 = mean over the rows whose label
    is not ignore_index of logsumexp(x_i) - x_i[label_i].  Evaluated in float64 on the given (rows, V) logits; returns
    (mean, per-row losses with 0 for ignored rows).  No labelled row -> NaN mean (ATen's 0/0)."""
    x = np.asarray(logits, dtype=np.float64)
    lab = np.asarray(labels, dtype=np.int64).reshape(-1)
    keep = lab != ignore_index
    m = x.max(axis=1, keepdims=True)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    lse = np.log(np.exp(x - m_safe).sum(axis=1)) + m_safe[:, 0]
    rows = np.zeros(x.shape[0], dtype=np.float64)
    idx = np.nonzero(keep)[0]
    rows[idx] = lse[idx] - x[idx, lab[idx]]
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = rows[idx].sum() / np.float64(idx.size) if idx.size else np.float64(np.nan)
    return mean, rows


def add_layer_norm(a: np.ndarray, b, weight: np.ndarray, bias, eps: float = 1e-12, q_res=None, q_out=None,
                   n_bits: int = 8) -> np.ndarray:
    """BertSelfOutput / BertOutput tail ``LayerNorm(hidden_states + input_tensor)`` (HF modeling_bert; wrapped by the
    reference at quantized_bert.py:540-605) and, with the taps, QuantizedBertSelfOutput.forward (:566-568): fp32
    restatement -- sum, optional res tap (delta, zero_float), biased-variance LayerNorm, optional output tap."""
    x = np.asarray(a, dtype=F32)
    if b is not None:
        x = (x + np.asarray(b, dtype=F32)).astype(F32)
    if q_res is not None:
        x = asym_quantize(x, q_res[0], q_res[1], n_bits)
    x64 = x.astype(np.float64)
    mu = x64.mean(-1, keepdims=True)
    var = x64.var(-1, keepdims=True)
    y = ((x64 - mu) / np.sqrt(var + eps)) * np.asarray(weight, dtype=np.float64)
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    y = y.astype(F32)
    if q_out is not None:
        y = asym_quantize(y, q_out[0], q_out[1], n_bits)
    return y
