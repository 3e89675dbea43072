"""Quantized attention wrappers: ``QuantizedBertSelfAttentionWithExtras``
(transformers_language/models/quantized_bert.py:221-440) and ``QuantizedOPTAttentionWithExtras``
(transformers_language/models/quantized_opt.py:54-274), plus the small ``QuantizedModel`` base
(quantization/base_quantized_model.py) they need.

Construction mirrors the reference: attributes are copied from the FP module, q/k/v(/out) projections
become ``QuantLinear`` (cached symmetric weight fake-quant + output activation tap) and three
``QuantizedActivation`` taps are added: scores after the scale and BEFORE the mask, probs after the
softmax, context after the gate (BERT, :434) or before it (OPT, :210).  ``gate_scaling_factor`` is dropped
by the BERT wrapper exactly like the reference (:422).

While any tap is estimating its range (``estimate_ranges``) the forward runs the unfused sequence: every
tap needs a global min/max of its input before that input is quantized.  Once the ranges are fixed the
whole chain -- three taps included -- is one fused tcgen05 kernel launch.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .bert_attention import BertSelfAttentionWithExtras
from .opt_attention import OPTAttentionWithExtras
from .quantization import (QuantizedActivation, QuantizedModule, QuantLinear, _set_layer_estimate_ranges,
                           _set_layer_estimate_ranges_train, _set_layer_fix_ranges, estimate_and_quantize_group)


class QuantizedModel(nn.Module):
    """State toggles of quantization/base_quantized_model.py:84-155 (the dummy-forward load_state_dict of
    :34-82 is unnecessary here: quantizer state tensors have a fixed shape)."""

    def quantized_weights(self):
        self.apply(lambda l: l.quantized_weights() if isinstance(l, QuantizedModule) else None)

    def full_precision_weights(self):
        self.apply(lambda l: l.full_precision_weights() if isinstance(l, QuantizedModule) else None)

    def quantized_acts(self):
        self.apply(lambda l: l.quantized_acts() if isinstance(l, QuantizedModule) else None)

    def full_precision_acts(self):
        self.apply(lambda l: l.full_precision_acts() if isinstance(l, QuantizedModule) else None)

    def quantized(self):
        self.apply(lambda l: l.quantized() if isinstance(l, QuantizedModule) else None)

    def full_precision(self):
        self.apply(lambda l: l.full_precision() if isinstance(l, QuantizedModule) else None)

    def fix_ranges(self):
        self.apply(_set_layer_fix_ranges)

    def estimate_ranges(self):
        self.apply(_set_layer_estimate_ranges)

    def estimate_ranges_train(self):
        self.apply(_set_layer_estimate_ranges_train)

    def learn_ranges(self):
        raise NotImplementedError("learn_ranges is outside the eval hot path")

    def set_quant_state(self, weight_quant, act_quant):
        (self.quantized_acts if act_quant else self.full_precision_acts)()
        (self.quantized_weights if weight_quant else self.full_precision_weights)()

    def reset_act_ranges(self):
        from .quantization import QuantizationManager

        def _fn(module):
            if isinstance(module, QuantizationManager) and module.quantizer.act_quant:
                module.reset_ranges()
        self.apply(_fn)


def _linear_params(quant_params):
    qp = dict(quant_params)
    qp.pop("quant_dict", None)
    return qp


def quantize_linear(lin: nn.Linear, **quant_params) -> QuantLinear:
    """quantize_model() restricted to nn.Linear (autoquant_utils.py:236-270 -> QuantLinear)."""
    return QuantLinear.from_linear(lin, **_linear_params(quant_params))


def quantized_fused_qkv(owner: nn.Module, mods, hidden_states: torch.Tensor, num_heads: int, head_dim: int):
    """Fixed ranges: the three QuantLinear projections as ONE GEMM over the concatenated (cached) fake-quantized weights and
    ONE segmented tap (each projection keeps its own range: ``oeh_fake_quant_asym_segments``) instead of 3 + 3 launches.
    Returns (q, k, v) as (B,H,T,d) strided views, or None while any tap is still estimating / taps disagree."""
    if (not owner.training and all(isinstance(m, QuantLinear) and m._qa_host and m.activation_function is None
                                   and not m._forward_hooks and not m._forward_pre_hooks for m in mods)
            and all(m.activation_quantizer.estimating for m in mods)):
        # Range estimation: the three projection taps see independent tensors, so their batch (min, max) pairs cross the
        # ranks in ONE all-reduce (estimate_and_quantize_group); each tap then updates / quantizes as on its own.
        ys = [m.linear_only(hidden_states) for m in mods]
        outs = estimate_and_quantize_group([m.activation_quantizer for m in mods], ys)
        if outs is None:
            return None
        B, T = hidden_states.shape[0], hidden_states.shape[1]
        return tuple(y.view(B, T, num_heads, head_dim).permute(0, 2, 1, 3) for y in outs)
    if owner.training or not all(isinstance(m, QuantLinear) and m.tap_fusable and m.activation_function is None
                                 and not m._forward_pre_hooks for m in mods):
        return None
    active = [m._qa_host for m in mods]
    if any(active) != all(active) or len({m.activation_quantizer.n_bits for m in mods}) != 1 \
            or any((m.bias is None) != (mods[0].bias is None) for m in mods):
        return None
    params = [m.get_params() for m in mods]                   # cached fake-quantized weights (hijacker.py:84-101)
    key = tuple((id(w), None if b is None else id(b)) for w, b in params)
    cache = owner.__dict__.get("_qkv_cache")
    if cache is None or cache[0] != key:
        with torch.no_grad():
            w = torch.cat([w for w, _ in params], dim=0)
            b = None if params[0][1] is None else torch.cat([b for _, b in params], dim=0)
        cache = (key, w, b, params)                           # `params` keeps the keyed tensors alive
        owner.__dict__["_qkv_cache"] = cache
    seg = cache[1].shape[0] // 3
    if (all(active) and hidden_states.dtype in (torch.float16, torch.bfloat16) and cache[1].dtype == hidden_states.dtype
            and seg % 128 == 0 and hidden_states.shape[-1] % 8 == 0 and not torch.is_grad_enabled()):
        # the three projections AND their three taps in one GEMM launch (per-segment ranges in the epilogue)
        y = ops.linear_fwd(hidden_states, cache[1], cache[2], qparams=[m.tap_qparams() for m in mods],
                           n_bits=mods[0].activation_quantizer.n_bits)
    else:
        y = nn.functional.linear(hidden_states, cache[1], cache[2])
        if all(active):
            y = ops.fake_quant_asym_segments(y, [m.tap_qparams() for m in mods], mods[0].activation_quantizer.n_bits)
    B, T = y.shape[0], y.shape[1]
    y = y.view(B, T, 3, num_heads, head_dim)
    return tuple(y[:, :, j].permute(0, 2, 1, 3) for j in range(3))


class QuantizedBertSelfAttentionWithExtras(QuantizedModel, BertSelfAttentionWithExtras):
    def __init__(self, org_model, **quant_params):
        nn.Module.__init__(self)
        self.quant_dict = dict(quant_params.get("quant_dict") or {})
        qp = _linear_params(quant_params)
        # copy attributes (quantized_bert.py:227-258)
        self.num_attention_heads = org_model.num_attention_heads
        self.attention_head_size = org_model.attention_head_size
        self.all_head_size = org_model.all_head_size
        self.position_embedding_type = getattr(org_model, "position_embedding_type", None)
        if self.position_embedding_type in ("relative_key", "relative_key_query"):
            self.max_position_embeddings = org_model.max_position_embeddings
            self.distance_embedding = org_model.distance_embedding
        self.is_decoder = org_model.is_decoder
        self.query = quantize_linear(org_model.query, **qp)
        self.key = quantize_linear(org_model.key, **qp)
        self.value = quantize_linear(org_model.value, **qp)
        self.dropout = org_model.dropout
        self.attn_scores_act_quantizer = QuantizedActivation(**qp)
        self.attn_probs_act_quantizer = QuantizedActivation(**qp)
        self.context_act_quantizer = QuantizedActivation(**qp)
        self.softmax_fn = org_model.softmax_fn
        self.attn_gate_type = org_model.attn_gate_type
        self.attn_gate_init = org_model.attn_gate_init
        self.attn_gate_mlp = org_model.attn_gate_mlp
        self.attn_gate_mlp2 = org_model.attn_gate_mlp2
        self.attn_gate_linear_all_features = org_model.attn_gate_linear_all_features
        self.alpha = org_model.alpha            # do not quantize for now (:256)
        self.gate_fn = org_model.gate_fn
        self.pooling_fn = org_model.pooling_fn
        # probes / flags the shared forward expects
        self.attn_scores = nn.Identity()
        self.attn_probs_before_dropout = nn.Identity()
        self.attn_probs_after_dropout = nn.Identity()
        self.skip_attn = False
        self.materialize_probs = False
        self.fused_fp32 = bool(getattr(org_model, "fused_fp32", False))
        self.gate_scaling_factor = 1.0
        self._gate_pack = None
        self._last_gate_logits = None
        self.last_gate_avg_prob_ = None

    def _taps(self):
        return (self.attn_scores_act_quantizer, self.attn_probs_act_quantizer, self.context_act_quantizer)

    def _fused_qkv(self, hidden_states):
        return quantized_fused_qkv(self, (self.query, self.key, self.value), hidden_states, self.num_attention_heads,
                                   self.attention_head_size)

    def _gate_scale(self) -> float:
        return 1.0                                # ``context_layer *= gate`` (quantized_bert.py:422)


class QuantizedOPTAttentionWithExtras(QuantizedModel, OPTAttentionWithExtras):
    def __init__(self, org_model, **quant_params):
        nn.Module.__init__(self)
        qp = _linear_params(quant_params)
        # copy attributes (quantized_opt.py:58-89)
        self.embed_dim = org_model.embed_dim
        self.num_heads = org_model.num_heads
        self.dropout = org_model.dropout
        self.head_dim = org_model.head_dim
        self.scaling = org_model.scaling
        self.is_decoder = org_model.is_decoder
        self.k_proj = quantize_linear(org_model.k_proj, **qp)
        self.v_proj = quantize_linear(org_model.v_proj, **qp)
        self.q_proj = quantize_linear(org_model.q_proj, **qp)
        self.out_proj = quantize_linear(org_model.out_proj, **qp)
        self.attn_scores_act_quantizer = QuantizedActivation(**qp)
        self.attn_probs_act_quantizer = QuantizedActivation(**qp)
        self.context_act_quantizer = QuantizedActivation(**qp)
        self.softmax_fn = org_model.softmax_fn
        self.attn_gate_type = org_model.attn_gate_type
        self.attn_gate_init = org_model.attn_gate_init
        self.attn_gate_mlp = org_model.attn_gate_mlp
        self.attn_gate_mlp2 = org_model.attn_gate_mlp2
        self.attn_gate_linear_all_features = org_model.attn_gate_linear_all_features
        self.alpha = org_model.alpha
        self.gate_fn = org_model.gate_fn
        self.pooling_fn = org_model.pooling_fn
        self.attn_scores = nn.Identity()
        self.attn_probs_before_dropout = nn.Identity()
        self.attn_probs_after_dropout = nn.Identity()
        self.skip_attn = False
        self.materialize_probs = False
        self.fused_fp32 = bool(getattr(org_model, "fused_fp32", False))
        self.causal_hint = getattr(org_model, "causal_hint", False)
        self.gate_scaling_factor = 1.0            # ``attn_output *= gate`` (quantized_opt.py:250)
        self._gate_pack = None
        self._last_gate_logits = None
        self.last_gate_avg_prob_ = None

    def _taps(self):
        return (self.attn_scores_act_quantizer, self.attn_probs_act_quantizer, self.context_act_quantizer)

    def _fused_qkv(self, hidden_states):
        return quantized_fused_qkv(self, (self.q_proj, self.k_proj, self.v_proj), hidden_states, self.num_heads, self.head_dim)
