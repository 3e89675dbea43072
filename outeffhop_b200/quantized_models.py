"""Model-level INT8 evaluation (SURVEY section 8a rows a13 + a15): the quantized BERT wrapper the ``--quantize`` branch of
validate_mlm_config.py builds (:445-476) and the range-estimation driver it runs.

* ``get_quant_config`` / ``val_qparams`` / ``DotDict``   -- transformers_language/quant_configs.py:7-33, utils.py:27-47, :74-91
* ``pass_data_for_range_estimation``                   -- transformers_language/utils.py:50-71
* ``QuantEmbedding``                                   -- quantization/autoquant_utils.py:73-90 (weights only, FP32Acts tap)
* ``QuantizedBertEmbeddings`` / ``QuantizedBertLayer`` / ``QuantizedBertModel`` / ``QuantizedBertForMaskedLM``
                                                       -- transformers_language/models/quantized_bert.py:146-218, :609-682,
                                                          :706-818, :857-926
The module tree keeps the reference's names (``bert.embeddings.sum_pos_embd_act_quantizer``,
``bert.encoder.layer.N.attention.self.query``, ``...intermediate.0``, ``...output.res_act_quantizer``) so quantizer state
can be compared / loaded name by name.  Two reference behaviours that follow from HF rather than from the paper and are
kept on purpose: the MLM head stays FP32 (``self.cls = org_model.cls``, :867), and the position of the FFN tap depends on
the activation MODULE: ``quantize_intermediate`` hands ``nn.Sequential(dense, act)`` to ``quantize_model``; HF's
``GELUActivation`` is not in ``activations_set``, so with a stock HF model the tap sits on the dense output BEFORE the GELU,
while the validate script's ``nn.GELU()`` replacement (validate_mlm_config.py:166-168; here ``hf_compat.fuse_bert_gelu``) is
fused into the QuantLinear and tapped AFTER the activation.

During estimation every tap runs minmax -> EMA -> qdq on the device (no host sync); with fixed ranges the attention core,
the residual blocks (add -> tap -> LayerNorm -> tap) and the loss are single kernels, and the step is CUDA-graph
capturable (``eval_harness.make_bert_int8_eval_step``).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .bert_attention import BertSelfAttentionWithExtras
from .quantization import (CurrentMinMaxEstimator, FP32Acts, QMethods, QuantizationManager, QuantizedActivation,
                           QuantizedModule, QuantLinear, RangeEstimators)
from .quantized_attention import QuantizedBertSelfAttentionWithExtras, QuantizedModel, _linear_params
from .quantized_layers import QuantizedBertOutput, QuantizedBertSelfOutput, QuantLayerNorm


class DotDict(dict):
    """transformers_language/utils.py:74-91: attribute access to dict entries (missing key -> None)."""

    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


def get_quant_config() -> DotDict:
    """transformers_language/quant_configs.py:7-33 (the defaults validate_mlm_config.py / validate_clm.py start from)."""
    config = DotDict()
    config.act_quant = DotDict({"cross_entropy_layer": None, "num_batches": 16, "options": {},
                                "quant_method": RangeEstimators.running_minmax, "std_dev": None})
    config.quant = DotDict({"act_quant": True, "n_bits": 8, "n_bits_act": 8, "num_candidates": None, "per_channel": False,
                            "percentile": None, "quant_setup": "all", "qmethod": QMethods.symmetric_uniform,
                            "qmethod_act": QMethods.asymmetric_uniform, "weight_quant": True,
                            "weight_quant_method": RangeEstimators.current_minmax})
    return config


def val_qparams(config) -> dict:
    """transformers_language/utils.py:27-47.  (The MSE weight-range search, ``weight_opt_method`` / ``num_candidates``,
    is outside the eval hot path: the reference's eval configs use current_minmax.)"""
    weight_range_options = {}
    if config.quant.num_candidates is not None:
        weight_range_options["num_candidates"] = config.quant.num_candidates
    return {
        "method": config.quant.qmethod.cls,
        "n_bits": config.quant.n_bits,
        "n_bits_act": config.quant.n_bits_act,
        "act_method": config.quant.qmethod_act.cls,
        "per_channel_weights": config.quant.per_channel,
        "percentile": config.quant.percentile,
        "quant_setup": config.quant.quant_setup,
        "weight_range_method": config.quant.weight_quant_method.cls,
        "weight_range_options": weight_range_options,
        "act_range_method": config.act_quant.quant_method.cls,
        "act_range_options": config.act_quant.options,
    }


def pass_data_for_range_estimation(loader: Iterable, model: nn.Module, act_quant=None, max_num_batches: int = 20,
                                   inp_idx: int = 0):
    """transformers_language/utils.py:50-71: ``max_num_batches`` forwards in eval / no_grad with every manager in its
    estimating state.  Dict batches are passed as keyword arguments, tuples by ``inp_idx``."""
    model.eval()
    batches = []
    device = next(model.parameters()).device
    with torch.no_grad():
        for i, data in enumerate(loader):
            if isinstance(data, (tuple, list)):
                x = data[inp_idx].to(device=device)
                batches.append(x.detach().cpu().numpy())
                model(x)
            else:
                model(**{k: v.to(device=device) for k, v in data.items()})
            if i >= max_num_batches - 1 or not act_quant:
                break
    return batches


class QuantEmbedding(QuantizedModule, nn.Embedding):
    """quantization/autoquant_utils.py:73-90: a lookup table is weights -- fake-quantized once (cached), never re-quantized
    as an activation (``activation_quantizer = FP32Acts()``)."""

    def __init__(self, *args, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        if activation is not None:
            raise NotImplementedError("embeddings have no fused activation")
        self.activation_function = None
        self.activation_quantizer = FP32Acts()
        if self.weight_range_method is CurrentMinMaxEstimator:
            weight_init_params = dict(percentile=self.percentile)
        else:
            weight_init_params = self.weight_range_options
        self.weight_quantizer = QuantizationManager(qmethod=self.method, init=self.weight_range_method,
                                                    per_channel=self.per_channel_weights, qparams=self.weight_qparams,
                                                    init_params=weight_init_params)

    @classmethod
    def from_embedding(cls, emb: nn.Embedding, **quant_params):
        q = cls(emb.num_embeddings, emb.embedding_dim, padding_idx=emb.padding_idx, max_norm=emb.max_norm,
                norm_type=emb.norm_type, scale_grad_by_freq=emb.scale_grad_by_freq, sparse=emb.sparse,
                **_linear_params(quant_params))
        q.weight = emb.weight
        return q.to(emb.weight.device)

    def get_params(self):                                          # hijacker.py:84-101
        if not self.training and self.cached_params:
            return self.cached_params
        weight = self.weight
        if self._qw_host:
            weight = self.weight_quantizer(weight)
        if self._caching and not self.training and self.cached_params is None:
            self.cached_params = (weight.detach().clone(), None)
        return weight, None

    def forward(self, x, offsets=None):
        weight, _ = self.get_params()
        return F.embedding(x.contiguous(), weight.contiguous(), self.padding_idx, self.max_norm, self.norm_type,
                           self.scale_grad_by_freq, self.sparse)


class QuantizedBertEmbeddings(QuantizedModel):
    """quantized_bert.py:146-218 (``quant_dict`` empty, as in the eval scripts)."""

    def __init__(self, org_model, **quant_params):
        super().__init__()
        qp = _linear_params(quant_params)
        self.word_embeddings = QuantEmbedding.from_embedding(org_model.word_embeddings, **qp)
        self.position_embeddings = QuantEmbedding.from_embedding(org_model.position_embeddings, **qp)
        self.token_type_embeddings = QuantEmbedding.from_embedding(org_model.token_type_embeddings, **qp)
        self.dropout = org_model.dropout
        position_ids = getattr(org_model, "position_ids", None)
        if position_ids is not None:
            self.register_buffer("position_ids", position_ids, persistent=False)
        else:
            self.position_ids = None
        self.position_embedding_type = getattr(org_model, "position_embedding_type", "absolute")
        self.sum_input_token_type_embd_act_quantizer = QuantizedActivation(**qp)
        self.sum_pos_embd_act_quantizer = QuantizedActivation(**qp)
        self.LayerNorm = QuantLayerNorm.from_layer_norm(org_model.LayerNorm, **qp)

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None):
        input_shape = input_ids.size() if input_ids is not None else inputs_embeds.size()[:-1]
        seq_length = input_shape[1]
        device = input_ids.device if input_ids is not None else inputs_embeds.device
        if position_ids is None:
            position_ids = self.position_ids[:, :seq_length]
        if token_type_ids is None:
            token_type_ids = torch.zeros(input_shape, dtype=torch.long, device=device)
        if inputs_embeds is None:
            inputs_embeds = self.word_embeddings(input_ids)
        embeddings = inputs_embeds + self.token_type_embeddings(token_type_ids)                 # :208-209
        embeddings = self.sum_input_token_type_embd_act_quantizer(embeddings)                   # :210
        if self.position_embedding_type == "absolute":
            embeddings = embeddings + self.position_embeddings(position_ids)                    # :212-214
            embeddings = self.sum_pos_embd_act_quantizer(embeddings)                            # :215
        embeddings = self.LayerNorm(embeddings)                                                 # :217
        return self.dropout(embeddings)


class _QuantizedBertAttention(nn.Module):
    """Container with HF's names: ``.self`` (quantized attention core) and ``.output`` (quantized residual block)."""

    def __init__(self, org_attention, **quant_params):
        super().__init__()
        core = org_attention.self
        if not isinstance(core, BertSelfAttentionWithExtras):
            raise NotImplementedError(
                "quantize the model after swap_bert_attention(): the wrapper table of quantized_bert.py:633-637 dispatches "
                f"on the attention class, got {type(core).__name__}")
        setattr(self, "self", QuantizedBertSelfAttentionWithExtras(core, **quant_params))
        self.output = QuantizedBertSelfOutput(org_attention.output, **quant_params)

    def forward(self, hidden_states, attention_mask=None, head_mask=None):
        self_out = getattr(self, "self")(hidden_states, attention_mask, head_mask)[0]
        return self.output(self_out, hidden_states)


class QuantizedBertLayer(QuantizedModel):
    """quantized_bert.py:620-682 (self-attention encoder layer; cross-attention / chunking are not used on the eval path)."""

    def __init__(self, org_model, **quant_params):
        super().__init__()
        if getattr(org_model, "add_cross_attention", False):
            raise NotImplementedError("cross-attention layers are outside the eval hot path")
        self.attention = _QuantizedBertAttention(org_model.attention, **quant_params)
        act = org_model.intermediate.intermediate_act_fn
        if not isinstance(act, nn.Module):
            if act is not F.gelu:
                raise NotImplementedError(f"intermediate activation {act} (quantized_bert.py:609-617)")
            act = nn.GELU()
        # quantize_intermediate (:609-617) hands Sequential(dense, act) to quantize_sequential: an activation from the
        # hijacker's activations_set (hijacker.py:14-24; nn.GELU included, also the nn.GELU() substituted for F.gelu) is
        # FUSED into the QuantLinear and runs BEFORE its tap (autoquant_utils.py:105-176, hijacker.py:115-117); any other
        # module (HF's GELUActivation is not an nn.GELU) stays a separate Sequential entry behind the tapped dense output.
        if isinstance(act, (nn.ReLU, nn.ReLU6, nn.Hardtanh, nn.Sigmoid, nn.Tanh, nn.GELU, nn.PReLU)):
            self.intermediate = nn.Sequential(QuantLinear.from_linear(org_model.intermediate.dense, activation=act,
                                                                      **_linear_params(quant_params)))
        else:
            self.intermediate = nn.Sequential(QuantLinear.from_linear(org_model.intermediate.dense,
                                                                      **_linear_params(quant_params)), act)
        self.output = QuantizedBertOutput(org_model.output, **quant_params)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, **unused):
        attention_output = self.attention(hidden_states, attention_mask, head_mask)
        intermediate_output = self.intermediate(attention_output)                               # :679
        return (self.output(intermediate_output, attention_output),)                            # :680


class _Encoder(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layer = nn.ModuleList(layers)


class QuantizedBertModel(QuantizedModel):
    """quantized_bert.py:706-818 without the pooler (BertForMaskedLM builds BertModel(add_pooling_layer=False))."""

    def __init__(self, org_model, **quant_params):
        super().__init__()
        self.config = org_model.config
        self.embeddings = QuantizedBertEmbeddings(org_model.embeddings, **quant_params)
        self.encoder = _Encoder([QuantizedBertLayer(l, **quant_params) for l in org_model.encoder.layer])
        if getattr(org_model, "pooler", None) is not None:
            raise NotImplementedError("pooler quantization (quantized_bert.py:684-703) is not on the MLM eval path")

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None):
        """``attention_mask``: None for the reference's chunked, un-padded eval data (validate_mlm_config.py:344-357), or
        an HF-4.31 extended additive mask (B,1,1,T) of 0 / finfo.min."""
        h = self.embeddings(input_ids=input_ids, token_type_ids=token_type_ids, position_ids=position_ids)
        for layer in self.encoder.layer:
            h = layer(h, attention_mask)[0]
        return h


class QuantizedBertForMaskedLM(QuantizedModel):
    """quantized_bert.py:857-926: quantized encoder, FP32 MLM head, mean CE over the labelled tokens."""

    def __init__(self, org_model, quant_setup=None, **quant_params):
        super().__init__()
        self.config = org_model.config
        quant_params.setdefault("quant_dict", {})
        self.bert = QuantizedBertModel(org_model.bert, **quant_params)
        self.cls = org_model.cls                                   # FP_logits setup (:866-870)

    def logits(self, input_ids, attention_mask=None, token_type_ids=None, position_ids=None):
        from .eval_harness import _padded_decoder
        h = self.bert(input_ids, attention_mask, token_type_ids, position_ids)
        pred = self.cls.predictions
        w, b, V = _padded_decoder(pred.decoder)
        logits = F.linear(pred.transform(h), w, b)
        return logits, V

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, labels=None, **unused):
        from transformers.modeling_outputs import MaskedLMOutput
        from . import ops
        logits, V = self.logits(input_ids, attention_mask, token_type_ids, position_ids)
        loss = None
        if labels is not None:
            loss = ops.ce_loss(logits.view(-1, logits.shape[-1]), labels.reshape(-1), vocab=V)
        return MaskedLMOutput(loss=loss, logits=logits[..., :V])


def quantize_bert_for_masked_lm(model, config: Optional[DotDict] = None, **overrides) -> QuantizedBertForMaskedLM:
    """validate_mlm_config.py:445-466: ``QuantizedBertForMaskedLM(model, **val_qparams(config))`` followed by
    ``set_quant_state(weight_quant, act_quant)``.  ``model``: an HF BertForMaskedLM whose attention modules are the
    drop-ins (``swap_bert_attention``)."""
    config = config or get_quant_config()
    qparams = val_qparams(config)
    qparams.update(overrides)
    qparams.setdefault("quant_dict", {})
    p = next(model.parameters())
    qmodel = QuantizedBertForMaskedLM(model, **qparams).to(p.device).eval()
    qmodel.set_quant_state(weight_quant=bool(config.quant.weight_quant), act_quant=bool(config.quant.act_quant))
    return qmodel


# =====================================================================================================================
# OPT (validate_clm.py --quantize; transformers_language/models/quantized_opt.py)
# =====================================================================================================================
from .opt_attention import OPTAttentionWithExtras                      # noqa: E402
from .quantized_attention import QuantizedOPTAttentionWithExtras       # noqa: E402
from .quantized_layers import _residual_tap                            # noqa: E402


class QuantizedOPTLearnedPositionalEmbedding(QuantizedModel):
    """quantized_opt.py:27-51: the learned position table (offset 2) as a weight-quantized lookup."""

    def __init__(self, org_model, **quant_params):
        super().__init__()
        self.offset = org_model.offset
        self.quant_embedding = QuantEmbedding.from_embedding(org_model, **quant_params)

    def forward(self, attention_mask: Optional[torch.Tensor] = None, past_key_values_length: int = 0,
                position_ids: Optional[torch.Tensor] = None):
        if position_ids is None:                                                                # :41-49
            attention_mask = attention_mask.long()
            position_ids = (torch.cumsum(attention_mask, dim=1).type_as(attention_mask) * attention_mask).long() - 1
            position_ids = position_ids[:, past_key_values_length:]
        return self.quant_embedding(position_ids + self.offset)


class QuantizedOPTDecoderLayer(QuantizedModel):
    """quantized_opt.py:277-384: [LN] -> attention -> residual add + tap -> [LN] -> fc1+ReLU (tap after the ReLU) -> fc2
    -> residual add + tap.  With fixed ranges each ``residual + x`` and its tap are one kernel (oeh_add_fake_quant)."""

    def __init__(self, org_model, **quant_params):
        super().__init__()
        qp = _linear_params(quant_params)
        self.embed_dim = org_model.embed_dim
        self.do_layer_norm_before = org_model.do_layer_norm_before
        self.dropout = org_model.dropout
        if not isinstance(org_model.activation_fn, nn.ReLU):
            raise NotImplementedError("quantized_opt.py:286 asserts a ReLU feed-forward activation")
        if not isinstance(org_model.self_attn, OPTAttentionWithExtras):
            raise NotImplementedError("quantize the model after swap_opt_attention() (quantized_opt.py:289 wraps "
                                      f"OPTAttentionWithExtras), got {type(org_model.self_attn).__name__}")
        self.self_attn = QuantizedOPTAttentionWithExtras(org_model.self_attn, **quant_params)
        self.self_attn_layer_norm = QuantLayerNorm.from_layer_norm(org_model.self_attn_layer_norm, **qp)
        self.fc1_act = QuantLinear.from_linear(org_model.fc1, activation=org_model.activation_fn, **qp)   # :292-294
        self.fc2 = QuantLinear.from_linear(org_model.fc2, **qp)
        self.final_layer_norm = QuantLayerNorm.from_layer_norm(org_model.final_layer_norm, **qp)
        self.self_attn_res_act_quantizer = QuantizedActivation(**qp)
        self.ffn_res_act_quantizer = QuantizedActivation(**qp)

    def _chainable(self, next_ln) -> bool:
        """``forward_chained`` applies: pre-LayerNorm layer in eval with every tap of the two residual blocks frozen."""
        if not self.do_layer_norm_before or self.training or not isinstance(next_ln, QuantLayerNorm):
            return False
        a = _chain_taps(self.final_layer_norm, (self.self_attn_res_act_quantizer,), (self.self_attn.out_proj,))
        b = _chain_taps(next_ln, (self.ffn_res_act_quantizer,), (self.fc2,))
        return a is not None and b is not None and self.final_layer_norm.weight.dtype == next_ln.weight.dtype

    def forward_chained(self, x, residual, next_ln, attention_mask=None):
        """The same layer for a caller that chains layers (QuantizedOPTDecoder.forward): ``x`` = this layer's LayerNorm-ed,
        tapped input (written by the previous layer's last kernel), ``residual`` = the tapped residual stream.  Two fused
        kernels replace 2 x (output tap, residual add + tap, LayerNorm + tap) = 6 launches (quantized_opt.py:338-370):
            out_proj tap -> + residual -> self_attn_res tap -> final_layer_norm -> its tap
            fc2 tap      -> + residual -> ffn_res tap       -> NEXT layer's self_attn_layer_norm (or the decoder's final one)
        Returns (next layer's input, new residual)."""
        attn = OPTAttentionWithExtras.forward(self.self_attn, hidden_states=x, attention_mask=attention_mask,
                                              _out_proj_linear_only=True)[0]
        out_proj, tap, ln = self.self_attn.out_proj, self.self_attn_res_act_quantizer, self.final_layer_norm
        n_bits = _chain_taps(ln, (tap,), (out_proj,))
        w, b = ln.get_params()
        x2, residual = ops.add_layer_norm(attn, residual, w, b, ln.eps, qparams_a=out_proj.tap_qparams(),
                                          qparams_res=tap.kernel_qparams() if tap.active else None,
                                          qparams_out=ln.activation_quantizer.quantizer.qparams if ln._qa_host else None,
                                          n_bits=n_bits, return_sum=True)
        shape = x2.shape
        f = self.fc1_act(x2.reshape(-1, shape[-1]))                                               # :360
        f = self.fc2.linear_only(f).view(shape)
        tap, ln = self.ffn_res_act_quantizer, next_ln
        n_bits = _chain_taps(ln, (tap,), (self.fc2,))
        w, b = ln.get_params()
        return ops.add_layer_norm(f, residual, w, b, ln.eps, qparams_a=self.fc2.tap_qparams(),
                                  qparams_res=tap.kernel_qparams() if tap.active else None,
                                  qparams_out=ln.activation_quantizer.quantizer.qparams if ln._qa_host else None,
                                  n_bits=n_bits, return_sum=True)

    def forward(self, hidden_states, attention_mask=None, layer_head_mask=None, output_attentions=False, use_cache=False,
                past_key_value=None, **unused):
        residual = hidden_states
        if self.do_layer_norm_before:
            hidden_states = self.self_attn_layer_norm(hidden_states)
        hidden_states, self_attn_weights, present_key_value = OPTAttentionWithExtras.forward(
            self.self_attn, hidden_states=hidden_states, past_key_value=past_key_value, attention_mask=attention_mask,
            layer_head_mask=layer_head_mask, output_attentions=output_attentions)
        hidden_states = F.dropout(hidden_states, p=self.dropout, training=self.training)
        hidden_states = _residual_tap(self.self_attn_res_act_quantizer, residual, hidden_states)        # :342-345
        if not self.do_layer_norm_before:
            hidden_states = self.self_attn_layer_norm(hidden_states)
        shape = hidden_states.shape
        hidden_states = hidden_states.reshape(-1, hidden_states.size(-1))
        residual = hidden_states
        if self.do_layer_norm_before:
            hidden_states = self.final_layer_norm(hidden_states)
        hidden_states = self.fc1_act(hidden_states)                                                      # :360
        tap, fc2 = self.ffn_res_act_quantizer, self.fc2
        if (not self.training and tap.fixed and fc2._qa_host and fc2.tap_fusable
                and fc2.activation_quantizer.n_bits == tap.activation_quantizer.n_bits):
            # fc2's output tap, the residual add and the residual tap in one kernel (:362-369)
            hidden_states = _residual_tap(tap, residual, fc2.linear_only(hidden_states),
                                          qparams_b=fc2.tap_qparams()).view(shape)
        else:
            hidden_states = fc2(hidden_states)
            hidden_states = F.dropout(hidden_states, p=self.dropout, training=self.training)
            hidden_states = _residual_tap(tap, residual, hidden_states).view(shape)                      # :367-370
        if not self.do_layer_norm_before:
            hidden_states = self.final_layer_norm(hidden_states)
        outputs = (hidden_states,)
        if output_attentions:
            outputs += (self_attn_weights,)
        if use_cache:
            outputs += (present_key_value,)
        return outputs


def _chain_taps(layer_norm: "QuantLayerNorm", taps, linears):
    """n_bits shared by a fused ``dense tap -> add -> residual tap -> LayerNorm -> output tap`` kernel, or None when the
    pieces cannot share one launch (a tap still estimating, observers attached, mixed bit widths)."""
    from .quantization import AsymmetricUniformQuantizer
    bits = set()
    for t in taps:
        if t.active:
            if not t.fixed or t._forward_hooks or t._forward_pre_hooks:
                return None
            bits.add(t.activation_quantizer.n_bits)
    for lin in linears:
        if not lin.tap_fusable:
            return None
        if lin._qa_host:
            bits.add(lin.activation_quantizer.n_bits)
    ln = layer_norm
    out = ln.activation_quantizer
    if ln._forward_hooks or ln._forward_pre_hooks or ln.bias is None or len(ln.normalized_shape) != 1:
        return None
    if ln._qa_host:
        if out.estimating or not isinstance(out.quantizer, AsymmetricUniformQuantizer):
            return None
        bits.add(out.n_bits)
    if len(bits) > 1:
        return None
    return bits.pop() if bits else 8


class QuantizedOPTDecoder(QuantizedModel):
    """quantized_opt.py:387-662 for the eval path (no KV cache, no head masks)."""

    def __init__(self, org_model, **quant_params):
        super().__init__()
        qp = _linear_params(quant_params)
        self.config = org_model.config
        self.dropout = org_model.dropout
        self.embed_tokens = QuantEmbedding.from_embedding(org_model.embed_tokens, **qp)
        self.embed_positions = QuantizedOPTLearnedPositionalEmbedding(org_model.embed_positions, **qp)
        self.project_out = None if org_model.project_out is None else QuantLinear.from_linear(org_model.project_out, **qp)
        self.project_in = None if org_model.project_in is None else QuantLinear.from_linear(org_model.project_in, **qp)
        self.final_layer_norm = (None if org_model.final_layer_norm is None
                                 else QuantLayerNorm.from_layer_norm(org_model.final_layer_norm, **qp))
        self.layers = nn.ModuleList([QuantizedOPTDecoderLayer(l, **quant_params) for l in org_model.layers])
        self.embed_sum_act_quantizer = QuantizedActivation(**qp)

    def forward(self, input_ids, attention_mask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None):
        """``attention_mask``: the (B,1,T,T) additive causal mask of HF-4.31's ``_prepare_decoder_attention_mask``
        (``eval_harness.make_causal_mask``) -- the eval data is un-padded -- or None for a single token."""
        tok, posq = self.embed_tokens, self.embed_positions.quant_embedding
        if (position_ids is not None and self.project_in is None and not self.training and input_ids.is_cuda
                and input_ids.dim() == 2 and tok.max_norm is None and posq.max_norm is None
                and not (tok._forward_hooks or posq._forward_hooks or self.embed_positions._forward_hooks)):
            # the two (weight-quantized, cached) tables gathered and summed in ONE launch (oeh_embed_sum)
            from . import ops
            w_tok, w_pos = tok.get_params()[0], posq.get_params()[0]
            emb_sum = ops.embed_sum(input_ids, w_tok.contiguous(), w_pos.contiguous(), position_ids, self.embed_positions.offset)
        else:
            inputs_embeds = tok(input_ids)
            if position_ids is None:
                ones = torch.ones(inputs_embeds.shape[:2], dtype=torch.bool, device=inputs_embeds.device)
                pos_embeds = self.embed_positions(ones, 0)
            else:
                pos_embeds = self.embed_positions(position_ids=position_ids)
            if self.project_in is not None:
                inputs_embeds = self.project_in(inputs_embeds)
            emb_sum = inputs_embeds + pos_embeds
        hidden_states = self.embed_sum_act_quantizer(emb_sum)                                    # :556-559
        layers = list(self.layers)
        if (layers and self.final_layer_norm is not None and not self.training and hidden_states.is_cuda
                and hidden_states.dtype in (torch.float16, torch.bfloat16, torch.float32)
                and all(l._chainable(nxt) for l, nxt in zip(layers, [l.self_attn_layer_norm for l in layers[1:]]
                                                            + [self.final_layer_norm]))):
            # Fixed ranges, pre-LayerNorm layers: every ``residual + x`` with its two taps shares ONE kernel with the
            # LayerNorm (and output tap) that follows it -- across layer boundaries too; the last one is final_layer_norm.
            x = layers[0].self_attn_layer_norm(hidden_states)
            for i, layer in enumerate(layers):
                nxt = layers[i + 1].self_attn_layer_norm if i + 1 < len(layers) else self.final_layer_norm
                x, hidden_states = layer.forward_chained(x, hidden_states, nxt, attention_mask)
            hidden_states = x
        else:
            for layer in layers:
                hidden_states = layer(hidden_states, attention_mask=attention_mask)[0]
            if self.final_layer_norm is not None:
                hidden_states = self.final_layer_norm(hidden_states)
        if self.project_out is not None:
            hidden_states = self.project_out(hidden_states)
        return hidden_states


class _QuantizedOPTModel(QuantizedModel):
    def __init__(self, org_model, **quant_params):
        super().__init__()
        self.decoder = QuantizedOPTDecoder(org_model.decoder, **quant_params)


class QuantizedOPTForCausalLM(QuantizedModel):
    """quantized_opt.py:719-889 with the ``quant_setup`` choices the eval scripts use: ``fp32_head`` (FP lm_head, the
    README's OPT recipe), ``fp32_head+int16_final_LN`` / ``fp32_head+int16_all_LN`` (16-bit LayerNorm quantizers) and
    ``all`` (lm_head quantized too).  Loss: shifted mean CE (HF OPTForCausalLM), evaluated by ``oeh_ce_loss``."""

    def __init__(self, org_model, quant_setup="fp32_head", **quant_params):
        super().__init__()
        self.config = org_model.config
        self.model = _QuantizedOPTModel(org_model.model, **quant_params)
        if quant_setup in ("fp32_head", "fp32_head+int16_final_LN", "fp32_head+int16_all_LN"):
            self.lm_head = org_model.lm_head
            lns = []
            if quant_setup == "fp32_head+int16_final_LN":
                lns = [self.model.decoder.final_layer_norm]
            elif quant_setup == "fp32_head+int16_all_LN":
                lns = [m for m in self.model.modules() if isinstance(m, QuantLayerNorm)]
            for ln in lns:
                ln.weight_quantizer.quantizer.n_bits = 16
                ln.activation_quantizer.quantizer.n_bits = 16
        elif quant_setup == "all":
            self.lm_head = QuantLinear.from_linear(org_model.lm_head, **_linear_params(quant_params))
        else:
            raise ValueError(f"Quantization setup '{quant_setup}' not supported.")

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, labels=None, **unused):
        from transformers.modeling_outputs import CausalLMOutputWithPast
        from . import ops
        from .eval_harness import _padded_decoder, make_causal_mask
        if attention_mask is None and input_ids.shape[1] > 1:
            p = next(self.parameters())
            attention_mask = make_causal_mask(input_ids.shape[0], input_ids.shape[1], p.dtype, input_ids.device)
        h = self.model.decoder(input_ids, attention_mask, position_ids)
        if isinstance(self.lm_head, QuantLinear):
            logits, V = self.lm_head(h), self.lm_head.out_features
        else:
            w, b, V = _padded_decoder(self.lm_head)
            logits = F.linear(h, w, b)
        loss = None
        if labels is not None:
            # HF's shift of logits / labels, applied to the label index inside the kernel
            loss = ops.ce_loss(logits.view(-1, logits.shape[-1]), labels.reshape(-1), vocab=V, shift_block=labels.shape[-1])
        return CausalLMOutputWithPast(loss=loss, logits=logits[..., :V])


def quantize_opt_for_causal_lm(model, config: Optional[DotDict] = None, quant_setup: str = "fp32_head",
                               **overrides) -> QuantizedOPTForCausalLM:
    """validate_clm.py:465-487: ``QuantizedOPTForCausalLM(model, **val_qparams(config))`` + ``set_quant_state``."""
    config = config or get_quant_config()
    qparams = val_qparams(config)
    qparams.update(overrides)
    qparams.pop("quant_setup", None)
    p = next(model.parameters())
    qmodel = QuantizedOPTForCausalLM(model, quant_setup=quant_setup, **qparams).to(p.device).eval()
    qmodel.set_quant_state(weight_quant=bool(config.quant.weight_quant), act_quant=bool(config.quant.act_quant))
    return qmodel
