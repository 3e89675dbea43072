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
kept on purpose: the MLM head stays FP32 (``self.cls = org_model.cls``, :867), and the FFN tap sits on the dense output
BEFORE the GELU (``quantize_intermediate`` hands ``nn.Sequential(dense, GELUActivation())`` to ``quantize_model``;
HF's ``GELUActivation`` is not in ``activations_set``, so the activation is neither fused into the QuantLinear nor tapped).

During estimation every tap runs minmax -> EMA -> qdq on the device (no host sync); with fixed ranges the attention core,
the residual blocks (add -> tap -> LayerNorm -> tap) and the loss are single kernels, and the step is CUDA-graph
capturable (``eval_harness.make_bert_int8_eval_step``).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
from torch import nn
from torch.nn import functional as F

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
        # quantize_intermediate (:609-617): QuantLinear tapped on the dense output, then the activation module as is
        self.intermediate = nn.Sequential(QuantLinear.from_linear(org_model.intermediate.dense, **_linear_params(quant_params)),
                                          act)
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
