"""Glue between the drop-in attention modules and the HuggingFace model shells they are swapped into.

The reference targets transformers 4.31 (requirement.txt:47): ``BertAttention`` calls
``self.self(hidden_states, attention_mask, head_mask, encoder_hidden_states, encoder_attention_mask,
past_key_value, output_attentions)`` and expects the reference's variable-length tuple.  transformers >= 4.48
calls ``self.self(hidden_states, attention_mask=..., past_key_values=..., **kw)`` and unpacks exactly
``(attn_output, attn_weights)``; ``OPTDecoderLayer`` likewise unpacks two values.  The ``*HF`` subclasses
below keep the reference constructors / state_dict keys and only adapt the call convention, so the
module-swap loops of validate_mlm_config.py:174-191 and validate_clm.py:147-168 work on either version:

    swap_bert_attention(model, attn_softmax="softmax1", attn_gate_type="conditional_per_token", ...)
    swap_opt_attention(model, attn_softmax="vanilla", ...)
"""
from __future__ import annotations

from typing import Optional

import torch
import transformers
from packaging import version

from .bert_attention import AttentionGateType, BertSelfAttentionWithExtras
from .opt_attention import OPTAttentionWithExtras
from .softmax import parse_attn_softmax

NEW_STYLE = version.parse(transformers.__version__) >= version.parse("4.48.0")


class BertSelfAttentionWithExtrasHF(BertSelfAttentionWithExtras):
    """Call-convention adapter for transformers >= 4.48 BertAttention."""

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False, **kw):
        if kw.get("past_key_values") is not None:
            raise NotImplementedError("KV caches are not used on the evaluation path")
        out = super().forward(hidden_states, attention_mask, head_mask, encoder_hidden_states,
                              encoder_attention_mask, past_key_value, bool(output_attentions))
        if not NEW_STYLE:
            return out
        return out[0], (out[1] if output_attentions and len(out) > 1 else None)


class OPTAttentionWithExtrasHF(OPTAttentionWithExtras):
    """Call-convention adapter for transformers >= 4.48 OPTDecoderLayer."""

    def forward(self, hidden_states, key_value_states=None, past_key_value=None, attention_mask=None,
                layer_head_mask=None, output_attentions=False, **kw):
        if kw.get("past_key_values") is not None:
            raise NotImplementedError("KV caches are not used on the evaluation path (pass use_cache=False)")
        extra = {k: kw[k] for k in ("_gate_logits", "_out_proj_linear_only") if k in kw}
        out, weights, past = super().forward(hidden_states, key_value_states, past_key_value, attention_mask,
                                             layer_head_mask, bool(output_attentions), **extra)
        if not NEW_STYLE:
            return out, weights, past
        return out, weights


def _gate_type(attn_gate_type):
    return AttentionGateType[attn_gate_type] if isinstance(attn_gate_type, str) else attn_gate_type


def swap_bert_attention(model, attn_softmax: str = "vanilla", alpha: Optional[float] = None,
                        max_seq_length: Optional[int] = None, skip_attn: bool = False,
                        attn_gate_type="none", attn_gate_init: Optional[float] = None, attn_gate_mlp: bool = False,
                        attn_gate_mlp2: bool = False, attn_gate_linear_all_features: bool = False,
                        fine_tuning: bool = False, fused_fp32: bool = False):
    """validate_mlm_config.py:174-191: replace every ``layer.attention.self`` and copy its weights.  ``fused_fp32``: let fp32
    activations take the fused kernel with fp16 operands (bert_attention._fusable_dtype; default: exact unfused fp32)."""
    softmax_fn = parse_attn_softmax(attn_softmax)
    bert = getattr(model, "bert", model)
    for layer in bert.encoder.layer:
        old = layer.attention.self
        new = BertSelfAttentionWithExtrasHF(
            model.config, softmax_fn=softmax_fn, alpha=alpha, max_seq_length=max_seq_length, skip_attn=skip_attn,
            attn_gate_type=_gate_type(attn_gate_type), attn_gate_init=attn_gate_init, attn_gate_mlp=attn_gate_mlp,
            attn_gate_mlp2=attn_gate_mlp2, attn_gate_linear_all_features=attn_gate_linear_all_features,
            fine_tuning=fine_tuning)
        new.load_state_dict(old.state_dict(), strict=False)
        p = next(old.parameters())
        new.fused_fp32 = bool(fused_fp32)
        layer.attention.self = new.to(device=p.device, dtype=p.dtype).train(old.training)
    return model


def swap_opt_attention(model, attn_softmax: str = "vanilla", alpha: Optional[float] = None,
                       max_seq_length: Optional[int] = None, skip_attn: bool = False, attn_gate_type="none",
                       attn_gate_init: Optional[float] = None, attn_gate_mlp: bool = False,
                       attn_gate_mlp2: bool = False, attn_gate_linear_all_features: bool = False,
                       fine_tuning: bool = False, causal_hint: bool = False, fused_fp32: bool = False):
    """validate_clm.py:147-168: replace every ``layer.self_attn`` and copy its weights (``fused_fp32``: see swap_bert_attention)."""
    softmax_fn = parse_attn_softmax(attn_softmax)
    cfg = model.config
    for layer in model.model.decoder.layers:
        old = layer.self_attn
        new = OPTAttentionWithExtrasHF(
            embed_dim=cfg.hidden_size, num_heads=cfg.num_attention_heads, dropout=cfg.attention_dropout,
            is_decoder=True, bias=cfg.enable_bias, softmax_fn=softmax_fn, alpha=alpha, max_seq_length=max_seq_length,
            skip_attn=skip_attn, attn_gate_type=_gate_type(attn_gate_type), attn_gate_init=attn_gate_init,
            attn_gate_mlp=attn_gate_mlp, attn_gate_mlp2=attn_gate_mlp2,
            attn_gate_linear_all_features=attn_gate_linear_all_features, fine_tuning=fine_tuning)
        new.load_state_dict(old.state_dict(), strict=False)
        new.causal_hint = causal_hint
        p = next(old.parameters())
        new.fused_fp32 = bool(fused_fp32)
        layer.self_attn = new.to(device=p.device, dtype=p.dtype).train(old.training)
    return model


class BertResidualLayerNormFused(torch.nn.Module):
    """BertSelfOutput / BertOutput (the HF modules the reference wraps at quantized_bert.py:540-605) with the residual
    add and the LayerNorm in one kernel: ``LayerNorm(dropout(dense(hidden_states)) + input_tensor)``.  Same submodule
    names (dense, LayerNorm, dropout), so state_dicts and the quantized wrappers keep working."""

    def __init__(self, org):
        super().__init__()
        self.dense, self.LayerNorm, self.dropout = org.dense, org.LayerNorm, org.dropout

    def forward(self, hidden_states, input_tensor):
        hidden_states = self.dropout(self.dense(hidden_states))
        ln = self.LayerNorm
        if self.training or (torch.is_grad_enabled() and hidden_states.requires_grad) \
                or ln._forward_hooks or ln._forward_pre_hooks:
            # training graph (autograd needs the ATen ops) / somebody observes the LayerNorm module itself (the
            # validate scripts hook named modules for the outlier metrics, validate_mlm_config.py:501-540)
            return ln(hidden_states + input_tensor)
        from . import ops
        return ops.add_layer_norm(hidden_states, input_tensor, ln.weight, ln.bias, ln.eps)


def fuse_bert_residual_blocks(model):
    """Swap every ``layer.attention.output`` / ``layer.output`` of a BERT encoder for the fused module (eval only)."""
    bert = getattr(model, "bert", model)
    for layer in bert.encoder.layer:
        layer.attention.output = BertResidualLayerNormFused(layer.attention.output).train(layer.training)
        layer.output = BertResidualLayerNormFused(layer.output).train(layer.training)
    return model


class GELUFused(torch.nn.GELU):
    """The erf GELU between the two FFN GEMMs through ``oeh_gelu_erf`` for 16-bit CUDA tensors in eval; anything else
    (training graphs, fp32, hooks on this module) takes ATen's kernel.  An ``nn.GELU`` subclass on purpose: the validate
    script replaces HF's ``GELUActivation`` with ``nn.GELU()`` before anything else (validate_mlm_config.py:166-168), and
    ``nn.GELU`` is in the hijacker's ``activations_set`` (hijacker.py:14-24), so ``quantize_intermediate``
    (quantized_bert.py:609-617) fuses it into the QuantLinear and taps AFTER the activation -- a model prepared with
    ``fuse_bert_gelu`` quantizes the same way."""

    def __init__(self):
        super().__init__(approximate="none")

    def forward(self, input):
        if self.training or (torch.is_grad_enabled() and input.requires_grad) or not input.is_cuda \
                or input.dtype not in (torch.float16, torch.bfloat16) or self._forward_hooks or self._forward_pre_hooks:
            return torch.nn.functional.gelu(input)
        from . import ops
        return ops.gelu(input)


def _is_erf_gelu(act) -> bool:
    if act is torch.nn.functional.gelu:
        return True
    if isinstance(act, torch.nn.GELU):
        return act.approximate == "none"
    return type(act).__name__ == "GELUActivation" and getattr(act, "act", None) is torch.nn.functional.gelu


class BertIntermediateFused(torch.nn.Module):
    """HF ``BertIntermediate`` (dense -> erf GELU; the pair the reference hands to ``quantize_intermediate``,
    quantized_bert.py:609-617) as ONE launch: the GELU runs in the epilogue of the library's tcgen05 GEMM (``oeh_linear_fwd``)
    for 16-bit CUDA tensors in eval; anything else (training graphs, fp32, hooks on either submodule) runs the two modules.
    Same submodule names (``dense``, ``intermediate_act_fn``)."""

    def __init__(self, org):
        super().__init__()
        self.dense, self.intermediate_act_fn = org.dense, org.intermediate_act_fn

    def forward(self, hidden_states):
        d, act = self.dense, self.intermediate_act_fn
        if (not self.training and hidden_states.is_cuda and hidden_states.dtype in (torch.float16, torch.bfloat16)
                and type(d) is torch.nn.Linear and d.weight.dtype == hidden_states.dtype and isinstance(act, GELUFused)
                and not (torch.is_grad_enabled() and (hidden_states.requires_grad or d.weight.requires_grad))
                and not any(m._forward_hooks or m._forward_pre_hooks for m in (d, act))
                and d.out_features % 8 == 0 and d.in_features % 8 == 0):
            from . import ops
            return ops.linear_fwd(hidden_states, d.weight, d.bias, act="gelu")
        return act(d(hidden_states))


def fuse_bert_gelu(model):
    """Swap the erf-GELU activation of every ``layer.intermediate`` for :class:`GELUFused` (other activations stay) and the
    ``BertIntermediate`` module itself for :class:`BertIntermediateFused` (dense + GELU in one GEMM launch)."""
    bert = getattr(model, "bert", model)
    for layer in bert.encoder.layer:
        if _is_erf_gelu(layer.intermediate.intermediate_act_fn):
            layer.intermediate.intermediate_act_fn = GELUFused().train(layer.training)
            if not isinstance(layer.intermediate, BertIntermediateFused):
                layer.intermediate = BertIntermediateFused(layer.intermediate).train(layer.training)
    return model


class BertEmbeddingsFused(torch.nn.Module):
    """HF ``BertEmbeddings`` (the embedding shell the reference keeps; quantized at quantized_bert.py:146-218) for eval with
    absolute positions: three gathers, two adds and the LayerNorm in ONE kernel (``oeh_embed_layer_norm``; the sums are rounded
    in the tensor dtype in HF's order).  Same submodule / buffer names, so state_dicts and the quantized wrapper keep working;
    anything else (training, ``inputs_embeds``, hooks on a submodule) runs the original forward."""

    def __init__(self, org):
        super().__init__()
        self.__dict__["_org"] = org                     # not a registered submodule: its children are re-registered below
        for name, mod in org.named_children():
            self.add_module(name, mod)
        for name, buf in org.named_buffers(recurse=False):
            self.register_buffer(name, buf, persistent=name in org.state_dict())
        self.position_embedding_type = getattr(org, "position_embedding_type", "absolute")

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None, past_key_values_length=0):
        mods = (self.word_embeddings, self.position_embeddings, self.token_type_embeddings, self.LayerNorm)
        w = self.word_embeddings.weight
        if (self.training or input_ids is None or inputs_embeds is not None or past_key_values_length != 0
                or self.position_embedding_type != "absolute" or not w.is_cuda or input_ids.dim() != 2
                or any(m._forward_hooks or m._forward_pre_hooks for m in mods)
                or (torch.is_grad_enabled() and w.requires_grad) or self.LayerNorm.weight.dtype != w.dtype):
            return self._org.forward(input_ids=input_ids, token_type_ids=token_type_ids, position_ids=position_ids,
                                     inputs_embeds=inputs_embeds, past_key_values_length=past_key_values_length)
        from . import ops
        ln = self.LayerNorm
        try:
            return ops.embed_layer_norm(input_ids, w, self.position_embeddings.weight, self.token_type_embeddings.weight,
                                        ln.weight, ln.bias, ln.eps, token_type_ids=token_type_ids, position_ids=position_ids)
        except NotImplementedError:                     # hidden size the register-resident kernel is not built for
            return self._org.forward(input_ids=input_ids, token_type_ids=token_type_ids, position_ids=position_ids)


class BertHeadTransformFused(torch.nn.Module):
    """HF ``BertPredictionHeadTransform`` (dense -> GELU -> LayerNorm, the FP32/FP16 MLM head the reference leaves
    unquantized, quantized_bert.py:866-870) with the library's kernels: GELU in the GEMM epilogue, LayerNorm in one pass."""

    def __init__(self, org):
        super().__init__()
        self.dense, self.transform_act_fn, self.LayerNorm = org.dense, org.transform_act_fn, org.LayerNorm

    def forward(self, hidden_states):
        d, ln = self.dense, self.LayerNorm
        if (self.training or not hidden_states.is_cuda or hidden_states.dtype not in (torch.float16, torch.bfloat16)
                or type(d) is not torch.nn.Linear or d.weight.dtype != hidden_states.dtype or not _is_erf_gelu(self.transform_act_fn)
                or (torch.is_grad_enabled() and (hidden_states.requires_grad or d.weight.requires_grad))
                or any(m._forward_hooks or m._forward_pre_hooks for m in (d, ln))
                or isinstance(self.transform_act_fn, torch.nn.Module) and (self.transform_act_fn._forward_hooks or self.transform_act_fn._forward_pre_hooks)
                or d.out_features % 8 or d.in_features % 8 or ln.bias is None):
            return ln(self.transform_act_fn(d(hidden_states)))
        from . import ops
        x = ops.linear_fwd(hidden_states, d.weight, d.bias, act="gelu")
        try:
            return ops.add_layer_norm(x, None, ln.weight, ln.bias, ln.eps)
        except NotImplementedError:
            return ln(x)


def fuse_bert_shell(model):
    """Swap ``bert.embeddings`` for :class:`BertEmbeddingsFused` and the MLM head's transform for :class:`BertHeadTransformFused`
    (eval only; other heads / embedding variants are left alone)."""
    bert = getattr(model, "bert", model)
    if type(bert.embeddings).__name__ == "BertEmbeddings":
        bert.embeddings = BertEmbeddingsFused(bert.embeddings).train(bert.embeddings.training)
    pred = getattr(getattr(model, "cls", None), "predictions", None)
    if pred is not None and type(pred.transform).__name__ == "BertPredictionHeadTransform":
        pred.transform = BertHeadTransformFused(pred.transform).train(pred.transform.training)
    return model


class OPTDecoderLayerFused(torch.nn.Module):
    """HF OPTDecoderLayer (pre-LayerNorm variants: 125m, 1.3b, ...; the module the reference wraps at
    quantized_opt.py:277-384) for eval with fewer, fused launches: LayerNorm kernels from liboeh, ``residual + attn`` and
    the following LayerNorm in ONE kernel (sum kept as the next residual), fc1 + bias + ReLU in one cuBLASLt call (the
    ReLU epilogue is exact).  Same submodule names, so state_dicts and the quantized wrapper keep working."""

    def __init__(self, org):
        super().__init__()
        self.embed_dim = org.embed_dim
        self.do_layer_norm_before = org.do_layer_norm_before
        self.dropout = org.dropout
        self.activation_fn = org.activation_fn
        self.self_attn, self.self_attn_layer_norm = org.self_attn, org.self_attn_layer_norm
        self.fc1, self.fc2, self.final_layer_norm = org.fc1, org.fc2, org.final_layer_norm
        self._org_forward = org.forward.__func__

    def _fusable(self, h) -> bool:
        ln1, ln2 = self.self_attn_layer_norm, self.final_layer_norm
        return (self.do_layer_norm_before and not self.training and h.is_cuda and isinstance(self.activation_fn, torch.nn.ReLU)
                and not (torch.is_grad_enabled() and h.requires_grad) and ln1.bias is not None and ln2.bias is not None
                and not any(m._forward_hooks or m._forward_pre_hooks for m in (ln1, ln2, self.fc1, self.fc2)))

    def forward(self, hidden_states, attention_mask=None, past_key_values=None, use_cache=False, position_ids=None, **kw):
        if not self._fusable(hidden_states):
            return self._org_forward(self, hidden_states, attention_mask=attention_mask, past_key_values=past_key_values,
                                     use_cache=use_cache, position_ids=position_ids, **kw)
        from . import ops
        ln1, ln2 = self.self_attn_layer_norm, self.final_layer_norm
        x = ops.add_layer_norm(hidden_states, None, ln1.weight, ln1.bias, ln1.eps)
        attn = self.self_attn(hidden_states=x, past_key_values=past_key_values, position_ids=position_ids,
                              attention_mask=attention_mask, **kw)[0]
        x2, residual = ops.add_layer_norm(hidden_states, attn, ln2.weight, ln2.bias, ln2.eps, return_sum=True)
        shape = x2.shape
        f = torch._addmm_activation(self.fc1.bias, x2.view(-1, shape[-1]), self.fc1.weight.t()) \
            if self.fc1.bias is not None else torch.relu(self.fc1(x2.view(-1, shape[-1])))
        f = self.fc2(f)
        return (residual.view(-1, shape[-1]) + f).view(shape)


    def gate_predictors(self, device):
        """Linear per-token gate predictors of this layer's attention (for the kernel that writes its input), or None."""
        fn = getattr(self.self_attn, "fusable_gate", None)
        return fn(device) if fn is not None and not self.self_attn.training else None

    def forward_chained(self, x, residual, next_ln, attention_mask=None, position_ids=None, gate_logits=None, next_gate=None):
        """The same layer for a caller that chains layers: ``x`` = this layer's LayerNorm-ed input (computed by the previous
        layer's last kernel), ``residual`` = the un-normalised stream.  Returns (LayerNorm_next(residual'), residual') with
        the final ``residual + ffn`` and the NEXT layer's first LayerNorm (``next_ln``) fused as well.  ``gate_logits``: this
        layer's per-token gate logits if the kernel that wrote ``x`` produced them; ``next_gate``: the next layer's gate
        predictors -- its logits are then emitted by this layer's last kernel and returned third."""
        from . import ops
        ln2 = self.final_layer_norm
        kw = {} if gate_logits is None else {"_gate_logits": gate_logits}
        attn = self.self_attn(hidden_states=x, past_key_values=None, position_ids=position_ids,
                              attention_mask=attention_mask, **kw)[0]
        x2, residual = ops.add_layer_norm(residual, attn, ln2.weight, ln2.bias, ln2.eps, return_sum=True)
        shape = x2.shape
        f = torch._addmm_activation(self.fc1.bias, x2.view(-1, shape[-1]), self.fc1.weight.t()) \
            if self.fc1.bias is not None else torch.relu(self.fc1(x2.view(-1, shape[-1])))
        f = self.fc2(f).view(shape)
        return ops.add_layer_norm(residual, f, next_ln.weight, next_ln.bias, next_ln.eps, return_sum=True, gate=next_gate)


def fuse_opt_decoder_layers(model):
    """Swap every decoder layer of an HF OPT model for the fused eval module."""
    dec = model.model.decoder
    for i, layer in enumerate(dec.layers):
        dec.layers[i] = OPTDecoderLayerFused(layer).train(layer.training)
    return model
