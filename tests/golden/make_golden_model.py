#!/usr/bin/env python
"""Golden fixture for the MODEL-LEVEL INT8 path (SURVEY 8a rows a13 + a15), generated from the LIVE reference
(build container only):

    python tests/golden/make_golden_model.py      -> tests/golden/quantized_bert_model.npz
    python tests/golden/make_golden_model.py opt  -> tests/golden/quantized_opt_model.npz   (QuantizedOPTDecoder, quantized_opt.py:387-662)

A small random-init BertForMaskedLM (2 layers, hidden 128 = 2 heads x 64, FFN 256, vocab 97, seq 32) whose attention is
the reference's BertSelfAttentionWithExtras (softmax1 / clipped softmax, optional per-token gate) is wrapped in the
reference's QuantizedBertForMaskedLM (transformers_language/models/quantized_bert.py:857-926) with the validate script's
parameters (validate_mlm_config.py:445-474: get_quant_config() + val_qparams(), W8A8, asymmetric activations with
running_minmax, symmetric weights with current_minmax, FP32 MLM head), ranges are estimated with
pass_data_for_range_estimation() (transformers_language/utils.py:50-71) on 3 batches, fixed, and one batch is evaluated.
Stored: the FP weights, the batches, every quantizer's (_delta, _zero_float) after estimation and the fixed-range loss /
logits.  HF-5.5 call-convention adapters as in SURVEY 8c.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402


def build(attn_softmax, gated, seed):
    from transformers import BertConfig, BertForMaskedLM
    from transformers_language.models.bert_attention import AttentionGateType, BertSelfAttentionWithExtras
    from transformers_language.models.softmax import SOFTMAX_MAPPING

    cfg = BertConfig(vocab_size=97, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                     max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = BertForMaskedLM(cfg).eval()
    with torch.no_grad():                                   # wider activations than the 0.02-std init gives
        for n, p in model.named_parameters():
            if "LayerNorm.weight" in n:
                p.copy_(1.0 + 0.2 * torch.randn_like(p))
            elif p.dim() == 2 and "embeddings" not in n:
                p.mul_(4.0)
    for layer in model.bert.encoder.layer:
        old = layer.attention.self
        new = BertSelfAttentionWithExtras(
            cfg, softmax_fn=SOFTMAX_MAPPING[attn_softmax], max_seq_length=32,
            attn_gate_type=AttentionGateType.conditional_per_token if gated else AttentionGateType.none,
            attn_gate_init=0.25 if gated else None)
        new.load_state_dict(old.state_dict(), strict=False)
        layer.attention.self = new.eval()
    return cfg, model


class _SelfAdapter(torch.nn.Module):
    """HF >= 4.48 BertAttention calls self.self(hidden_states, attention_mask=..., **kw) and unpacks two values."""

    def __init__(self, ref):
        super().__init__()
        self.ref = ref

    def forward(self, hidden_states, attention_mask=None, **kw):
        return self.ref(hidden_states, attention_mask)[0], None


def gen_case(out, name, attn_softmax, gated, seed):
    from quantization.quantizers import QMethods  # noqa: F401
    from transformers_language.models.quantized_bert import QuantizedBertForMaskedLM
    from transformers_language.quant_configs import get_quant_config
    from transformers_language.utils import pass_data_for_range_estimation, val_qparams

    cfg, model = build(attn_softmax, gated, seed)
    for k, v in model.state_dict().items():
        out[f"{name}::w::{k}"] = v.numpy().copy()
    g = torch.Generator().manual_seed(seed + 1)
    B, T = 4, 32
    batches = []
    for _ in range(4):
        ids = torch.randint(0, 97, (B, T), generator=g)
        masked = torch.rand(B, T, generator=g) < 0.3
        masked[:, 0] = True
        labels = torch.where(masked, ids, torch.full_like(ids, -100))
        batches.append({"input_ids": ids, "labels": labels})
    out[f"{name}::input_ids"] = np.stack([b["input_ids"].numpy() for b in batches])
    out[f"{name}::labels"] = np.stack([b["labels"].numpy() for b in batches])

    raw = [layer.attention.self for layer in model.bert.encoder.layer]
    for layer in model.bert.encoder.layer:                  # HF-5.5 call convention, FP pass only
        layer.attention.self = _SelfAdapter(layer.attention.self)
    with torch.no_grad():
        fp_loss = float(model(**batches[3]).loss)
    for layer, r in zip(model.bert.encoder.layer, raw):     # quantize_model dispatches on the exact reference type
        layer.attention.self = r
    qcfg = get_quant_config()
    qparams = val_qparams(qcfg)
    qparams["quant_dict"] = {}
    qmodel = QuantizedBertForMaskedLM(model, **qparams).eval()
    # the reference's QuantizedBertModel.forward predates HF 5.x: drive the quantized submodules in its order
    # (quantized_bert.py:725-818: embeddings -> encoder layers -> [no pooler]; head in FP)
    def forward(input_ids, labels=None):
        h = qmodel.bert.embeddings(input_ids=input_ids)
        for layer in qmodel.bert.encoder.layer:            # QuantizedBertLayer.forward (:648-682) with HF-4.31's
            att = layer.attention                           # BertAttention.forward: self -> output(self_out, hidden)
            self_out = att.self(h, None, None)[0]
            att_out = att.output(self_out, h)
            h = layer.output(layer.intermediate(att_out), att_out)
        logits = qmodel.cls(h)
        loss = None
        if labels is not None:
            loss = torch.nn.functional.cross_entropy(logits.view(-1, cfg.vocab_size), labels.view(-1))
        return loss, logits

    class _Wrap(torch.nn.Module):                           # what pass_data_for_range_estimation calls: model(**batch)
        def __init__(self):
            super().__init__()
            self.q = qmodel

        def forward(self, input_ids, labels=None):
            return forward(input_ids, labels)

    wrap = _Wrap().eval()
    qmodel.set_quant_state(weight_quant=True, act_quant=True)          # validate_mlm_config.py:466
    pass_data_for_range_estimation(loader=batches[:3], model=wrap, act_quant=True, max_num_batches=3)   # :469-474
    qmodel.fix_ranges()                                                # :476
    names, deltas, zfs = [], [], []
    for n, m in qmodel.named_modules():
        q = getattr(m, "quantizer", None)
        if q is not None and hasattr(q, "_delta") and q._delta is not None and m.__class__.__name__ == "QuantizationManager":
            d = q._delta.detach().float().reshape(-1)
            if d.numel() == 1:
                names.append(n)
                deltas.append(float(d))
                zf = getattr(q, "_zero_float", None)
                zfs.append(float(zf) if zf is not None else float("nan"))
    out[f"{name}::qnames"] = np.array(names)
    out[f"{name}::qdelta"] = np.array(deltas, dtype=np.float32)
    out[f"{name}::qzero_float"] = np.array(zfs, dtype=np.float32)
    with torch.no_grad():
        loss, logits = forward(**batches[3])
    out[f"{name}::loss_fixed"] = np.array(float(loss), dtype=np.float64)
    out[f"{name}::logits_fixed"] = logits.numpy()
    out[f"{name}::loss_fp"] = np.array(fp_loss, dtype=np.float64)
    out[f"{name}::attn_softmax"] = np.array(attn_softmax)
    out[f"{name}::gated"] = np.array(int(gated))
    print(name, "fp loss", fp_loss, "int8 loss", float(loss), "quantizers", len(names))


def main():
    G._install_shims()
    out = {}
    gen_case(out, "sm1", "softmax1", False, 4000)
    gen_case(out, "clipgate", "clipped(-.025:1)", True, 4001)
    np.savez_compressed(os.path.join(G.OUT, "quantized_bert_model.npz"), **out)
    print("wrote", os.path.join(G.OUT, "quantized_bert_model.npz"))




# ------------------------------------------------------------------------------------------------ OPT (causal LM)
def build_opt(attn_softmax, gated, seed):
    import transformers.models.opt.modeling_opt as mo
    for name in ("_expand_mask", "_make_causal_mask"):      # quantized_opt.py:8-14: import-time names only (the mask is
        if not hasattr(mo, name):                           # built below with HF-4.31 semantics)
            setattr(mo, name, lambda *a, **k: None)
    from transformers import OPTConfig, OPTForCausalLM
    from transformers_language.models.bert_attention import AttentionGateType
    from transformers_language.models.opt_attention import OPTAttentionWithExtras
    from transformers_language.models.softmax import SOFTMAX_MAPPING

    cfg = OPTConfig(vocab_size=101, hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2,
                    word_embed_proj_dim=128, max_position_embeddings=64, do_layer_norm_before=True, dropout=0.0,
                    attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = OPTForCausalLM(cfg).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "layer_norm.weight" in n:
                p.copy_(1.0 + 0.2 * torch.randn_like(p))
            elif p.dim() == 2 and "embed" not in n:
                p.mul_(4.0)
    for layer in model.model.decoder.layers:
        old = layer.self_attn
        new = OPTAttentionWithExtras(
            embed_dim=cfg.hidden_size, num_heads=cfg.num_attention_heads, dropout=0.0, is_decoder=True, bias=True,
            softmax_fn=SOFTMAX_MAPPING[attn_softmax], max_seq_length=32,
            attn_gate_type=AttentionGateType.conditional_per_token if gated else AttentionGateType.none,
            attn_gate_init=0.25 if gated else None)
        new.load_state_dict(old.state_dict(), strict=False)
        layer.self_attn = new.eval()
    return cfg, model


def causal_mask_431(B, T, dtype):
    """HF-4.31 ``_make_causal_mask`` + all-ones ``_expand_mask``: (B,1,T,T), 0 on/below the diagonal, finfo.min above."""
    m = torch.full((T, T), torch.finfo(dtype).min, dtype=dtype).triu(1)
    return m[None, None].expand(B, 1, T, T).contiguous()


def gen_opt_case(out, name, attn_softmax, gated, seed):
    import transformers.models.opt.modeling_opt as mo
    for stub in ("_expand_mask", "_make_causal_mask"):
        if not hasattr(mo, stub):
            setattr(mo, stub, lambda *a, **k: None)
    from transformers_language.models.quantized_opt import QuantizedOPTDecoder
    from transformers_language.quant_configs import get_quant_config
    from transformers_language.utils import pass_data_for_range_estimation, val_qparams

    cfg, model = build_opt(attn_softmax, gated, seed)
    for k, v in model.state_dict().items():
        out[f"{name}::w::{k}"] = v.numpy().copy()
    g = torch.Generator().manual_seed(seed + 1)
    B, T = 4, 32
    batches = []
    for _ in range(4):
        ids = torch.randint(0, 101, (B, T), generator=g)
        batches.append({"input_ids": ids, "labels": ids.clone()})
    out[f"{name}::input_ids"] = np.stack([b["input_ids"].numpy() for b in batches])
    qparams = val_qparams(get_quant_config())
    qdec = QuantizedOPTDecoder(model.model.decoder, **qparams).eval()      # what QuantizedOPTForCausalLM wraps (fp32_head)
    mask = causal_mask_431(B, T, torch.float32)
    ones = torch.ones(B, T, dtype=torch.bool)

    def forward(input_ids, labels=None):
        # QuantizedOPTDecoder.forward (quantized_opt.py:459-662) for un-padded blocks, then the FP32 lm_head (:731-732)
        h = qdec.embed_tokens(input_ids) + qdec.embed_positions(ones, 0)
        h = qdec.embed_sum_act_quantizer(h)
        for layer in qdec.layers:
            h = layer(h, attention_mask=mask)[0]
        h = qdec.final_layer_norm(h)
        logits = model.lm_head(h)
        loss = None
        if labels is not None:
            loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, cfg.vocab_size), labels[:, 1:].reshape(-1))
        return loss, logits

    class _Wrap(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q = qdec

        def forward(self, input_ids, labels=None):
            return forward(input_ids, labels)

    qdec.set_quant_state(weight_quant=True, act_quant=True)
    pass_data_for_range_estimation(loader=batches[:3], model=_Wrap().eval(), act_quant=True, max_num_batches=3)
    qdec.fix_ranges()
    names, deltas, zfs = [], [], []
    for n, m in qdec.named_modules():
        q = getattr(m, "quantizer", None)
        if q is not None and m.__class__.__name__ == "QuantizationManager" and getattr(q, "_delta", None) is not None:
            d = q._delta.detach().float().reshape(-1)
            if d.numel() == 1:
                names.append(n)
                deltas.append(float(d))
                zf = getattr(q, "_zero_float", None)
                zfs.append(float(zf) if zf is not None else float("nan"))
    out[f"{name}::qnames"] = np.array(names)
    out[f"{name}::qdelta"] = np.array(deltas, dtype=np.float32)
    out[f"{name}::qzero_float"] = np.array(zfs, dtype=np.float32)
    with torch.no_grad():
        loss, logits = forward(**batches[3])
    out[f"{name}::loss_fixed"] = np.array(float(loss), dtype=np.float64)
    out[f"{name}::logits_fixed"] = logits.numpy()
    out[f"{name}::attn_softmax"] = np.array(attn_softmax)
    out[f"{name}::gated"] = np.array(int(gated))
    print(name, "int8 loss", float(loss), "quantizers", len(names))


def main_opt():
    G._install_shims()
    out = {}
    gen_opt_case(out, "sm1", "softmax1", False, 6789)
    gen_opt_case(out, "gate", "vanilla", True, 6790)
    np.savez_compressed(os.path.join(G.OUT, "quantized_opt_model.npz"), **out)
    print("wrote", os.path.join(G.OUT, "quantized_opt_model.npz"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "opt":
        main_opt()
    else:
        main()
