#!/usr/bin/env python
"""Golden fixtures for the FULL-SIZE INT8 configs, generated from the LIVE reference (build container only, CPU, fp32):

    python tests/golden/make_golden_fullsize.py bert   -> tests/golden/fullsize_bert_int8.npz
    python tests/golden/make_golden_fullsize.py opt    -> tests/golden/fullsize_opt_int8.npz      (BASELINE config 4)

bert: BERT-base (BertConfig() defaults, random init from torch.manual_seed(4000)), 32 x 128 tokens, the validate script's
      preparation (validate_mlm_config.py:166-191: GELUActivation -> nn.GELU(), attention -> BertSelfAttentionWithExtras
      with --attn_softmax "clippedsoftmax1(-.025:1)"), --quantize with the default W8A8 config (:445-476), ranges estimated
      on 4 batches, fixed, 2 batches evaluated.
opt:  OPT-125m shape (random init from torch.manual_seed(6789)), 4 x 512 tokens, softmax1, quant_setup fp32_head
      (validate_clm.py:418-499), 4 estimation batches, 2 evaluated.

Each model is run TWICE: as built, and with every Linear bias of the encoder / decoder layers perturbed by N(0, 1e-6)
(pre-activations have std ~0.55 here, so this is the size of the last-bit differences between two correct fp32 GEMM
implementations, e.g. MKL on the CPU vs cuBLAS).  The spread between the two REFERENCE runs (``*_pert`` entries) is the
reference's own sensitivity to such differences -- the noise floor any other implementation is compared against.

Weights are NOT stored: both sides build the HF model from the same seed with the same torch (the fixture carries a
float64 checksum of the state_dict so a drift in the initialiser shows up as a checksum failure, not as a parity failure).
Batches come from outeffhop_b200.eval_harness.synthetic_{mlm,clm}_batch (seeded, CPU).  Stored: every activation / weight
quantizer's (_delta, _zero_float) after estimation, the fixed-range losses, the FP losses and a slice of the logits.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as G  # noqa: E402
import make_golden_model as GM  # noqa: E402

N_EST, N_EVAL = 4, 2


def state_checksum(model) -> np.ndarray:
    """[sum, sum of |.|] over every tensor of the state_dict, float64."""
    s = a = 0.0
    for v in model.state_dict().values():
        if v.is_floating_point():
            s += float(v.double().sum())
            a += float(v.double().abs().sum())
    return np.array([s, a], dtype=np.float64)


def perturb_linear_biases(root, seed, sigma=1e-6):
    """b += N(0, sigma) on every nn.Linear under ``root``: a stand-in for the last-bit differences between two correct
    fp32 GEMM implementations (biases are not quantized, hijacker.py:84-101)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in root.modules():
            if isinstance(m, torch.nn.Linear) and m.bias is not None:
                m.bias.add_(torch.randn(m.bias.shape, generator=g) * sigma)


def collect_ranges(qmodel, out, strip=""):
    names, deltas, zfs = [], [], []
    for n, m in qmodel.named_modules():
        q = getattr(m, "quantizer", None)
        if q is not None and m.__class__.__name__ == "QuantizationManager" and getattr(q, "_delta", None) is not None:
            d = q._delta.detach().float().reshape(-1)
            if d.numel() == 1:
                names.append(n[len(strip):] if strip and n.startswith(strip) else n)
                deltas.append(float(d))
                zf = getattr(q, "_zero_float", None)
                zfs.append(float(zf) if zf is not None else float("nan"))
    out["qnames"] = np.array(names)
    out["qdelta"] = np.array(deltas, dtype=np.float32)
    out["qzero_float"] = np.array(zfs, dtype=np.float32)
    return len(names)


def gen_bert():
    from torch import nn
    from transformers import BertConfig, BertForMaskedLM
    from outeffhop_b200.eval_harness import synthetic_mlm_batch
    from transformers_language.models.bert_attention import AttentionGateType, BertSelfAttentionWithExtras
    from transformers_language.models.quantized_bert import QuantizedBertForMaskedLM
    from transformers_language.models.softmax import SOFTMAX_MAPPING
    from transformers_language.quant_configs import get_quant_config
    from transformers_language.utils import pass_data_for_range_estimation, val_qparams

    seed, attn_softmax, B, T = 4000, "clippedsoftmax1(-.025:1)", 32, 128
    cfg = BertConfig()
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = BertForMaskedLM(cfg).eval()
    out = {"w_checksum": state_checksum(model), "attn_softmax": np.array(attn_softmax), "seed": np.array(seed)}
    for layer in model.bert.encoder.layer:                                    # validate_mlm_config.py:166-168
        layer.intermediate.intermediate_act_fn = nn.GELU()
    for layer in model.bert.encoder.layer:                                    # :174-191
        old = layer.attention.self
        new = BertSelfAttentionWithExtras(cfg, softmax_fn=SOFTMAX_MAPPING[attn_softmax], max_seq_length=T,
                                          attn_gate_type=AttentionGateType.none)
        new.load_state_dict(old.state_dict(), strict=False)
        layer.attention.self = new.eval()
    batches = [synthetic_mlm_batch(B, T, seed=seed + i, pin=False) for i in range(N_EST + N_EVAL)]

    raw = [layer.attention.self for layer in model.bert.encoder.layer]
    for layer in model.bert.encoder.layer:
        layer.attention.self = GM._SelfAdapter(layer.attention.self)
    with torch.no_grad():
        out["loss_fp"] = np.array([float(model(**b).loss) for b in batches[N_EST:]], dtype=np.float64)
    for layer, r in zip(model.bert.encoder.layer, raw):
        layer.attention.self = r

    def run(tag):
        qparams = val_qparams(get_quant_config())
        qparams["quant_dict"] = {}
        qmodel = QuantizedBertForMaskedLM(model, **qparams).eval()

        def forward(input_ids, labels=None):                                  # quantized_bert.py:725-818, :648-682
            h = qmodel.bert.embeddings(input_ids=input_ids)
            for layer in qmodel.bert.encoder.layer:
                att = layer.attention
                self_out = att.self(h, None, None)[0]
                att_out = att.output(self_out, h)
                h = layer.output(layer.intermediate(att_out), att_out)
            logits = qmodel.cls(h)
            loss = None
            if labels is not None:
                loss = torch.nn.functional.cross_entropy(logits.view(-1, cfg.vocab_size), labels.view(-1))
            return loss, logits

        class _Wrap(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.q = qmodel

            def forward(self, input_ids, labels=None):
                return forward(input_ids, labels)

        qmodel.set_quant_state(weight_quant=True, act_quant=True)                               # :466
        pass_data_for_range_estimation(loader=batches[:N_EST], model=_Wrap().eval(), act_quant=True, max_num_batches=N_EST)
        qmodel.fix_ranges()                                                                     # :476
        res = {}
        n = collect_ranges(qmodel, res)
        losses, slices = [], []
        with torch.no_grad():
            for b in batches[N_EST:]:
                loss, logits = forward(**b)
                losses.append(float(loss))
                slices.append(logits[:2, :16, :256].numpy().copy())
        res["loss_fixed"] = np.array(losses, dtype=np.float64)
        res["logits_slice"] = np.stack(slices)
        for k2, v2 in res.items():
            out[k2 + tag] = v2
        return n

    n = run("")
    perturb_linear_biases(model.bert.encoder, seed + 99)
    run("_pert")
    print("bert: fp", out["loss_fp"], "int8", out["loss_fixed"], "quantizers", n)
    np.savez_compressed(os.path.join(G.OUT, "fullsize_bert_int8.npz"), **out)


def gen_opt():
    import transformers.models.opt.modeling_opt as mo
    for name in ("_expand_mask", "_make_causal_mask"):
        if not hasattr(mo, name):
            setattr(mo, name, lambda *a, **k: None)
    from transformers import OPTConfig, OPTForCausalLM
    from outeffhop_b200.eval_harness import synthetic_clm_batch
    from transformers_language.models.bert_attention import AttentionGateType
    from transformers_language.models.opt_attention import OPTAttentionWithExtras
    from transformers_language.models.quantized_opt import QuantizedOPTDecoder
    from transformers_language.models.softmax import SOFTMAX_MAPPING
    from transformers_language.quant_configs import get_quant_config
    from transformers_language.utils import pass_data_for_range_estimation, val_qparams

    seed, attn_softmax, B, T = 6789, "softmax1", 4, 512
    cfg = OPTConfig(hidden_size=768, num_hidden_layers=12, ffn_dim=3072, num_attention_heads=12, word_embed_proj_dim=768,
                    max_position_embeddings=2048, do_layer_norm_before=True, vocab_size=50272)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = OPTForCausalLM(cfg).eval()
    out = {"w_checksum": state_checksum(model), "attn_softmax": np.array(attn_softmax), "seed": np.array(seed)}
    for layer in model.model.decoder.layers:                                  # validate_clm.py:147-168
        old = layer.self_attn
        new = OPTAttentionWithExtras(embed_dim=cfg.hidden_size, num_heads=cfg.num_attention_heads, dropout=cfg.attention_dropout,
                                     is_decoder=True, bias=True, softmax_fn=SOFTMAX_MAPPING[attn_softmax], max_seq_length=T,
                                     attn_gate_type=AttentionGateType.none)
        new.load_state_dict(old.state_dict(), strict=False)
        layer.self_attn = new.eval()
    batches = [synthetic_clm_batch(B, T, seed=seed + i, pin=False) for i in range(N_EST + N_EVAL)]
    mask = GM.causal_mask_431(B, T, torch.float32)
    ones = torch.ones(B, T, dtype=torch.bool)

    def run(tag):
        qparams = val_qparams(get_quant_config())
        qparams["quant_dict"] = {}
        qdec = QuantizedOPTDecoder(model.model.decoder, **qparams).eval()     # quant_setup fp32_head: FP lm_head (:731-732)

        def forward(input_ids, labels=None):                                  # quantized_opt.py:459-662
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

        if not tag:                                                           # FP losses (quantizers off, weights FP)
            qdec.set_quant_state(weight_quant=False, act_quant=False)
            with torch.no_grad():
                out["loss_fp"] = np.array([float(forward(**b)[0]) for b in batches[N_EST:]], dtype=np.float64)
        qdec.set_quant_state(weight_quant=True, act_quant=True)                                 # validate_clm.py:478-480
        pass_data_for_range_estimation(loader=batches[:N_EST], model=_Wrap().eval(), act_quant=True, max_num_batches=N_EST)
        qdec.fix_ranges()                                                                       # :495
        res = {}
        n = collect_ranges(qdec, res)
        losses, slices = [], []
        with torch.no_grad():
            for b in batches[N_EST:]:
                loss, logits = forward(**b)
                losses.append(float(loss))
                slices.append(logits[:2, :16, :256].numpy().copy())
        res["loss_fixed"] = np.array(losses, dtype=np.float64)
        res["logits_slice"] = np.stack(slices)
        for k2, v2 in res.items():
            out[k2 + tag] = v2
        return n

    n = run("")
    perturb_linear_biases(model.model.decoder.layers, seed + 99)
    run("_pert")
    print("opt: fp", out["loss_fp"], "int8", out["loss_fixed"], "quantizers", n)
    np.savez_compressed(os.path.join(G.OUT, "fullsize_opt_int8.npz"), **out)


if __name__ == "__main__":
    G._install_shims()
    torch.set_num_threads(os.cpu_count() or 8)
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    if which in ("bert", "both"):
        gen_bert()
    if which in ("opt", "both"):
        gen_opt()
