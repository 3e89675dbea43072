"""Model-level INT8 path (SURVEY 8a rows a13 + a15): QuantizedBertForMaskedLM + pass_data_for_range_estimation against
the LIVE reference's own QuantizedBertForMaskedLM (fixture tests/golden/quantized_bert_model.npz written by
tests/golden/make_golden_model.py): same FP weights, same batches, fp32 (the reference's INT8-eval precision).
Tolerances: the reference ran on CPU (MKL GEMMs), this on cuBLAS -- activations agree to ~1e-5 relative, so estimated
ranges must agree to 1e-3 relative (a min/max that moves by one fp32 GEMM rounding; at most 3 of the 30 may move by one
8-bit step of their input, 0.4 %, when a code flips at a bin edge upstream), the symmetric weight ranges bit-exactly, and the fixed-range loss to 2e-3 (codes flip at bin edges; the contract is perplexity within 0.1 %)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quantized_bert_model.npz")


def test_quant_config_and_params_match_reference_defaults():
    from outeffhop_b200 import quantized_models as QM
    from outeffhop_b200.quantization import (AsymmetricUniformQuantizer, CurrentMinMaxEstimator, RunningMinMaxEstimator,
                                             SymmetricUniformQuantizer)
    cfg = QM.get_quant_config()
    assert cfg.quant.n_bits == 8 and cfg.quant.n_bits_act == 8 and cfg.act_quant.num_batches == 16
    assert cfg.quant.percentile is None and cfg.nothing_here is None            # DotDict: missing key -> None
    qp = QM.val_qparams(cfg)
    assert qp["method"] is SymmetricUniformQuantizer and qp["act_method"] is AsymmetricUniformQuantizer
    assert qp["weight_range_method"] is CurrentMinMaxEstimator and qp["act_range_method"] is RunningMinMaxEstimator
    assert qp["quant_setup"] == "all" and qp["per_channel_weights"] is False and qp["weight_range_options"] == {}
    cfg.quant.percentile = 99.999
    assert QM.val_qparams(cfg)["percentile"] == 99.999


def test_pass_data_for_range_estimation_driver_semantics():
    from outeffhop_b200.quantized_models import pass_data_for_range_estimation

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.calls = []

        def forward(self, x=None, **kw):
            self.calls.append(x if x is not None else sorted(kw))

    m = M().train()
    loader = [{"input_ids": torch.zeros(1), "labels": torch.ones(1)} for _ in range(10)]
    pass_data_for_range_estimation(loader, m, act_quant=True, max_num_batches=4)
    assert len(m.calls) == 4 and m.calls[0] == ["input_ids", "labels"] and not m.training
    m.calls.clear()
    pass_data_for_range_estimation(loader, m, act_quant=False, max_num_batches=4)   # weights only: one batch (utils.py:68)
    assert len(m.calls) == 1
    m.calls.clear()
    got = pass_data_for_range_estimation([(torch.full((2,), 3.0), 0)] * 3, m, act_quant=True, max_num_batches=2)
    assert len(m.calls) == 2 and len(got) == 2 and float(got[0][0]) == 3.0


def _build(g, name):
    from transformers import BertConfig, BertForMaskedLM
    from outeffhop_b200 import hf_compat
    cfg = BertConfig(vocab_size=97, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                     max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg._attn_implementation = "eager"
    model = BertForMaskedLM(cfg).eval()
    gated = bool(int(g[f"{name}::gated"]))
    kw = dict(attn_gate_type="conditional_per_token", attn_gate_init=0.25) if gated else {}
    hf_compat.swap_bert_attention(model, attn_softmax=str(g[f"{name}::attn_softmax"]), max_seq_length=32, **kw)
    sd = {k.split("::w::")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}::w::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "position_ids" not in k], missing
    assert not unexpected, unexpected
    return model.cuda().eval()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sm1", "clipgate"])
def test_quantized_bert_model_vs_live_reference(name):
    from outeffhop_b200 import quantized_models as QM
    from outeffhop_b200.quantization import QuantizationManager
    g = np.load(GOLD)
    model = _build(g, name)
    ids, labels = torch.from_numpy(g[f"{name}::input_ids"]), torch.from_numpy(g[f"{name}::labels"])
    batches = [{"input_ids": ids[i], "labels": labels[i]} for i in range(4)]
    qmodel = QM.quantize_bert_for_masked_lm(model)
    QM.pass_data_for_range_estimation(batches[:3], qmodel, act_quant=True, max_num_batches=3)
    qmodel.fix_ranges()
    ours = {}
    for n, m in qmodel.named_modules():
        if isinstance(m, QuantizationManager) and m.quantizer.is_initialized:
            d = m.quantizer._delta
            if d is not None and d.numel() == 1:
                zf = getattr(m.quantizer, "_zero_float", None)
                ours[n] = (float(d), float(zf) if zf is not None else float("nan"))
    names = [str(n) for n in g[f"{name}::qnames"]]
    assert set(names) == set(ours), (sorted(set(names) - set(ours)), sorted(set(ours) - set(names)))
    loose = 0
    for n, d_ref, z_ref in zip(names, g[f"{name}::qdelta"], g[f"{name}::qzero_float"]):
        d, z = ours[n]
        if n.endswith("weight_quantizer"):
            assert np.float32(d) == np.float32(d_ref), (n, d, d_ref)              # weights are identical: bit-exact
        else:
            # one 8-bit step of the tensor feeding this tap (1/255 = 0.4 %) when a code flips at a bin edge upstream
            assert abs(d - d_ref) <= 6e-3 * abs(d_ref), (n, d, d_ref)
            assert abs(z - z_ref) <= 1.0, (n, z, z_ref)
            loose += abs(d - d_ref) > 1e-3 * abs(d_ref)
    assert loose <= 3, loose                                        # ... and that is the exception, not the rule
    with torch.no_grad():
        out = qmodel(input_ids=ids[3].cuda(), labels=labels[3].cuda())
    loss_ref = float(g[f"{name}::loss_fixed"])
    assert abs(float(out.loss) - loss_ref) <= 2e-3 * abs(loss_ref), (float(out.loss), loss_ref)
    diff = (out.logits.float().cpu().numpy() - g[f"{name}::logits_fixed"])
    assert np.abs(diff).mean() < 0.02 * np.abs(g[f"{name}::logits_fixed"]).mean()
    # quantization must be doing something, and must stay close to FP
    assert abs(loss_ref - float(g[f"{name}::loss_fp"])) > 1e-6


@pytest.mark.gpu
def test_int8_eval_step_graph_equals_eager_and_uses_fused_kernels():
    """fp16 model, fixed ranges: the captured INT8 step (fused attention with taps, one-kernel residual blocks, fused
    loss) reproduces the eager fixed-range forward."""
    from outeffhop_b200 import eval_harness as EH, ops, quantized_models as QM
    model = EH.build_bert_base(seed=4000, num_hidden_layers=2, attn_softmax="softmax1")
    qmodel = QM.quantize_bert_for_masked_lm(model)
    batches = [EH.synthetic_mlm_batch(8, 128, seed=4000 + i) for i in range(4)]
    calls = {"attn": 0, "ln": 0}
    o_attn, o_ln = ops.attention_fwd, ops.add_layer_norm
    step = EH.make_bert_int8_eval_step(qmodel, batches[0], batches[:3], use_graph=True)
    ops.attention_fwd = lambda *a, **k: (calls.__setitem__("attn", calls["attn"] + 1), o_attn(*a, **k))[1]
    ops.add_layer_norm = lambda *a, **k: (calls.__setitem__("ln", calls["ln"] + 1), o_ln(*a, **k))[1]
    try:
        with torch.no_grad():
            eager = float(qmodel(input_ids=batches[3]["input_ids"].cuda(), labels=batches[3]["labels"].cuda()).loss)
    finally:
        ops.attention_fwd, ops.add_layer_norm = o_attn, o_ln
    # per layer: 1 fused attention + 2 one-kernel residual blocks; + the embeddings' LayerNorm with its tap + the LayerNorm of
    # the FP MLM head's transform (hf_compat.BertHeadTransformFused)
    assert calls["attn"] == 2 and calls["ln"] == 6, calls
    step.load(batches[3])
    got = float(step.run())
    assert abs(got - eager) <= 1e-3 * abs(eager), (got, eager)
    assert np.isfinite(got)


# ------------------------------------------------------------------------------------------------ OPT (config 4)
GOLD_OPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quantized_opt_model.npz")


def _build_opt(g, name):
    from transformers import OPTConfig, OPTForCausalLM
    from outeffhop_b200 import hf_compat
    cfg = OPTConfig(vocab_size=101, hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2,
                    word_embed_proj_dim=128, max_position_embeddings=64, do_layer_norm_before=True, dropout=0.0,
                    attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    model = OPTForCausalLM(cfg).eval()
    gated = bool(int(g[f"{name}::gated"]))
    kw = dict(attn_gate_type="conditional_per_token", attn_gate_init=0.25) if gated else {}
    hf_compat.swap_opt_attention(model, attn_softmax=str(g[f"{name}::attn_softmax"]), max_seq_length=32, **kw)
    sd = {k.split("::w::")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}::w::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.cuda().eval()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sm1", "gate"])
def test_quantized_opt_model_vs_live_reference(name):
    """QuantizedOPTDecoder + FP32 lm_head (quant_setup fp32_head) against the live reference: same tolerances as BERT."""
    from outeffhop_b200 import quantized_models as QM
    from outeffhop_b200.quantization import QuantizationManager
    g = np.load(GOLD_OPT)
    model = _build_opt(g, name)
    ids = torch.from_numpy(g[f"{name}::input_ids"])
    batches = [{"input_ids": ids[i], "labels": ids[i].clone()} for i in range(4)]
    qmodel = QM.quantize_opt_for_causal_lm(model, quant_setup="fp32_head")
    QM.pass_data_for_range_estimation(batches[:3], qmodel, act_quant=True, max_num_batches=3)
    qmodel.fix_ranges()
    ours = {}
    for n, m in qmodel.named_modules():
        if isinstance(m, QuantizationManager) and m.quantizer.is_initialized:
            d = m.quantizer._delta
            if d is not None and d.numel() == 1:
                zf = getattr(m.quantizer, "_zero_float", None)
                ours[n[len("model.decoder."):]] = (float(d), float(zf) if zf is not None else float("nan"))
    names = [str(n) for n in g[f"{name}::qnames"]]
    assert set(names) == set(ours), (sorted(set(names) - set(ours)), sorted(set(ours) - set(names)))
    loose = 0
    for n, d_ref, z_ref in zip(names, g[f"{name}::qdelta"], g[f"{name}::qzero_float"]):
        d, z = ours[n]
        if n.endswith("weight_quantizer"):
            assert np.float32(d) == np.float32(d_ref), (n, d, d_ref)
        else:
            assert abs(d - d_ref) <= 6e-3 * abs(d_ref), (n, d, d_ref)
            assert abs(z - z_ref) <= 1.0, (n, z, z_ref)
            loose += abs(d - d_ref) > 1e-3 * abs(d_ref)
    assert loose <= 3, loose
    with torch.no_grad():
        out = qmodel(input_ids=ids[3].cuda(), labels=ids[3].cuda())
    loss_ref = float(g[f"{name}::loss_fixed"])
    assert abs(float(out.loss) - loss_ref) <= 2e-3 * abs(loss_ref), (float(out.loss), loss_ref)
    diff = out.logits.float().cpu().numpy() - g[f"{name}::logits_fixed"]
    assert np.abs(diff).mean() < 0.02 * np.abs(g[f"{name}::logits_fixed"]).mean()


@pytest.mark.gpu
def test_opt_int8_eval_step_graph_and_quant_setups():
    """BASELINE config 4 shape at 2 layers: OPT fp16, softmax1, fp32_head; captured fixed-range step == eager; the
    int16-LayerNorm setups switch the LayerNorm quantizers to 16 bit; unknown setups raise like the reference."""
    from outeffhop_b200 import eval_harness as EH, quantized_models as QM
    from outeffhop_b200.quantized_layers import QuantLayerNorm
    model = EH.build_opt_125m(seed=6789, num_hidden_layers=2, attn_softmax="softmax1")
    qmodel = QM.quantize_opt_for_causal_lm(model, quant_setup="fp32_head")
    batches = [EH.synthetic_clm_batch(2, 512, seed=6789 + i) for i in range(4)]
    step = EH.make_opt_int8_eval_step(qmodel, batches[0], batches[:3], use_graph=True)
    with torch.no_grad():
        eager = float(qmodel(input_ids=batches[3]["input_ids"].cuda(), labels=batches[3]["labels"].cuda()).loss)
    step.load(batches[3])
    got = float(step.run())
    assert np.isfinite(got) and abs(got - eager) <= 1e-3 * abs(eager), (got, eager)
    m2 = EH.build_opt_125m(seed=1, num_hidden_layers=1)
    q2 = QM.quantize_opt_for_causal_lm(m2, quant_setup="fp32_head+int16_all_LN")
    assert all(ln.activation_quantizer.quantizer.n_bits == 16 for ln in q2.modules() if isinstance(ln, QuantLayerNorm))
    with pytest.raises(ValueError):
        QM.quantize_opt_for_causal_lm(EH.build_opt_125m(seed=1, num_hidden_layers=1), quant_setup="nope")
