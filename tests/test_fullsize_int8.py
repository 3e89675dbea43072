"""FULL-SIZE INT8 parity against the live reference (fixtures tests/golden/fullsize_{bert,opt}_int8.npz written by
tests/golden/make_golden_fullsize.py from /root/reference on CPU, fp32):

* BERT-base 32 x 128, clippedsoftmax1(-.025:1), W8A8, 4 estimation batches (validate_mlm_config.py:166-191, :445-476)
* OPT-125m 4 x 512, softmax1, quant_setup fp32_head -- BASELINE config 4 (validate_clm.py:147-168, :418-499)

Both sides build the HF model from the same seed (checksum-checked), run the same seeded batches.

What can be pinned, and what cannot (MEASURED, see the ``*_pert`` entries of the fixtures): the W8A8 network is chaotic in
the last bits of its inputs.  The generator runs the REFERENCE twice, the second time with N(0, 1e-6) added to every
Linear bias (the size of the difference between two correct fp32 GEMM implementations).  Reference vs reference then
differs by: activation ranges 5e-7 (median) in layer 0 growing to ~1e-2 in layer 11 (max 0.11 BERT / 0.026 OPT: a code
that flips at a bin edge moves a tensor by 1/255 of its range, the next tap's max moves with it, and clippedsoftmax1
turns a 1 % change of the largest probability into a 4 % change of its clipped value); logits 6 % of their mean
magnitude; loss 2.5e-3 / 1.1e-3 (BERT) and 8e-4 / 6e-4 (OPT) -- i.e. the reference reproduces ITS OWN perplexity only to
0.25 % on this random-init model.  north_star's "perplexity within 0.1 %" is therefore not a property any second
implementation can have here; the contract tested is:
  * FP (unquantized) loss: |d| <= 2e-5 (fp32) / 2e-4 (fp16)                  -- no chaos without the quantizers
  * weight ranges bit-exact (fp32)
  * layer-0 / embedding taps (nothing upstream to amplify): median <= 1e-5, max <= 3e-3 relative (fp32)
  * everything downstream no further from the reference than 3x the reference's distance from itself
    (range median / p90, logits), 4x for the single worst range, and |d loss| <= 1.5x the reference's own + 5e-4.
"""
import json
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
N_EST, N_EVAL = 4, 2


def _checksum(model):
    s = a = 0.0
    for v in model.state_dict().values():
        if v.is_floating_point():
            s += float(v.double().sum())
            a += float(v.double().abs().sum())
    return np.array([s, a])


def _record(name, payload):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "fullsize_int8_parity.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **payload}, default=float) + "\n")
    except OSError:
        pass


def _ranges(qmodel, strip=""):
    from outeffhop_b200.quantization import QuantizationManager
    ours = {}
    for n, m in qmodel.named_modules():
        if isinstance(m, QuantizationManager) and m.quantizer.is_initialized:
            d = m.quantizer._delta
            if d is not None and d.numel() == 1:
                zf = getattr(m.quantizer, "_zero_float", None)
                ours[n[len(strip):] if strip and n.startswith(strip) else n] = (
                    float(d), float(zf) if zf is not None else float("nan"))
    return ours


def _compare_ranges(g, ours):
    """Per-quantizer relative difference of ``_delta`` (activation taps) / bit-equality (weights)."""
    names = [str(n) for n in g["qnames"]]
    assert set(names) == set(ours), (sorted(set(names) - set(ours))[:5], sorted(set(ours) - set(names))[:5])
    rels, weight_mismatch, zdiff = {}, [], 0.0
    for n, d_ref, z_ref in zip(names, g["qdelta"], g["qzero_float"]):
        d, z = ours[n]
        if n.endswith("weight_quantizer"):
            if np.float32(d) != np.float32(d_ref):
                weight_mismatch.append((n, d, float(d_ref)))
            continue
        rels[n] = abs(d - d_ref) / abs(d_ref)
        zdiff = max(zdiff, abs(z - z_ref))
    return rels, weight_mismatch, zdiff


def _range_stats(rels):
    v = np.array(sorted(rels.values()))
    worst = sorted(rels.items(), key=lambda kv: -kv[1])[:8]
    return dict(n_act=len(v), median=float(np.median(v)), p90=float(np.quantile(v, 0.9)), max=float(v.max()),
                n_gt_1e3=int((v > 1e-3).sum()), n_gt_6e3=int((v > 6e-3).sum()), worst=[(k, float(x)) for k, x in worst])


def _self_sensitivity(g):
    """Reference vs perturbed reference (see the module docstring)."""
    names = [str(n) for n in g["qnames"]]
    act = [i for i, n in enumerate(names) if not n.endswith("weight_quantizer")]
    rel = np.abs(g["qdelta_pert"][act] - g["qdelta"][act]) / np.abs(g["qdelta"][act])
    return dict(median=float(np.median(rel)), p90=float(np.quantile(rel, 0.9)), max=float(rel.max()),
                dloss=float(np.abs(g["loss_fixed_pert"] - g["loss_fixed"]).max()),
                logits=float(np.abs(g["logits_slice_pert"] - g["logits_slice"]).mean()))


def _first_layer(name: str) -> bool:
    return ("embed" in name or ".layer.0." in name or name.startswith("layers.0.")) and "final_layer_norm" not in name


def _check(g, rels, wbad, losses, logits_diff, dtype):
    ref = _self_sensitivity(g)
    st = _range_stats(rels)
    if dtype == torch.float32:
        assert not wbad, wbad[:3]                                  # identical weights: bit-exact
        first = np.array([v for n, v in rels.items() if _first_layer(n)])
        assert len(first) >= 8 and np.median(first) <= 1e-5 and first.max() <= 3e-3, (len(first), np.median(first), first.max())
    assert st["median"] <= 3 * ref["median"] and st["p90"] <= 3 * ref["p90"] and st["max"] <= 4 * ref["max"], (st, ref)
    dl = np.abs(np.array(losses) - g["loss_fixed"]).max()
    assert dl <= 1.5 * ref["dloss"] + 5e-4, (dl, ref["dloss"])
    assert logits_diff <= 1.5 * ref["logits"], (logits_diff, ref["logits"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16], ids=["fp32", "fp16"])
def test_bert_base_int8_fullsize_vs_live_reference(dtype):
    from outeffhop_b200 import eval_harness as EH, quantized_models as QM
    g = np.load(os.path.join(HERE, "golden", "fullsize_bert_int8.npz"))
    seed, B, T = int(g["seed"]), 32, 128
    model = EH.build_bert_base(seed=seed, dtype=torch.float32, attn_softmax=str(g["attn_softmax"]))
    np.testing.assert_allclose(_checksum(model), g["w_checksum"], rtol=1e-9)    # same random init as the reference run
    model = model.to(dtype)
    batches = [EH.synthetic_mlm_batch(B, T, seed=seed + i) for i in range(N_EST + N_EVAL)]
    with torch.no_grad():
        fp = [float(EH.bert_mlm_loss(model, b["input_ids"].cuda(), b["labels"].cuda())) for b in batches[N_EST:]]
    qmodel = QM.quantize_bert_for_masked_lm(model)
    QM.pass_data_for_range_estimation(batches[:N_EST], qmodel, act_quant=True, max_num_batches=N_EST)
    qmodel.fix_ranges()
    rels, wbad, zdiff = _compare_ranges(g, _ranges(qmodel))
    losses = []
    with torch.no_grad():
        for b in batches[N_EST:]:
            out = qmodel(input_ids=b["input_ids"].cuda(), labels=b["labels"].cuda())
            losses.append(float(out.loss))
    sl = out.logits.reshape(B, T, -1)[:2, :16, :256].float().cpu().numpy()
    ref_sl = g["logits_slice"][-1]
    _record(f"bert_{dtype}", dict(loss=losses, loss_ref=g["loss_fixed"].tolist(), loss_fp=fp, loss_fp_ref=g["loss_fp"].tolist(),
                                  ref_self=_self_sensitivity(g), ranges=_range_stats(rels), weight_mismatch=wbad[:3], zdiff=zdiff,
                                  logits_mean_abs_diff=float(np.abs(sl - ref_sl).mean()), logits_mean_abs=float(np.abs(ref_sl).mean())))
    np.testing.assert_allclose(fp, g["loss_fp"], atol=2e-4 if dtype == torch.float16 else 2e-5, rtol=0)
    _check(g, rels, wbad, losses, float(np.abs(sl - ref_sl).mean()), dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16], ids=["fp32", "fp16"])
def test_opt_125m_int8_config4_fullsize_vs_live_reference(dtype):
    from outeffhop_b200 import eval_harness as EH, quantized_models as QM
    g = np.load(os.path.join(HERE, "golden", "fullsize_opt_int8.npz"))
    seed, B, T = int(g["seed"]), 4, 512
    model = EH.build_opt_125m(seed=seed, dtype=torch.float32, attn_softmax=str(g["attn_softmax"]))
    np.testing.assert_allclose(_checksum(model), g["w_checksum"], rtol=1e-9)
    model = model.to(dtype)
    batches = [EH.synthetic_clm_batch(B, T, seed=seed + i) for i in range(N_EST + N_EVAL)]
    mask = EH.make_causal_mask(B, T, dtype, "cuda")
    pos = torch.arange(T, device="cuda").unsqueeze(0).expand(B, T).contiguous()
    with torch.no_grad():
        fp = [float(EH.opt_clm_loss(model, b["input_ids"].cuda(), b["labels"].cuda(), mask, pos)) for b in batches[N_EST:]]
    qmodel = QM.quantize_opt_for_causal_lm(model, quant_setup="fp32_head")
    QM.pass_data_for_range_estimation(batches[:N_EST], qmodel, act_quant=True, max_num_batches=N_EST)
    qmodel.fix_ranges()
    rels, wbad, zdiff = _compare_ranges(g, _ranges(qmodel, strip="model.decoder."))
    losses = []
    with torch.no_grad():
        for b in batches[N_EST:]:
            out = qmodel(input_ids=b["input_ids"].cuda(), labels=b["labels"].cuda())
            losses.append(float(out.loss))
    sl = out.logits.reshape(B, T, -1)[:2, :16, :256].float().cpu().numpy()
    ref_sl = g["logits_slice"][-1]
    _record(f"opt_{dtype}", dict(loss=losses, loss_ref=g["loss_fixed"].tolist(), loss_fp=fp, loss_fp_ref=g["loss_fp"].tolist(),
                                 ref_self=_self_sensitivity(g), ranges=_range_stats(rels),
                                 weight_mismatch=wbad[:3], zdiff=zdiff, logits_mean_abs_diff=float(np.abs(sl - ref_sl).mean()),
                                 logits_mean_abs=float(np.abs(ref_sl).mean())))
    np.testing.assert_allclose(fp, g["loss_fp"], atol=2e-4 if dtype == torch.float16 else 2e-5, rtol=0)
    _check(g, rels, wbad, losses, float(np.abs(sl - ref_sl).mean()), dtype)
