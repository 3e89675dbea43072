#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own Python modules (pure PyTorch, CPU, fp32) with the
small shims SURVEY.md section 8c lists, feeds them seeded synthetic inputs and stores
inputs + outputs as .npz.  The reference ships no tests or golden vectors of its
own, so these files are what pins oracle/ and the CUDA path to the reference.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference/OutEffHop"
OUT = os.path.dirname(os.path.abspath(__file__))

warnings.filterwarnings("ignore")


def _install_shims():
    import transformers  # noqa: F401  (must be imported before timm is stubbed)
    import transformers.pytorch_utils as pu
    import transformers.modeling_utils as mu

    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward   # quantized_bert.py:28
    for name in ("find_pruneable_heads_and_indices", "prune_linear_layer"):
        if not hasattr(mu, name) and hasattr(pu, name):
            setattr(mu, name, getattr(pu, name))
    # hijacker.py:5-6 imports two timm activations only to put them in a set
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    act_me = types.ModuleType("timm.models.layers.activations_me")

    class Swish(torch.nn.Module):
        pass

    class SwishMe(torch.nn.Module):
        pass

    layers.Swish = Swish
    act_me.SwishMe = SwishMe
    timm.models, models.layers, layers.activations_me = models, layers, act_me
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                        "timm.models.layers.activations_me": act_me})
    sys.path.insert(0, REF)


def gen_softmax():
    from transformers_language.models.softmax import SOFTMAX_MAPPING

    g = torch.Generator().manual_seed(1001)
    rows = []
    for bias in (-8.0, 0.0, 8.0):
        rows.append(torch.randn(3, 37, generator=g) * 2.0 + bias)
    x = torch.cat(rows, 0)
    fmin = torch.finfo(torch.float32).min
    masked = x.clone()
    masked[:, 20:] = masked[:, 20:] + fmin          # HF extended mask on the tail
    allmasked = torch.full((1, 37), fmin)
    big = torch.tensor([[100.0] * 37, [-100.0] * 37, [-90.0] + [-95.0] * 36])
    x = torch.cat([x, masked, allmasked, big], 0)   # (22, 37)
    out = {"x": x.numpy()}
    keys, kinds, gammas, etas = [], [], [], []
    for key, fn in SOFTMAX_MAPPING.items():
        if key == "entmax":
            continue
        keys.append(key)
        kw = getattr(fn, "keywords", {}) or {}
        name = getattr(fn, "func", fn).__name__
        kinds.append({"softmax": 0, "softmax_1": 1, "clipped_softmax": 2, "clipped_softmax1": 3}[name])
        gammas.append(kw.get("gamma", 0.0))
        etas.append(kw.get("eta", 1.0))
        out["y::" + key] = fn(x, dim=-1).numpy()
    out["keys"] = np.array(keys)
    out["kinds"] = np.array(kinds, dtype=np.int32)
    out["gammas"] = np.array(gammas, dtype=np.float64)
    out["etas"] = np.array(etas, dtype=np.float64)
    out["all_keys"] = np.array(list(SOFTMAX_MAPPING.keys()))
    np.savez_compressed(os.path.join(OUT, "softmax.npz"), **out)
    print("softmax.npz", len(keys), "keys")


def _sd_numpy(mod):
    """state_dict as float16 arrays (parameters were rounded to fp16-representable values by
    _round_params_fp16, so this is lossless and halves the fixture size)."""
    return {k: v.detach().numpy().astype(np.float16) for k, v in mod.state_dict().items()}


def _round_params_fp16(mod):
    """Make every parameter exactly representable in fp16 so the fp16 CUDA path and the fp32
    reference start from bit-identical weights."""
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(p.half().float())
    return mod


def _randn_fp16(*shape, scale=1.0):
    return (torch.randn(*shape) * scale).half().float()


_BASE = {}


def _load_base(mod, family):
    """Share one set of q/k/v(/out) projection weights per model family across all cases."""
    names = [n for n, _ in mod.named_parameters() if not n.startswith("alpha")]
    if family not in _BASE:
        _BASE[family] = {n: p.detach().clone() for n, p in mod.named_parameters() if n in names}
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n in _BASE[family]:
                p.copy_(_BASE[family][n])
    return mod


def gen_attention_modules():
    from transformers import BertConfig
    from transformers_language.models.softmax import SOFTMAX_MAPPING
    from transformers_language.models.bert_attention import BertSelfAttentionWithExtras, AttentionGateType
    from transformers_language.models.opt_attention import OPTAttentionWithExtras

    out = {}
    cfg = BertConfig(hidden_size=128, num_attention_heads=2, attention_probs_dropout_prob=0.0)
    fmin = torch.finfo(torch.float32).min
    softmax_keys = ["vanilla", "softmax1", "clipped(-.025:1)", "clippedsoftmax1(-.025:1)", "clipped(0:1.03)"]

    def make_bert(key, gated, **kw):
        gt = AttentionGateType.conditional_per_token if gated else AttentionGateType.none
        m = BertSelfAttentionWithExtras(cfg, softmax_fn=SOFTMAX_MAPPING[key], attn_gate_type=gt,
                                        attn_gate_init=0.25 if gated else None, **kw).eval()
        if gated and isinstance(m.alpha, torch.nn.ModuleList):   # make the gate data dependent
            for fc in m.alpha:
                for lin in ([fc] if isinstance(fc, torch.nn.Linear) else [l for l in fc if isinstance(l, torch.nn.Linear)]):
                    torch.nn.init.normal_(lin.weight, std=0.3)
        return _round_params_fp16(_load_base(m, "bert"))

    # (key, gated, masked): all five keys plain and gated+masked, plus the two mixed cases for softmax1
    combos = [(k, g, g) for k in softmax_keys for g in (False, True)] + [("softmax1", False, True), ("softmax1", True, False)]
    cases = []
    for idx, (key, gated, masked) in enumerate(combos):
        torch.manual_seed(2000 + idx)
        B, T = 2, (64 if not masked else 50)
        m = make_bert(key, gated)
        hidden = _randn_fp16(B, T, 128, scale=1.5)
        mask = None
        if masked:  # HF-4.31 extended mask (B,1,1,T): pad the tail of sample 1
            mask = torch.zeros(B, 1, 1, T)
            mask[1, :, :, 37:] = fmin
        with torch.no_grad():
            y = m(hidden, attention_mask=mask)[0]
        tag = f"bert{idx}"
        cases.append((tag, key, int(gated), int(masked)))
        out[tag + "::hidden"] = hidden.numpy().astype(np.float16)
        out[tag + "::out"] = y.numpy()
        if mask is not None:
            out[tag + "::mask"] = mask.numpy()
        for k, v in _sd_numpy(m).items():
            if k.startswith("alpha") or idx == 0:
                out[(tag if k.startswith("alpha") else "bert_base") + "::sd::" + k] = v
    out["bert_cases"] = np.array(cases)

    # gate predictor variants used by the BERT gate scripts (--attn_gate_mlp) and the bigger MLP
    for tag, kw in (("bert_mlp", dict(attn_gate_mlp=True)), ("bert_mlp2", dict(attn_gate_mlp2=True))):
        torch.manual_seed(2500 + len(tag))
        m = make_bert("softmax1", True, **kw)
        hidden = _randn_fp16(2, 64, 128, scale=1.5)
        with torch.no_grad():
            y = m(hidden)[0]
        out[tag + "::hidden"] = hidden.numpy().astype(np.float16)
        out[tag + "::out"] = y.numpy()
        for k, v in _sd_numpy(m).items():
            if k.startswith("alpha"):
                out[tag + "::sd::" + k] = v

    # alpha override (bert_attention.py:88-94): gamma = -alpha / max_seq_length, eta = 1
    torch.manual_seed(3500)
    m = _round_params_fp16(_load_base(BertSelfAttentionWithExtras(cfg, alpha=4.0, max_seq_length=64).eval(), "bert"))
    hidden = _randn_fp16(2, 64, 128)
    with torch.no_grad():
        y = m(hidden)[0]
    out["bert_alpha::hidden"] = hidden.numpy().astype(np.float16)
    out["bert_alpha::out"] = y.numpy()

    # OPT: causal HF-4.31 mask (B,1,T,T) of 0 / finfo.min, fp32 (the reference's eval precision)
    cases = []
    for idx, (key, gated) in enumerate([(k, bool(i % 2)) for i, k in enumerate(softmax_keys)] + [("softmax1", False)]):
        torch.manual_seed(3000 + idx)
        B, T = 2, 72
        gt = AttentionGateType.conditional_per_token if gated else AttentionGateType.none
        m = OPTAttentionWithExtras(embed_dim=128, num_heads=2, is_decoder=True,
                                   softmax_fn=SOFTMAX_MAPPING[key], attn_gate_type=gt,
                                   attn_gate_init=0.25 if gated else None).eval()
        if gated:
            for fc in m.alpha:
                torch.nn.init.normal_(fc.weight, std=0.3)
        m = _round_params_fp16(_load_base(m, "opt"))
        hidden = _randn_fp16(B, T, 128, scale=1.5)
        mask = torch.full((T, T), fmin).triu(1)[None, None].expand(B, 1, T, T).contiguous()
        with torch.no_grad():
            y = m(hidden, attention_mask=mask)[0]
        tag = f"opt{idx}"
        cases.append((tag, key, int(gated)))
        out[tag + "::hidden"] = hidden.numpy().astype(np.float16)
        out[tag + "::out"] = y.numpy()
        for k, v in _sd_numpy(m).items():
            if k.startswith("alpha") or idx == 0:
                out[(tag if k.startswith("alpha") else "opt_base") + "::sd::" + k] = v
    out["opt_cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "attention_modules.npz"), **out)
    print("attention_modules.npz", len(out))


def gen_quantizer():
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    from quantization.quantization_manager import QuantizationManager
    from quantization.base_quantized_classes import QuantizedActivation

    out = {}
    AQ = QMethods.asymmetric_uniform.cls
    g = torch.Generator().manual_seed(4001)
    ranges = [(-1.0, 3.0), (0.0, 1.0), (-7.3, 0.0), (-0.02, 0.57), (-50.0, 120.0), (0.25, 0.75), (0.0, 0.0)]
    qcases = []
    for n_bits in (8, 4):
        for ri, (lo, hi) in enumerate(ranges):
            q = AQ(n_bits=n_bits)
            q.set_quant_range(lo, hi)
            delta, zf = q._delta.clone(), q._zero_float.clone()
            scale = q.scale
            zp = q.zero_point
            # random values + exact half-way points + saturating outliers
            x = torch.randn(2048, generator=g) * (hi - lo + 1e-3) * 0.6 + (hi + lo) / 2
            k = torch.arange(-8, 2 ** n_bits + 8, dtype=torch.float32)
            half = (k + 0.5 - zp) * scale
            exact = (k - zp) * scale
            x = torch.cat([x, half, exact, torch.tensor([0.0, -0.0, 1e30, -1e30, 1e-30])])
            tag = f"q{n_bits}_{ri}"
            qcases.append((tag, n_bits, lo, hi))
            out[tag + "::x"] = x.numpy()
            out[tag + "::delta"] = delta.numpy()
            out[tag + "::zero_float"] = zf.numpy()
            out[tag + "::scale"] = scale.numpy()
            out[tag + "::zero_point"] = zp.numpy()
            out[tag + "::x_int"] = q.to_integer_forward(x).numpy()
            out[tag + "::x_q"] = q(x).numpy()
            if n_bits == 8 and ri < 4:   # fp16 tensors through the same quantizer
                xh = x.half()
                out[tag + "::x_h"] = xh.numpy()
                out[tag + "::x_q_h"] = q(xh).numpy()
                out[tag + "::x_int_h"] = q.to_integer_forward(xh).numpy()
                xb = x.bfloat16()
                out[tag + "::x_b"] = xb.float().numpy()
                out[tag + "::x_q_b"] = q(xb).float().numpy()
    out["qcases"] = np.array(qcases)

    # running min/max sequences through QuantizationManager in estimate mode, then fixed
    seqs = []
    for si in range(3):
        qm = QuantizationManager(qmethod=AQ, init=RangeEstimators.running_minmax.cls,
                                 qparams=dict(n_bits=8, scale_domain="linear", act_quant=True))
        tag = f"rm{si}"
        xs, ys, mins, maxs, deltas, zfs = [], [], [], [], [], []
        for step in range(5):
            x = torch.randn(3, 257, generator=g) * (1.0 + step * 0.7) + (si - 1) * 2.0
            if step == 2:
                x[0, 0] = 40.0   # outlier batch
            y = qm(x)
            xs.append(x.numpy()); ys.append(y.numpy())
            mins.append(qm.range_estimator.current_xmin.numpy().copy())
            maxs.append(qm.range_estimator.current_xmax.numpy().copy())
            deltas.append(qm.quantizer._delta.numpy().copy())
            zfs.append(qm.quantizer._zero_float.numpy().copy())
        qm.fix_ranges()
        xf = torch.randn(3, 257, generator=g) * 3.0
        out[tag + "::x"] = np.stack(xs); out[tag + "::y"] = np.stack(ys)
        out[tag + "::cur_min"] = np.stack(mins); out[tag + "::cur_max"] = np.stack(maxs)
        out[tag + "::delta"] = np.stack(deltas); out[tag + "::zero_float"] = np.stack(zfs)
        out[tag + "::x_fixed"] = xf.numpy(); out[tag + "::y_fixed"] = qm(xf).numpy()
        seqs.append(tag)
    out["rm_cases"] = np.array(seqs)

    # the SURVEY section-4 KAT: [-1,2],[-3,4],[.5,1]
    from quantization.range_estimators import RunningMinMaxEstimator
    est = RunningMinMaxEstimator()
    kat = []
    for lo, hi in ((-1.0, 2.0), (-3.0, 4.0), (0.5, 1.0)):
        a, b = est(torch.tensor([lo, hi]))
        kat.append((a.item(), b.item()))
    out["rm_kat"] = np.array(kat, dtype=np.float32)

    # QuantizedActivation passthrough / active
    qa = QuantizedActivation(method=AQ, act_method=AQ, n_bits=8, n_bits_act=8)
    x = torch.randn(4, 33, generator=g)
    out["qa::x"] = x.numpy()
    out["qa::y_fp"] = qa(x).numpy()
    qa.quantized_acts()
    out["qa::y_q"] = qa(x).numpy()
    np.savez_compressed(os.path.join(OUT, "quantizer.npz"), **out)
    print("quantizer.npz", len(qcases), "quantizer cases")


def gen_quantized_attention():
    """QuantizedBertSelfAttentionWithExtras / QuantizedOPTAttentionWithExtras:
    3 estimation batches (running_minmax) then one fixed-range forward."""
    import transformers.models.opt.modeling_opt as mo
    for name in ("_expand_mask", "_make_causal_mask"):      # quantized_opt.py:8-14 import-time only
        if not hasattr(mo, name):
            setattr(mo, name, lambda *a, **k: None)
    from transformers import BertConfig
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    from transformers_language.models.softmax import SOFTMAX_MAPPING
    from transformers_language.models.bert_attention import BertSelfAttentionWithExtras, AttentionGateType
    from transformers_language.models.opt_attention import OPTAttentionWithExtras
    from transformers_language.models.quantized_bert import QuantizedBertSelfAttentionWithExtras
    from transformers_language.models.quantized_opt import QuantizedOPTAttentionWithExtras

    qparams = dict(method=QMethods.symmetric_uniform.cls, n_bits=8, n_bits_act=8,
                   act_method=QMethods.asymmetric_uniform.cls, per_channel_weights=False, percentile=None,
                   quant_setup="all", weight_range_method=RangeEstimators.current_minmax.cls,
                   weight_range_options={}, act_range_method=RangeEstimators.running_minmax.cls,
                   act_range_options={})
    out = {}
    fmin = torch.finfo(torch.float32).min
    cfg = BertConfig(hidden_size=128, num_attention_heads=2, attention_probs_dropout_prob=0.0)

    def keep(k):   # quantizer / estimator buffers only; the weights are the shared fp module's
        return any(t in k for t in ("_delta", "_zero_float", "current_xmin", "current_xmax", "_signed"))

    def run(tag, qm, fp, key, gated, T, mask):
        qm.set_quant_state(weight_quant=True, act_quant=True)
        xs, ys = [], []
        with torch.no_grad():
            for step in range(3):
                x = _randn_fp16(2, T, 128, scale=1.0 + 0.5 * step)
                xs.append(x.numpy().astype(np.float16))
                ys.append((qm(x, attention_mask=mask) if mask is not None else qm(x))[0].numpy())
            qm.fix_ranges()
            xf = _randn_fp16(2, T, 128, scale=1.3)
            yf = (qm(xf, attention_mask=mask) if mask is not None else qm(xf))[0]
        out[tag + "::key"] = np.array(key); out[tag + "::gated"] = np.array(int(gated))
        out[tag + "::x_est"] = np.stack(xs); out[tag + "::y_est"] = np.stack(ys)
        out[tag + "::x_fixed"] = xf.numpy().astype(np.float16); out[tag + "::y_fixed"] = yf.numpy()
        for k, v in qm.state_dict().items():
            if keep(k):
                out[tag + "::sd::" + k] = v.detach().numpy()
        for k, v in _sd_numpy(fp).items():
            if k.startswith("alpha"):
                out[tag + "::fp::" + k] = v

    for ci, (key, gated) in enumerate((("softmax1", False), ("clippedsoftmax1(-.025:1)", True), ("vanilla", True))):
        torch.manual_seed(5000 + ci)
        gt = AttentionGateType.conditional_per_token if gated else AttentionGateType.none
        fp = BertSelfAttentionWithExtras(cfg, softmax_fn=SOFTMAX_MAPPING[key], attn_gate_type=gt,
                                         attn_gate_init=0.25 if gated else None).eval()
        if gated:
            for fc in fp.alpha:
                torch.nn.init.normal_(fc.weight, std=0.3)
        fp = _round_params_fp16(_load_base(fp, "bert"))
        qm = QuantizedBertSelfAttentionWithExtras(fp, quant_dict={}, **qparams).eval()
        run(f"qbert{ci}", qm, fp, key, gated, 48, None)
    for ci, (key, gated) in enumerate((("softmax1", False), ("vanilla", True))):
        torch.manual_seed(6000 + ci)
        gt = AttentionGateType.conditional_per_token if gated else AttentionGateType.none
        fp = OPTAttentionWithExtras(embed_dim=128, num_heads=2, is_decoder=True, softmax_fn=SOFTMAX_MAPPING[key],
                                    attn_gate_type=gt, attn_gate_init=0.25 if gated else None).eval()
        if gated:
            for fc in fp.alpha:
                torch.nn.init.normal_(fc.weight, std=0.3)
        fp = _round_params_fp16(_load_base(fp, "opt"))
        qm = QuantizedOPTAttentionWithExtras(fp, **qparams).eval()
        T = 40
        mask = torch.full((T, T), fmin).triu(1)[None, None].expand(2, 1, T, T).contiguous()
        run(f"qopt{ci}", qm, fp, key, gated, T, mask)
    np.savez_compressed(os.path.join(OUT, "quantized_attention.npz"), **out)
    print("quantized_attention.npz", len(out))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    torch.set_num_threads(1)   # deterministic reduction order
    _install_shims()
    gen_softmax()
    gen_attention_modules()
    gen_quantizer()
    gen_quantized_attention()
