#!/usr/bin/env python
"""Golden fixtures for the SURVEY section 8f rows, generated from the LIVE reference (build container only):

    python tests/golden/make_golden_next.py      -> tests/golden/next_rows.npz

* percentile range estimation: RunningMinMaxEstimator / CurrentMinMaxEstimator with ``percentile``
  (quantization/range_estimators.py:62-69, :89-94) behind a QuantizationManager with the asymmetric quantizer;
* QuantizedBertSelfOutput (quantized_bert.py:540-569): dense -> residual add -> tap -> LayerNorm, 3 estimation batches
  and one fixed-range forward;
* kurtosis() and the per-sample inf-norm (transformers_language/utils.py:9-20, validate_mlm_config.py:513-518);
* ViTSelfAttentionWithExtras (vit_attention.py:77-275) and STanHop's Association (cross_models/hopfield.py:19-51).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402

OUT = G.OUT
STANHOP = "/root/reference/STanHop_time_seeries"


def gen_percentile(out):
    from quantization.quantization_manager import QuantizationManager
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators

    g = torch.Generator().manual_seed(7001)
    for ci, (pct, shape) in enumerate(((99.999, (2, 4, 48, 64)), (99.9, (3, 1000)), (75.0, (257,)))):
        mgr = QuantizationManager(qmethod=QMethods.asymmetric_uniform.cls, init=RangeEstimators.running_minmax.cls,
                                  qparams=dict(n_bits=8, act_quant=True), init_params=dict(percentile=pct))
        xs, ys, ranges = [], [], []
        for step in range(3):
            x = torch.randn(*shape, generator=g) * (1.0 + step)
            flat = x.view(-1)
            flat[::97] *= 25.0                                    # outliers: the reason percentile ranges exist
            xs.append(x.numpy().copy())
            ys.append(mgr(x).numpy())
            est = mgr.range_estimator
            ranges.append([float(est.current_xmin), float(est.current_xmax), float(mgr.quantizer._delta),
                           float(mgr.quantizer._zero_float)])
            assert est.current_xmin.dtype == torch.float64
        mgr.fix_ranges()
        xf = torch.randn(*shape, generator=g) * 2.0
        out[f"pct{ci}::percentile"] = np.array(pct)
        out[f"pct{ci}::x"] = np.stack(xs)
        out[f"pct{ci}::y"] = np.stack(ys)
        out[f"pct{ci}::ranges"] = np.array(ranges, dtype=np.float64)
        out[f"pct{ci}::x_fixed"] = xf.numpy()
        out[f"pct{ci}::y_fixed"] = mgr(xf).numpy()
    # CurrentMinMaxEstimator keeps the reference's (p, 100-p) order
    est = RangeEstimators.current_minmax.cls(percentile=99.0)
    x = torch.randn(5000, generator=g)
    lo, hi = est(x)
    out["pct_cur::x"] = x.numpy()
    out["pct_cur::range"] = np.array([float(lo), float(hi)], dtype=np.float64)


def gen_residual_block(out):
    from transformers import BertConfig
    from transformers.models.bert.modeling_bert import BertSelfOutput
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    from transformers_language.models.quantized_bert import QuantizedBertSelfOutput

    qparams = dict(method=QMethods.symmetric_uniform.cls, n_bits=8, n_bits_act=8,
                   act_method=QMethods.asymmetric_uniform.cls, per_channel_weights=False, percentile=None,
                   quant_setup="all", weight_range_method=RangeEstimators.current_minmax.cls,
                   weight_range_options={}, act_range_method=RangeEstimators.running_minmax.cls,
                   act_range_options={})
    torch.manual_seed(7100)
    cfg = BertConfig(hidden_size=128, num_attention_heads=2, hidden_dropout_prob=0.0)
    fp = BertSelfOutput(cfg).eval()
    with torch.no_grad():
        fp.LayerNorm.weight.copy_(1.0 + 0.1 * torch.randn(128))
        fp.LayerNorm.bias.copy_(0.1 * torch.randn(128))
    qm = QuantizedBertSelfOutput(fp, quant_dict={}, **qparams).eval()
    qm.set_quant_state(weight_quant=True, act_quant=True)
    hs, ins, ys = [], [], []
    with torch.no_grad():
        for step in range(3):
            h, i = torch.randn(2, 24, 128) * (1 + step), torch.randn(2, 24, 128)
            hs.append(h.numpy()); ins.append(i.numpy()); ys.append(qm(h, i).numpy())
        qm.fix_ranges()
        hf, inf_ = torch.randn(2, 24, 128) * 1.5, torch.randn(2, 24, 128)
        yf = qm(hf, inf_)
    out["res::h_est"], out["res::in_est"], out["res::y_est"] = np.stack(hs), np.stack(ins), np.stack(ys)
    out["res::h_fixed"], out["res::in_fixed"], out["res::y_fixed"] = hf.numpy(), inf_.numpy(), yf.numpy()
    for k, v in fp.state_dict().items():
        out["res::fp::" + k] = v.numpy()
    for k, v in qm.state_dict().items():
        if any(t in k for t in ("_delta", "_zero_float", "current_xmin", "current_xmax", "_signed")):
            out["res::sd::" + k] = v.detach().numpy()


def gen_outlier(out):
    from transformers_language.utils import kurtosis
    g = torch.Generator().manual_seed(7200)
    x = torch.randn(6, 40, 96, generator=g)
    x[1, 3, 5] = 60.0
    x[4, :, 17] *= 12.0
    x[5] += 3.0                                                    # non-zero mean
    xv = x.view(x.size(0), -1)
    out["outlier::x"] = x.numpy()
    out["outlier::inf"] = xv.norm(dim=1, p=np.inf).numpy()
    out["outlier::kurt"] = kurtosis(xv).numpy()


def gen_vit(out):
    from transformers_language.models.softmax import SOFTMAX_MAPPING
    from transformers_language.models.vit_attention import ViTSelfAttentionWithExtras, AttentionGateType
    for ci, (key, gated) in enumerate((("softmax1", False), ("clippedsoftmax1(-.025:1)", False), ("vanilla", True))):
        torch.manual_seed(7300 + ci)
        gt = AttentionGateType.conditional_per_token if gated else AttentionGateType.none
        m = ViTSelfAttentionWithExtras(dim=128, num_heads=2, qkv_bias=True, softmax_fn=SOFTMAX_MAPPING[key],
                                       attn_gate_type=gt, attn_gate_init=0.25 if gated else None).eval()
        if gated:
            for fc in m.alpha:
                torch.nn.init.normal_(fc.weight, std=0.3)
        G._round_params_fp16(m)
        x = G._randn_fp16(2, 197, 128)
        with torch.no_grad():
            y = m(x)
        out[f"vit{ci}::key"], out[f"vit{ci}::gated"] = np.array(key), np.array(int(gated))
        out[f"vit{ci}::x"], out[f"vit{ci}::y"] = x.numpy().astype(np.float16), y.numpy()
        for k, v in G._sd_numpy(m).items():
            out[f"vit{ci}::sd::" + k] = v


def gen_hopfield(out):
    sys.path.insert(0, STANHOP)
    from cross_models.hopfield import Association
    for ci, mode in enumerate(("softmax", "softmax1", "clip")):
        torch.manual_seed(7400 + ci)
        m = Association(scale=None, attention_dropout=0.0, eta=1.1, gamma=-0.1, mode=mode).eval()
        q, k, v = G._randn_fp16(2, 33, 2, 64), G._randn_fp16(2, 50, 2, 64), G._randn_fp16(2, 50, 2, 64)
        with torch.no_grad():
            y = m(q, k, v)
        out[f"hop{ci}::mode"] = np.array(mode)
        out[f"hop{ci}::q"], out[f"hop{ci}::k"], out[f"hop{ci}::v"] = (t.numpy().astype(np.float16) for t in (q, k, v))
        out[f"hop{ci}::y"] = y.numpy()
    # mode='clip_softmax1' raises in the reference (ClipSoftmax_1.__init__, clip_softmax.py:52); record that fact and
    # the function the module was meant to apply
    try:
        Association(mode="clip_softmax1")
        out["hop_clip1::ctor_raises"] = np.array(0)
    except TypeError:
        out["hop_clip1::ctor_raises"] = np.array(1)
    from cross_models.clip_softmax import clipped_softmax_1
    q, k, v = G._randn_fp16(2, 33, 2, 64), G._randn_fp16(2, 50, 2, 64), G._randn_fp16(2, 50, 2, 64)
    scores = torch.einsum("blhe,bshe->bhls", q, k)
    A = clipped_softmax_1(scores / 8.0, dim=-1, eta=1.1, gamma=-0.1)
    out["hop_clip1::q"], out["hop_clip1::k"], out["hop_clip1::v"] = (t.numpy().astype(np.float16) for t in (q, k, v))
    out["hop_clip1::y"] = torch.einsum("bhls,bshd->blhd", A, v).contiguous().numpy()


if __name__ == "__main__":
    if not os.path.isdir(G.REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    torch.set_num_threads(1)
    G._install_shims()
    out = {}
    gen_percentile(out)
    gen_residual_block(out)
    gen_outlier(out)
    gen_vit(out)
    gen_hopfield(out)
    np.savez_compressed(os.path.join(OUT, "next_rows.npz"), **out)
    print("next_rows.npz", len(out), os.path.getsize(os.path.join(OUT, "next_rows.npz")))
