"""CPU-side checks: the C-ABI library loads and exports every symbol include/oeh.h declares, host-side
logic (CLI-string table, enums, host range math, error conventions) matches the reference."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from outeffhop_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "oeh.h")).read()
    declared = set(re.findall(r"\b(oeh_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.oeh_version() == 100
    assert handle.oeh_last_error() is not None


def test_attn_args_struct_layout_matches_header():
    """ctypes mirror of OehAttnArgs: field order/size sanity (8-byte pointers, natural alignment)."""
    from outeffhop_b200._lib import OehAttnArgs
    assert ctypes.sizeof(OehAttnArgs) % 8 == 0
    assert OehAttnArgs.q.offset == 0 and OehAttnArgs.q_strides.offset == 32 and OehAttnArgs.B.offset == 128
    assert OehAttnArgs.gamma.offset == 160 and OehAttnArgs.eta.offset == 168 and OehAttnArgs.mask.offset == 184


def test_host_set_quant_range_kat():
    from outeffhop_b200 import ops
    d, zf, s, zp = ops.set_quant_range_host(-1.0, 3.0, 8)
    assert np.float32(d) == np.float32(0.015686275) and zp == 64.0 and abs(zf - 63.749996) < 1e-5
    d, zf, s, zp = ops.set_quant_range_host(0.25, 0.75, 8)          # x_min is clamped to <= 0
    assert zp == 0.0 and np.float32(d) == np.float32(0.75) / np.float32(255)
    d, zf, s, zp = ops.set_quant_range_host(0.0, 0.0, 8)            # x_max is clamped to >= eps
    assert np.float32(s) == np.float32(1e-8)


def test_host_bad_args_map_to_value_error():
    from outeffhop_b200 import _lib
    out = (ctypes.c_float * 4)()
    rc = _lib.lib().oeh_set_quant_range_host(0.0, 1.0, 99, 1e-8, out)
    assert rc == _lib.OEH_ERR_BAD_ARG and "n_bits" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_softmax_mapping_table_matches_reference(golden):
    from outeffhop_b200.softmax import SOFTMAX_MAPPING, SoftmaxFn, parse_attn_softmax
    g = golden("softmax.npz")
    assert list(SOFTMAX_MAPPING.keys()) == [str(k) for k in g["all_keys"]]          # 40 keys, same order
    for key, kind, gamma, eta in zip(g["keys"], g["kinds"], g["gammas"], g["etas"]):
        fn = SOFTMAX_MAPPING[str(key)]
        assert isinstance(fn, SoftmaxFn) and (fn.kind, fn.gamma, fn.eta) == (int(kind), float(gamma), float(eta)), key
    assert SOFTMAX_MAPPING["clipped(-.005:1.005)"].gamma == -0.003                  # table quirk softmax.py:57
    assert SOFTMAX_MAPPING["clippedsoftmax1(-.025:1)"].eta == 1.1                   # table quirk softmax.py:61
    fn = parse_attn_softmax("clippedsoftmax1(-.05:1.2)")
    assert (fn.kind, fn.gamma, fn.eta) == (3, -0.05, 1.2)
    assert parse_attn_softmax("clipped(-.025:1)") is SOFTMAX_MAPPING["clipped(-.025:1)"]
    with pytest.raises(KeyError):
        parse_attn_softmax("sparsemax")


def test_as_softmax_fn_recognises_reference_style_callables():
    from functools import partial
    from outeffhop_b200 import softmax as S
    assert S.as_softmax_fn(torch.nn.functional.softmax).kind == 0
    assert S.as_softmax_fn(S.softmax_1).kind == 1
    fn = S.as_softmax_fn(partial(S.clipped_softmax1, gamma=-0.01, eta=1.0))
    assert (fn.kind, fn.gamma, fn.eta) == (3, -0.01, 1.0)
    assert S.as_softmax_fn(partial(S.clipped_softmax)).gamma == -0.1              # reference defaults eta=1.1, gamma=-0.1
    assert S.as_softmax_fn(lambda x, dim=-1: x) is None
    assert S.as_softmax_fn(S.SOFTMAX_MAPPING["entmax"]) is None


def test_enums_and_logit():
    from outeffhop_b200 import AttentionGateType, logit
    from outeffhop_b200.quantization import QMethods, Qstates, RangeEstimators
    assert AttentionGateType.list_names() == ["none", "unconditional_per_head", "conditional_per_head",
                                              "conditional_per_token"]
    assert AttentionGateType["conditional_per_token"].value == 3 and str(AttentionGateType.none) == "none"
    assert abs(logit(0.25) + 1.0986122886681098) < 1e-12
    assert QMethods.asymmetric_uniform.cls.__name__ == "AsymmetricUniformQuantizer"
    assert RangeEstimators.running_minmax.cls.__name__ == "RunningMinMaxEstimator"
    assert Qstates.list_names() == ["estimate_ranges", "fix_ranges", "learn_ranges", "estimate_ranges_train"]


def test_module_construction_and_state_dict_keys_without_gpu():
    from transformers import BertConfig
    import outeffhop_b200 as o
    cfg = BertConfig(hidden_size=128, num_attention_heads=2)
    m = o.BertSelfAttentionWithExtras(cfg, softmax_fn=o.SOFTMAX_MAPPING["softmax1"],
                                      attn_gate_type=o.AttentionGateType.conditional_per_token, attn_gate_init=0.25,
                                      attn_gate_mlp=True)
    keys = set(m.state_dict())
    assert {"query.weight", "key.bias", "value.weight", "alpha.0.0.weight", "alpha.1.2.bias"} <= keys
    m2 = o.BertSelfAttentionWithExtras(cfg, attn_gate_type=o.AttentionGateType.conditional_per_token, attn_gate_init=0.25)
    assert abs(m2.alpha[0].bias.item() + 1.0986123) < 1e-6                         # logit(0.25)
    m3 = o.BertSelfAttentionWithExtras(cfg, attn_gate_type=o.AttentionGateType.unconditional_per_head)
    assert m3.alpha.shape == (2,)
    op = o.OPTAttentionWithExtras(128, 2, is_decoder=True, alpha=12.0, max_seq_length=512, attn_softmax="softmax1")
    assert op.softmax_fn.kind == 3 and op.softmax_fn.gamma == -12.0 / 512          # `==`, not the reference's `is`
    assert {"q_proj.weight", "k_proj.bias", "v_proj.weight", "out_proj.bias"} <= set(op.state_dict())
    with pytest.raises(RuntimeError):                                              # no CPU fallback
        with torch.no_grad():
            m(torch.randn(1, 4, 128))
    with pytest.raises(NotImplementedError):
        o.AsymmetricUniformQuantizer(n_bits=8, per_channel=True)
    est = o.RunningMinMaxEstimator(percentile=99.999)                              # range_estimators.py:89-94, now on device
    assert est.percentile == 99.999


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from outeffhop_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.lib()


# ---- host logic of the eval-step fusions (CPU tensors: these helpers are pure tensor plumbing) ------------------------
def test_fused_qkv_projection_views_cache_and_opt_outs():
    import torch
    from torch import nn
    from outeffhop_b200.bert_attention import fused_qkv_projection
    torch.manual_seed(0)
    H, d = 4, 8
    owner = nn.Module().eval()
    mods = tuple(nn.Linear(H * d, H * d) for _ in range(3))
    x = torch.randn(2, 5, H * d)
    with torch.no_grad():
        q, k, v = fused_qkv_projection(owner, mods, x, H, d)
        for got, m in zip((q, k, v), mods):
            want = m(x).view(2, 5, H, d).permute(0, 2, 1, 3)
            assert got.shape == want.shape and got.stride(3) == 1
            assert torch.allclose(got, want, atol=1e-6)
        cache = owner.__dict__["_qkv_cache"]
        fused_qkv_projection(owner, mods, x, H, d)
        assert owner.__dict__["_qkv_cache"] is cache                      # parameters untouched: cache reused
        mods[1].weight.mul_(2.0)                                          # in-place update bumps ._version
        q2, k2, v2 = fused_qkv_projection(owner, mods, x, H, d)
        assert owner.__dict__["_qkv_cache"] is not cache
        assert torch.allclose(k2, mods[1](x).view(2, 5, H, d).permute(0, 2, 1, 3), atol=1e-6)
    assert "_qkv_cache" not in dict(owner.named_buffers()) and not list(owner.parameters())
    mods[0].register_forward_hook(lambda *a: None)                        # somebody observes `query`: no fusion
    assert fused_qkv_projection(owner, mods, x, H, d) is None
    assert fused_qkv_projection(nn.Module().train(), tuple(nn.Linear(4, 4) for _ in range(3)), torch.randn(1, 1, 4), 1, 4) is None


def test_padded_decoder_is_zero_padded_and_cached():
    import torch
    from torch import nn
    from outeffhop_b200.eval_harness import _padded_decoder
    lin = nn.Linear(16, 30522 % 1000 + 3)                                 # 525 outputs -> padded to 576
    w, b, V = _padded_decoder(lin)
    assert V == 525 and w.shape == (576, 16) and b.shape == (576,)
    assert torch.equal(w[:V], lin.weight) and torch.equal(b[:V], lin.bias)
    assert float(w[V:].abs().sum()) == 0.0 and float(b[V:].abs().sum()) == 0.0
    assert _padded_decoder(lin)[0] is w
    aligned = nn.Linear(16, 128)
    assert _padded_decoder(aligned)[0] is aligned.weight                  # already a multiple of 64: no copy
