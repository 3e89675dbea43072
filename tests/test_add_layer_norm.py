"""Fused residual add + LayerNorm (+ fake-quant taps) -- SURVEY 8f rank 2 -- against the oracle and ATen.
Tolerance: the kernel rounds the sum to the tensor dtype like ATen and normalises in fp32; results agree with
F.layer_norm(a + b) to 1 ulp of the tensor dtype (fp32: 2e-6 relative to the row scale)."""
import numpy as np
import pytest
import torch

from oracle import oeh_oracle as O  # checker only

_T = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}
_TOL = {"float32": 3e-6, "float16": 2e-3, "bfloat16": 1.6e-2}


def test_oracle_add_layer_norm_equals_aten_fp32():
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(37, 768, generator=g) * 2, torch.randn(37, 768, generator=g)
    w, bias = torch.rand(768, generator=g) + 0.5, torch.randn(768, generator=g) * 0.1
    want = torch.nn.functional.layer_norm(a + b, (768,), w, bias, 1e-12).numpy()
    got = O.add_layer_norm(a.numpy(), b.numpy(), w.numpy(), bias.numpy(), 1e-12)
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-6)


def test_symbol_declared():
    import os
    from outeffhop_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "oeh.h")).read()
    for name in ("oeh_add_layer_norm", "oeh_tap_add_fake_quant", "oeh_fake_quant_asym_segments"):
        assert name + "(" in hdr and name in _lib.SYMBOLS
        getattr(_lib.lib(), name)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("rows,cols", [(1, 8), (5, 64), (4096, 768), (33, 1024), (7, 2048), (64, 3072 // 4)])
@pytest.mark.parametrize("residual", [True, False])
def test_add_layer_norm_vs_aten_and_oracle(dtype, rows, cols, residual):
    from outeffhop_b200 import ops
    if dtype == "float32" and cols > 1024:
        pytest.skip("fp32 rows hold at most 1024 columns in registers")
    t = _T[dtype]
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + cols)
    a = (torch.randn(rows, cols, device="cuda", generator=g) * 3).to(t)
    b = torch.randn(rows, cols, device="cuda", generator=g).to(t) if residual else None
    w = (torch.rand(cols, device="cuda", generator=g) + 0.5).to(t)
    bias = (torch.randn(cols, device="cuda", generator=g) * 0.2).to(t)
    got = ops.add_layer_norm(a, b, w, bias, 1e-12)
    x = a + b if residual else a
    want = torch.nn.functional.layer_norm(x, (cols,), w, bias, 1e-12)
    assert got.dtype == t and got.shape == a.shape
    assert (got.float() - want.float()).abs().max().item() <= _TOL[dtype] * max(1.0, want.float().abs().max().item())
    f = lambda z: z.float().cpu().numpy()
    orc = O.add_layer_norm(f(x), None, f(w), f(bias), 1e-12)
    assert np.abs(f(got) - orc).max() <= _TOL[dtype] * max(1.0, np.abs(orc).max())


@pytest.mark.gpu
def test_add_layer_norm_with_taps_fp32_matches_unfused_chain():
    """INT8 residual block, fixed ranges: one kernel == add_fake_quant -> F.layer_norm -> fake_quant (bit-exact taps;
    LayerNorm itself to 1 fp32 ulp, so at most a few output codes may differ by one bin)."""
    from outeffhop_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    a, b = torch.randn(512, 768, device="cuda", generator=g) * 2, torch.randn(512, 768, device="cuda", generator=g)
    w, bias = torch.rand(768, device="cuda", generator=g) + 0.5, torch.randn(768, device="cuda", generator=g) * 0.1
    q_res = torch.tensor([0.05, 131.0], device="cuda")
    q_out = torch.tensor([0.03, 120.0], device="cuda")
    got = ops.add_layer_norm(a, b, w, bias, 1e-12, qparams_res=q_res, qparams_out=q_out)
    s = ops.add_fake_quant(a, b, q_res)
    ln = torch.nn.functional.layer_norm(s, (768,), w, bias, 1e-12)
    want = ops.fake_quant_asym(ln, q_out)
    diff = (got - want).abs()
    assert diff.max().item() <= 0.03 * 1.001 and (diff > 0).float().mean().item() < 1e-3
    orc = O.add_layer_norm(a.cpu().numpy(), b.cpu().numpy(), w.cpu().numpy(), bias.cpu().numpy(), 1e-12,
                           q_res=(np.float32(0.05), np.float32(131.0)), q_out=(np.float32(0.03), np.float32(120.0)))
    d2 = np.abs(got.cpu().numpy() - orc)
    assert d2.max() <= 0.03 * 1.001 and (d2 > 0).mean() < 1e-3


@pytest.mark.gpu
def test_add_layer_norm_rejects_bad_shapes():
    from outeffhop_b200 import ops
    a = torch.randn(4, 12, device="cuda", dtype=torch.float16)      # 12 is not a multiple of 8
    w = torch.ones(12, device="cuda", dtype=torch.float16)
    with pytest.raises(NotImplementedError):
        ops.add_layer_norm(a, None, w, w)
    with pytest.raises(ValueError):
        ops.add_layer_norm(a, a[:2], w, w)
    with pytest.raises(RuntimeError):
        ops.add_layer_norm(a.cpu(), None, w.cpu(), w.cpu())


@pytest.mark.gpu
def test_fused_bert_residual_modules_match_hf():
    from outeffhop_b200 import eval_harness as EH
    model = EH.build_bert_base(seed=1, num_hidden_layers=2, fuse_residual=False)
    b = EH.synthetic_mlm_batch(4, 128, seed=9)
    ids, lab = b["input_ids"].cuda(), b["labels"].cuda()
    with torch.no_grad():
        want = EH.bert_mlm_loss_aten(model, ids, lab).item()
        EH.fuse_bert_residual_blocks(model)
        got = EH.bert_mlm_loss(model, ids, lab).item()
    assert abs(got - want) < 2e-3 * abs(want)
    sd = model.state_dict()
    assert "bert.encoder.layer.0.attention.output.LayerNorm.weight" in sd and "bert.encoder.layer.1.output.dense.bias" in sd


# ---- taps folded into consumer kernels (fixed-range INT8 step) --------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_add_layer_norm_with_producer_tap_equals_unfused_chain(dtype):
    """qparams_a: y == add_layer_norm(fake_quant(a), b, ...) bit for bit (same qdq arithmetic in both kernels)."""
    from outeffhop_b200 import ops
    t = _T[dtype]
    g = torch.Generator(device="cuda").manual_seed(11)
    a = (torch.randn(300, 768, device="cuda", generator=g) * 2).to(t)
    b = torch.randn(300, 768, device="cuda", generator=g).to(t)
    w = (torch.rand(768, device="cuda", generator=g) + 0.5).to(t)
    bias = (torch.randn(768, device="cuda", generator=g) * 0.1).to(t)
    qa, qr, qo = (torch.tensor(v, device="cuda") for v in ([0.04, 120.0], [0.05, 131.0], [0.03, 125.0]))
    fused = ops.add_layer_norm(a, b, w, bias, 1e-12, qparams_a=qa, qparams_res=qr, qparams_out=qo)
    chain = ops.add_layer_norm(ops.fake_quant_asym(a, qa), b, w, bias, 1e-12, qparams_res=qr, qparams_out=qo)
    assert torch.equal(fused, chain)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_add_layer_norm_common_chain_body_equals_general_body(dtype):
    """16-bit launches with taps branch once to a body that holds only the quantizer's common chain (QdqVec::fast(): 8-bit codes,
    normal 1 / scale, zero point in range); anything else runs the general body.  Both must give the bits of the standalone
    chain fake_quant(a) -> + b -> fake_quant -> LayerNorm -> fake_quant:
      * all taps common                      -> fast body
      * ONE tap off the common chain (a 12-bit grid: codes not exact in the tensor dtype) -> the general body for all three taps,
        whose 8-bit taps must still equal the fast body's arithmetic (checked through the unfused chain)."""
    from outeffhop_b200 import ops
    t = _T[dtype]
    g = torch.Generator(device="cuda").manual_seed(23)
    a = (torch.randn(257, 768, device="cuda", generator=g) * 2).to(t)
    b = torch.randn(257, 768, device="cuda", generator=g).to(t)
    w = (torch.rand(768, device="cuda", generator=g) + 0.5).to(t)
    bias = (torch.randn(768, device="cuda", generator=g) * 0.1).to(t)
    qa, qr, qo = (torch.tensor(v, device="cuda") for v in ([0.04, 120.0], [0.05, 131.0], [0.03, 125.0]))

    def unfused(n_bits, qr_):
        s = ops.fake_quant_asym(a, qa, n_bits) + b
        s = ops.fake_quant_asym(s, qr_, n_bits)
        y = torch.nn.functional.layer_norm(s.float(), (768,), w.float(), bias.float(), 1e-12).to(t)
        return ops.fake_quant_asym(y, qo, n_bits), s

    y8, s8 = ops.add_layer_norm(a, b, w, bias, 1e-12, qparams_a=qa, qparams_res=qr, qparams_out=qo, return_sum=True)
    want8, wsum8 = unfused(8, qr)
    assert torch.equal(s8, wsum8)                                  # the taps and the sum: bit for bit
    assert float((y8.float() - want8.float()).abs().max()) <= 2 * 0.03 + 1e-6   # LayerNorm itself: fp32 two-pass vs ATen, <= 2 codes
    # same taps through the general body: zero point outside [0, 255] on the residual tap disables fast() for the launch
    qr_out = torch.tensor([0.05, 300.0], device="cuda")
    yg, sg = ops.add_layer_norm(a, b, w, bias, 1e-12, qparams_a=qa, qparams_res=qr_out, qparams_out=qo, return_sum=True)
    _, wsumg = unfused(8, qr_out)
    assert torch.equal(sg, wsumg)
    # and a launch whose ONLY tap is common must equal the same tap inside a general-body launch, element for element
    y_fast = ops.add_layer_norm(a, b, w, bias, 1e-12, qparams_out=qo)
    y_gen = ops.add_layer_norm(a, b, w, bias, 1e-12, qparams_out=qo, qparams_a=torch.tensor([1e-3, 300.0], device="cuda"))
    a_tapped = ops.fake_quant_asym(a, torch.tensor([1e-3, 300.0], device="cuda"))
    assert torch.equal(y_gen, ops.add_layer_norm(a_tapped, b, w, bias, 1e-12, qparams_out=qo))
    assert y_fast.shape == y_gen.shape


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("n", [1, 4099, 1 << 20])
def test_tap_add_fake_quant_equals_unfused_chain(dtype, n):
    from outeffhop_b200 import ops
    t = _T[dtype]
    g = torch.Generator(device="cuda").manual_seed(n)
    a = (torch.randn(n, device="cuda", generator=g) * 2).to(t)
    b = (torch.randn(n, device="cuda", generator=g) * 3).to(t)
    qb, qs = torch.tensor([0.05, 128.0], device="cuda"), torch.tensor([0.07, 119.0], device="cuda")
    fused = ops.add_fake_quant(a, b, qs, qparams_b=qb)
    chain = ops.add_fake_quant(a, ops.fake_quant_asym(b, qb), qs)
    assert torch.equal(fused, chain)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16"])
@pytest.mark.parametrize("n_seg", [1, 3, 4])
def test_fake_quant_segments_equals_per_segment_fake_quant(dtype, n_seg):
    from outeffhop_b200 import ops
    t = _T[dtype]
    g = torch.Generator(device="cuda").manual_seed(5 + n_seg)
    x = (torch.randn(2, 37, n_seg * 768, device="cuda", generator=g) * 3).to(t)
    qps = [torch.tensor([0.02 * (s + 1), 100.0 + 10 * s], device="cuda") for s in range(n_seg)]
    got = ops.fake_quant_asym_segments(x, qps)
    for s in range(n_seg):
        want = ops.fake_quant_asym(x[..., s * 768:(s + 1) * 768].contiguous(), qps[s])
        assert torch.equal(got[..., s * 768:(s + 1) * 768], want)
    with pytest.raises(ValueError):
        ops.fake_quant_asym_segments(x[..., :100], qps[:3] if n_seg >= 3 else qps * 3)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
def test_add_layer_norm_returns_the_sum_for_pre_ln_blocks(dtype):
    from outeffhop_b200 import ops
    t = _T[dtype]
    g = torch.Generator(device="cuda").manual_seed(3)
    a = (torch.randn(2, 50, 768, device="cuda", generator=g) * 2).to(t)
    b = torch.randn(2, 50, 768, device="cuda", generator=g).to(t)
    w, bias = torch.ones(768, device="cuda", dtype=t), torch.zeros(768, device="cuda", dtype=t)
    y, s = ops.add_layer_norm(a, b, w, bias, 1e-5, return_sum=True)
    assert torch.equal(s, a + b)                                            # ATen's add, bit for bit
    assert torch.equal(y, ops.add_layer_norm(a, b, w, bias, 1e-5))


@pytest.mark.gpu
def test_fused_opt_decoder_layer_matches_hf_layer():
    from outeffhop_b200 import eval_harness as EH
    model = EH.build_opt_125m(seed=3, num_hidden_layers=2, attn_softmax="softmax1", fuse_residual=False)
    b = EH.synthetic_clm_batch(2, 512, seed=1)
    ids, lab = b["input_ids"].cuda(), b["labels"].cuda()
    mask = EH.make_causal_mask(2, 512, torch.float16, "cuda")
    pos = torch.arange(512, device="cuda").unsqueeze(0).expand(2, 512).contiguous()
    with torch.no_grad():
        want = EH.opt_clm_loss(model, ids, lab, mask, pos, aten_loss=True).item()
        EH.fuse_opt_decoder_layers(model)
        got = EH.opt_clm_loss(model, ids, lab, mask, pos).item()
    assert abs(got - want) < 2e-3 * abs(want), (got, want)
    assert "model.decoder.layers.1.final_layer_norm.weight" in model.state_dict()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_add_layer_norm_emits_gate_logits(dtype):
    """oeh_add_layer_norm_gate: the LayerNorm kernel also produces the per-token gate logits of the attention module that
    consumes its output (Linear(64,1) predictors, bert_attention.py:303-320 / opt_attention.py:285-302) -- identical y, and
    logits equal to oeh_gate_logits on that y."""
    from outeffhop_b200 import ops
    torch.manual_seed(5)
    B, T, H = 3, 70, 12
    a = torch.randn(B, T, 64 * H, device="cuda").to(dtype)
    b = torch.randn(B, T, 64 * H, device="cuda").to(dtype)
    w = (1 + 0.1 * torch.randn(64 * H, device="cuda")).to(dtype)
    bias = (0.1 * torch.randn(64 * H, device="cuda")).to(dtype)
    gw = torch.randn(H, 64, device="cuda") * 0.2
    gb = torch.randn(H, device="cuda")
    y0, s0 = ops.add_layer_norm(a, b, w, bias, 1e-5, return_sum=True)
    y1, s1, logits = ops.add_layer_norm(a, b, w, bias, 1e-5, return_sum=True, gate=(gw, gb))
    assert torch.equal(y0, y1) and torch.equal(s0, s1) and logits.shape == (B, H, T)
    ref = ops.gate_logits(y0, H, gw, gb)
    assert torch.allclose(logits, ref, atol=2e-5, rtol=1e-5), float((logits - ref).abs().max())
    y2, logits2 = ops.add_layer_norm(a, None, w, bias, 1e-5, gate=(gw, gb))            # plain LayerNorm + gate
    assert torch.equal(y2, ops.add_layer_norm(a, None, w, bias, 1e-5))
    assert torch.allclose(logits2, ops.gate_logits(y2, H, gw, gb), atol=2e-5, rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_fused_bert_embeddings_and_head_transform_match_hf(dtype):
    """hf_compat.fuse_bert_shell: BertEmbeddings (3 gathers + 2 adds + LayerNorm) and BertPredictionHeadTransform
    (dense + GELU + LayerNorm) as single kernels vs the HF modules they replace."""
    from transformers import BertConfig, BertForMaskedLM
    from outeffhop_b200 import hf_compat
    cfg = BertConfig(num_hidden_layers=1)
    torch.manual_seed(0)
    ref = BertForMaskedLM(cfg).cuda().to(dtype).eval()
    with torch.no_grad():
        ref.bert.embeddings.LayerNorm.weight.add_(0.1 * torch.randn_like(ref.bert.embeddings.LayerNorm.weight))
    import copy
    fused = hf_compat.fuse_bert_shell(copy.deepcopy(ref))
    assert isinstance(fused.bert.embeddings, hf_compat.BertEmbeddingsFused)
    assert list(fused.state_dict()) == list(ref.state_dict())
    ids = torch.randint(0, cfg.vocab_size, (4, 128), device="cuda")
    tt = torch.randint(0, 2, (4, 128), device="cuda")
    with torch.no_grad():
        for kw in ({}, {"token_type_ids": tt}, {"position_ids": torch.arange(127, -1, -1, device="cuda")[None]}):
            a, b = fused.bert.embeddings(input_ids=ids, **kw), ref.bert.embeddings(input_ids=ids, **kw)
            tol = 1e-5 if dtype == torch.float32 else (2e-3 if dtype == torch.float16 else 2e-2)
            assert (a.float() - b.float()).abs().max().item() <= tol, kw      # same sums; LayerNorm rounding differs by <= 1 ulp
        if dtype != torch.float32:
            h = torch.randn(4, 128, cfg.hidden_size, device="cuda").to(dtype)
            a, b = fused.cls.predictions.transform(h), ref.cls.predictions.transform(h)
            assert (a.float() - b.float()).abs().max().item() <= (6e-3 if dtype == torch.float16 else 5e-2)
