"""GPU parity of the SURVEY section 8f rows through the C ABI: exact percentile ranges (vs numpy and vs the live
reference's QuantizationManager fixtures), the fused residual tap, outlier metrics, and the ViT / STanHop callers of
the fused attention kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oeh_oracle as O  # noqa: E402  (checker only)

F32 = np.float32


def _ops():
    from outeffhop_b200 import ops
    return ops


def _mods():
    import outeffhop_b200 as o
    return o


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------- percentile
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 4099, (1 << 20) + 5, 12582912])
def test_percentile_bit_exact_vs_numpy(dtype, n):
    ops = _ops()
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 3).astype(dtype)
    if n > 64:
        x[::97] *= 25
    xd = _dev(x)
    for p in (99.999, 99.9, 75.0, 50.0, 100.0):
        got = ops.percentile(xd, 100 - p, p).cpu().numpy()
        want = np.percentile(x, (100 - p, p))
        assert got.dtype == np.float64 and got[0] == want[0] and got[1] == want[1], (dtype, n, p, got, want)


def test_percentile_ties_zeros_nan_unaligned():
    ops = _ops()
    x = np.zeros(100003, F32)
    x[:7] = -3.5
    x[-5:] = 9.25
    x[50:60] = -0.0
    for p in (99.999, 99.99, 60.0):
        got = ops.percentile(_dev(x), 100 - p, p).cpu().numpy()
        assert np.array_equal(got, np.percentile(x, (100 - p, p))), p
    y = np.full(4097, 1.25, np.float16)                      # every element equal: all ranks in one bin every pass
    assert np.array_equal(ops.percentile(_dev(y), 0.1, 99.9).cpu().numpy(), [1.25, 1.25])
    rng = np.random.default_rng(5)
    z = rng.standard_normal(10007).astype(F32)
    base = _dev(np.concatenate([np.zeros(3, F32), z]))
    view = base[3:]                                          # 12-byte offset: exercises the ragged head
    assert np.array_equal(ops.percentile(view, 1.0, 99.0).cpu().numpy(), np.percentile(z, (1.0, 99.0)))
    z[77] = np.nan
    assert np.isnan(ops.percentile(_dev(z), 1.0, 99.0).cpu().numpy()).all()
    inf = z.copy()
    inf[77] = np.inf
    inf[78] = -np.inf
    with np.errstate(invalid="ignore"):
        want = np.percentile(inf, (0.0, 100.0))
    assert np.array_equal(ops.percentile(_dev(inf), 0.0, 100.0).cpu().numpy(), want, equal_nan=True)   # numpy: inf*0
    with np.errstate(invalid="ignore"):
        want = np.percentile(inf, (10.0, 90.0))
    assert np.array_equal(ops.percentile(_dev(inf), 10.0, 90.0).cpu().numpy(), want)
    with pytest.raises(TypeError):
        ops.percentile(torch.zeros(8, device="cuda", dtype=torch.bfloat16), 1, 99)
    with pytest.raises(IndexError):
        ops.percentile(torch.zeros(0, device="cuda"), 1, 99)
    with pytest.raises(ValueError):
        ops.percentile(torch.zeros(8, device="cuda"), -1, 99)


def test_percentile_stepwise_api_equals_one_shot():
    """The pass-by-pass entry points (the ones a multi-GPU caller all-reduces between) give the one-shot result;
    two shards whose histograms are summed give the percentile of the concatenation."""
    import ctypes as C
    from outeffhop_b200 import _lib as L
    ops = _ops()
    lib = L.lib()
    rng = np.random.default_rng(9)
    a, b = (rng.standard_normal(50000) * 2).astype(F32), (rng.standard_normal(50000) * 5 + 1).astype(F32)
    st = torch.cuda.current_stream().cuda_stream
    wsa = torch.zeros(lib.oeh_percentile_workspace_bytes(), dtype=torch.uint8, device="cuda")
    wsb = torch.zeros_like(wsa)
    xa, xb = _dev(a), _dev(b)
    ranks, gamma = (C.c_int64 * 4)(), (C.c_double * 2)()
    L.check(lib.oeh_percentile_plan(100000, 0.05, 99.95, ranks, gamma))
    for ws in (wsa, wsb):
        L.check(lib.oeh_percentile_begin(ws.data_ptr(), ws.numel(), ranks, st))
    ha, hb = (w[:ops.PCT_HIST_WORDS * 4].view(torch.int32) for w in (wsa, wsb))
    for p in range(lib.oeh_percentile_passes(L.OEH_F32)):
        L.check(lib.oeh_percentile_hist(xa.data_ptr(), 50000, L.OEH_F32, p, wsa.data_ptr(), st))
        L.check(lib.oeh_percentile_hist(xb.data_ptr(), 50000, L.OEH_F32, p, wsb.data_ptr(), st))
        tot = ha + hb                                          # what dist.all_reduce(SUM) leaves on every rank
        ha.copy_(tot)
        hb.copy_(tot)
        for ws in (wsa, wsb):
            L.check(lib.oeh_percentile_scan(L.OEH_F32, p, ws.data_ptr(), st))
    outs = []
    for ws in (wsa, wsb):
        o = torch.empty(2, dtype=torch.float64, device="cuda")
        L.check(lib.oeh_percentile_finish(L.OEH_F32, gamma, o.data_ptr(), ws.data_ptr(), st))
        outs.append(o.cpu().numpy())
    want = np.percentile(np.concatenate([a, b]), (0.05, 99.95))
    assert np.array_equal(outs[0], want) and np.array_equal(outs[1], want)


def test_percentile_manager_vs_reference_golden(golden):
    """QuantizationManager(asymmetric_uniform, running_minmax, percentile=p): float64 ranges and fp32 outputs
    bit-exact with the live reference over 3 estimation batches and a fixed-range batch."""
    o = _mods()
    g = golden("next_rows.npz")
    for ci in range(3):
        pct = float(g[f"pct{ci}::percentile"])
        mgr = o.QuantizationManager(qmethod=o.AsymmetricUniformQuantizer, init=o.RunningMinMaxEstimator,
                                    qparams=dict(n_bits=8, act_quant=True), init_params=dict(percentile=pct))
        for step in range(3):
            y = mgr(_dev(g[f"pct{ci}::x"][step]))
            est = mgr.range_estimator
            assert est.current_xmin.dtype == torch.float64 and mgr.quantizer._delta.dtype == torch.float64
            got = [float(est.current_xmin), float(est.current_xmax), float(mgr.quantizer._delta),
                   float(mgr.quantizer._zero_float)]
            assert got == list(g[f"pct{ci}::ranges"][step]), (ci, step, got)
            assert np.array_equal(y.cpu().numpy(), g[f"pct{ci}::y"][step]), (ci, step)
        mgr.fix_ranges()
        assert np.array_equal(mgr(_dev(g[f"pct{ci}::x_fixed"])).cpu().numpy(), g[f"pct{ci}::y_fixed"])
    est = o.CurrentMinMaxEstimator(percentile=99.0)
    lo, hi = est(_dev(g["pct_cur::x"]))
    assert [float(lo), float(hi)] == list(g["pct_cur::range"])


# ------------------------------------------------------------------------------------------- residual tap
@pytest.mark.parametrize("n", [1, 7, 4099, (1 << 20) + 3])
def test_add_fake_quant_fp32_bit_exact(n):
    ops = _ops()
    rng = np.random.default_rng(n)
    a, b = (rng.standard_normal(n) * 2).astype(F32), (rng.standard_normal(n) * 2).astype(F32)
    a[::53] *= 40
    delta, zf = O.asym_set_quant_range(-6.0, 9.0)
    scale, zp = O.asym_scale_zero_point(delta, zf)
    qp = torch.tensor([float(scale), float(zp)], device="cuda")
    got = ops.add_fake_quant(_dev(a), _dev(b), qp).cpu().numpy()
    assert np.array_equal(got, O.add_fake_quant(a, b, delta, zf))
    base = _dev(np.concatenate([np.zeros(1, F32), a]))          # unaligned operand
    assert np.array_equal(ops.add_fake_quant(base[1:], _dev(b), qp).cpu().numpy(), got)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_add_fake_quant_lowp_equals_unfused_kernels(dtype):
    ops = _ops()
    torch.manual_seed(3)
    a, b = (torch.randn(100003, device="cuda") * 3).to(dtype), torch.randn(100003, device="cuda").to(dtype)
    qp = torch.tensor([0.047, 121.0], device="cuda")
    assert torch.equal(ops.add_fake_quant(a, b, qp), ops.fake_quant_asym(a + b, qp))


def test_quantized_bert_self_output_vs_reference_golden(golden):
    o = _mods()
    from transformers import BertConfig
    from transformers.models.bert.modeling_bert import BertSelfOutput
    g = golden("next_rows.npz")
    fp = BertSelfOutput(BertConfig(hidden_size=128, num_attention_heads=2, hidden_dropout_prob=0.0)).eval()
    fp.load_state_dict({k[len("res::fp::"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("res::fp::")})
    qparams = dict(method=o.SymmetricUniformQuantizer, n_bits=8, n_bits_act=8, act_method=o.AsymmetricUniformQuantizer,
                   per_channel_weights=False, percentile=None, quant_setup="all",
                   weight_range_method=o.CurrentMinMaxEstimator, weight_range_options={},
                   act_range_method=o.RunningMinMaxEstimator, act_range_options={})
    qm = o.QuantizedBertSelfOutput(fp.cuda(), quant_dict={}, **qparams).eval()
    qm.set_quant_state(weight_quant=True, act_quant=True)
    step = float(g["res::sd::LayerNorm.activation_quantizer.quantizer._delta"])
    with torch.no_grad():
        for s in range(3):
            y = qm(_dev(g["res::h_est"][s]), _dev(g["res::in_est"][s]))
            d = np.abs(y.cpu().numpy() - g["res::y_est"][s])
            # estimated ranges inherit the GEMM's summation order, so scales differ in the last bits: count bin flips only
            assert d.max() <= 2.5 * step and (d > 0.01 * step).mean() < 0.03, (s, d.max(), (d > 0.01 * step).mean())
        qm.fix_ranges()
        assert qm.res_act_quantizer.fixed                                     # -> the fused add + qdq kernel
        y = qm(_dev(g["res::h_fixed"]), _dev(g["res::in_fixed"]))
    sd = qm.state_dict()
    for k in g.files:
        if k.startswith("res::sd::") and "_signed" not in k:
            np.testing.assert_allclose(sd[k[len("res::sd::"):]].float().cpu().numpy(), g[k], rtol=2e-6, err_msg=k)
    d = np.abs(y.cpu().numpy() - g["res::y_fixed"])
    assert d.max() <= 1.001 * step and (d > 0.01 * step).mean() < 0.02, (d.max(), (d > 0.01 * step).mean())


# ------------------------------------------------------------------------------------------- outlier metrics
def test_outlier_stats_vs_reference_golden_and_oracle(golden):
    ops, o = _ops(), _mods()
    g = golden("next_rows.npz")
    x = _dev(g["outlier::x"])
    inf, kurt = ops.outlier_stats(x)
    assert np.array_equal(inf.cpu().numpy(), g["outlier::inf"])
    np.testing.assert_allclose(kurt.cpu().numpy(), g["outlier::kurt"], rtol=2e-5)
    np.testing.assert_allclose(o.kurtosis(x.view(x.size(0), -1)).cpu().numpy(), g["outlier::kurt"], rtol=2e-5)
    torch.manual_seed(1)
    for dtype in (torch.float32, torch.float16, torch.bfloat16):
        for shape in ((32, 128, 768), (3, 1001), (1, 5), (5, 1 << 20)):
            t = (torch.randn(*shape, device="cuda") * 1.7 + 0.3).to(dtype)
            t.view(-1)[::1013] *= 30
            meter = torch.zeros(3, dtype=torch.float64, device="cuda")
            i1, k1 = ops.outlier_stats(t, meter=meter)
            i2, k2 = ops.outlier_stats(t, meter=meter)
            assert torch.equal(i1, i2) and torch.equal(k1, k2)                # deterministic
            ri, rk = O.outlier_stats(t.float().cpu().numpy())
            assert np.array_equal(i1.cpu().numpy(), ri.astype(F32)), (dtype, shape)
            np.testing.assert_allclose(k1.cpu().numpy(), rk, rtol=1e-5, err_msg=f"{dtype} {shape}")
            m = meter.cpu().numpy()
            np.testing.assert_allclose(m, [2 * i1.double().sum().item(), 2 * k1.double().sum().item(), 2 * shape[0]],
                                       rtol=1e-12)
    t = torch.randn(4, 100, device="cuda")
    t[2, 5] = float("nan")
    i, k = ops.outlier_stats(t)
    assert torch.isnan(i[2]) and torch.isnan(k[2]) and not torch.isnan(i[[0, 1, 3]]).any()


def test_outlier_meter_on_bert_hooks():
    o = _mods()
    from outeffhop_b200.eval_harness import build_bert_base, synthetic_mlm_batch
    model = build_bert_base(seed=1, dtype=torch.float16, attn_softmax="softmax1", num_hidden_layers=2)
    meter = o.OutlierMeter()
    handles = o.attach_bert_outlier_hooks(model, meter)
    batch = {k: v.cuda() for k, v in synthetic_mlm_batch(4, 32, seed=2).items()}
    with torch.no_grad():
        model(**batch)
    s = meter.summary()
    for h in handles:
        h.remove()
    assert {"max_ffn_out_inf_norm", "max_LN_out_inf_norm", "max_LN_inp_inf_norm", "avg_kurtosis", "max_kurtosis"} <= set(s)
    assert "bert.encoder.layer.1.output.LayerNorm.input" in s and all(np.isfinite(v) for v in s.values())


# ------------------------------------------------------------------------------------------- ViT / STanHop callers
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 4e-3), (torch.float32, 3e-4)])
def test_vit_attention_vs_reference_golden(golden, dtype, tol):
    o = _mods()
    g = golden("next_rows.npz")
    for ci in range(3):
        key, gated = str(g[f"vit{ci}::key"]), int(g[f"vit{ci}::gated"])
        gt = o.AttentionGateType.conditional_per_token if gated else o.AttentionGateType.none
        m = o.ViTSelfAttentionWithExtras(dim=128, num_heads=2, qkv_bias=True, softmax_fn=o.SOFTMAX_MAPPING[key],
                                         attn_gate_type=gt, attn_gate_init=0.25 if gated else None)
        sd = {k[len(f"vit{ci}::sd::"):]: torch.from_numpy(g[k].astype(F32)) for k in g.files if k.startswith(f"vit{ci}::sd::")}
        m.load_state_dict(sd)
        m = m.cuda().to(dtype).eval()
        with torch.no_grad():
            y = m(torch.from_numpy(g[f"vit{ci}::x"].astype(F32)).cuda().to(dtype))
        np.testing.assert_allclose(y.float().cpu().numpy(), g[f"vit{ci}::y"], atol=tol, rtol=0, err_msg=f"vit{ci} {key}")


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 3e-3), (torch.float32, 2e-5)])
def test_hopfield_association_vs_reference_golden(golden, dtype, tol):
    o = _mods()
    g = golden("next_rows.npz")
    cases = [(f"hop{ci}", str(g[f"hop{ci}::mode"])) for ci in range(3)] + [("hop_clip1", "clip_softmax1")]
    assert int(g["hop_clip1::ctor_raises"]) == 1           # the reference cannot construct this mode at all
    for tag, mode in cases:
        m = o.Association(scale=None, attention_dropout=0.0, eta=1.1, gamma=-0.1, mode=mode).eval()
        q, k, v = (torch.from_numpy(g[f"{tag}::{n}"].astype(F32)).cuda().to(dtype) for n in "qkv")
        with torch.no_grad():
            y = m(q, k, v)
        assert y.shape == (2, 33, 2, 64) and y.is_contiguous()
        np.testing.assert_allclose(y.float().cpu().numpy(), g[f"{tag}::y"], atol=tol, rtol=0, err_msg=f"{tag} {mode}")
    with pytest.raises(NotImplementedError):
        o.Association(mode="entmax")
    hp = o.Hopfield(128, 2, mode="softmax1").cuda().half().eval()
    with torch.no_grad():
        out = hp(*(torch.randn(2, 20, 128, device="cuda", dtype=torch.float16) for _ in range(3)))
    assert out.shape == (2, 20, 128) and torch.isfinite(out).all()
