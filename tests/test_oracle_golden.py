"""Pin oracle/ (numpy restatement) to the live reference via tests/golden/*.npz and the
SURVEY.md section-4 known-answer table.  CPU only."""
import numpy as np
import pytest

from oracle import oeh_oracle as O

F32 = np.float32


# ---------------------------------------------------------------- KAT table
def test_kat_softmax1():
    np.testing.assert_allclose(O.softmax_1(np.array([0, 0], F32)), [1 / 3, 1 / 3], rtol=1e-7)
    np.testing.assert_allclose(O.softmax_1(np.array([0], F32)), [0.5])
    np.testing.assert_allclose(O.softmax_1(np.array([1, 2, 3], F32)),
                               [0.087144323, 0.236882836, 0.643914282], rtol=2e-7)
    np.testing.assert_allclose(O.softmax_1(np.array([100, 100], F32)), [0.5, 0.5])
    assert np.array_equal(O.softmax_1(np.array([-100, -100], F32)), [0, 0])
    assert np.array_equal(O.softmax_1(np.array([-90, -90], F32)), [0, 0])
    fmin = np.finfo(F32).min
    np.testing.assert_allclose(O.softmax_1(np.array([0, fmin], F32)), [0.5, 0])
    assert np.array_equal(O.softmax_1(np.array([fmin, fmin], F32)), [0, 0])
    np.testing.assert_allclose(O.softmax_vanilla(np.array([fmin, fmin], F32)), [0.5, 0.5])


def test_kat_clipped():
    f = O.softmax_mapping("clipped(-.025:1)")
    np.testing.assert_allclose(f(np.array([0, 0], F32)), [0.4875, 0.4875], rtol=1e-6)
    np.testing.assert_allclose(f(np.array([10, 0, 0, 0], F32)), [0.999860287, 0, 0, 0], rtol=1e-6, atol=1e-9)
    g = O.softmax_mapping("clippedsoftmax1(-.025:1)")   # eta is 1.1 in the table (softmax.py:61)
    np.testing.assert_allclose(g(np.array([0, 0], F32)), [0.35, 0.35], rtol=1e-6)
    np.testing.assert_allclose(g(np.array([10, 0], F32)), [1.0, 0.0])


def test_kat_quantizer():
    delta, zf = O.asym_set_quant_range(-1.0, 3.0, 8)
    assert delta == F32(0.015686275) and abs(zf - F32(63.749996)) < 1e-5
    scale, zp = O.asym_scale_zero_point(delta, zf, 8)
    assert zp == 64
    x = np.array([0.5, -2, 5, delta / 2, 1.5 * delta, -1, 3], F32)
    assert O.asym_to_integer(x, delta, zf).tolist() == [96, 0, 255, 64, 66, 0, 255]   # half-to-even
    np.testing.assert_allclose(O.asym_quantize(x, delta, zf),
                               [0.501960814, -1.003921628, 2.996078491, 0, 0.031372551, -1.003921628, 2.996078491],
                               rtol=1e-7)


def test_kat_running_minmax():
    est = O.RunningMinMax(0.9)
    got = [est(np.array(p, F32)) for p in ([-1, 2], [-3, 4], [0.5, 1])]
    exp = [(-1, 2), (-1.200000048, 2.200000048), (-1.030000091, 2.079999924)]
    for (a, b), (ea, eb) in zip(got, exp):
        assert a == F32(ea) and b == F32(eb)


def test_logit():
    assert abs(O.logit(0.25) - (-1.0986122886681098)) < 1e-12


# ---------------------------------------------------------------- golden: softmax table
def test_softmax_table_matches_reference(golden):
    g = golden("softmax.npz")
    all_keys = [str(k) for k in g["all_keys"]]
    assert all_keys == list(O.SOFTMAX_TABLE.keys())          # same 40 keys, same order
    for key, kind, gamma, eta in zip(g["keys"], g["kinds"], g["gammas"], g["etas"]):
        assert O.SOFTMAX_TABLE[str(key)] == (int(kind), float(gamma), float(eta)), key


def test_softmax_outputs_match_reference(golden):
    g = golden("softmax.npz")
    x = g["x"]
    for key in g["keys"]:
        y = O.softmax_mapping(str(key))(x, dim=-1)
        ref = g["y::" + str(key)]
        assert y.dtype == ref.dtype
        np.testing.assert_allclose(y, ref, rtol=2e-6, atol=6e-8, err_msg=str(key))  # 1 ulp of a prob near 1
        assert np.array_equal(ref[np.abs(ref - y) == 0] == 0, y[np.abs(ref - y) == 0] == 0)
        assert (y[ref == 0] <= 6e-8).all() and (ref[y == 0] <= 6e-8).all(), key   # clip / overflow zeros agree


# ---------------------------------------------------------------- golden: quantizer (bit exact)
def _bits(a):
    return np.ascontiguousarray(a).view({2: np.uint16, 4: np.uint32}[a.dtype.itemsize])


def assert_bit_equal(a, b, msg=""):
    """bit-exact up to the sign of zero (ATen's clamp may return either zero)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, msg
    same = (_bits(a) == _bits(b)) | ((a == 0) & (b == 0))
    assert same.all(), f"{msg}: {np.count_nonzero(~same)} mismatches, first at {np.argwhere(~same)[:3].tolist()}"


def test_quantizer_codes_bit_exact(golden):
    g = golden("quantizer.npz")
    for tag, n_bits, lo, hi in g["qcases"]:
        n_bits = int(n_bits)
        delta, zf = O.asym_set_quant_range(float(lo), float(hi), n_bits)
        assert_bit_equal(delta, g[f"{tag}::delta"], tag + " delta")
        assert_bit_equal(zf, g[f"{tag}::zero_float"], tag + " zero_float")
        scale, zp = O.asym_scale_zero_point(delta, zf, n_bits)
        assert_bit_equal(scale, g[f"{tag}::scale"], tag + " scale")
        assert_bit_equal(zp, g[f"{tag}::zero_point"], tag + " zp")
        x = g[f"{tag}::x"]
        assert_bit_equal(O.asym_to_integer(x, delta, zf, n_bits), g[f"{tag}::x_int"], tag + " x_int")
        assert_bit_equal(O.asym_quantize(x, delta, zf, n_bits), g[f"{tag}::x_q"], tag + " x_q")
        if f"{tag}::x_h" in g.files:
            xh = g[f"{tag}::x_h"]
            assert_bit_equal(O.asym_to_integer(xh, delta, zf, n_bits, lowp_mode="cpu"), g[f"{tag}::x_int_h"], tag + " x_int fp16")
            assert_bit_equal(O.asym_quantize(xh, delta, zf, n_bits, lowp_mode="cpu"), g[f"{tag}::x_q_h"], tag + " x_q fp16")


def test_running_minmax_manager_bit_exact(golden):
    g = golden("quantizer.npz")
    assert_bit_equal(np.array([list(p) for p in _kat_pairs()], F32), g["rm_kat"])
    for tag in g["rm_cases"]:
        qm = O.QuantizationManager(8, 0.9)
        xs, ys = g[f"{tag}::x"], g[f"{tag}::y"]
        for i in range(xs.shape[0]):
            y = qm(xs[i])
            assert_bit_equal(qm.estimator.current_xmin, g[f"{tag}::cur_min"][i], f"{tag} min step {i}")
            assert_bit_equal(qm.estimator.current_xmax, g[f"{tag}::cur_max"][i], f"{tag} max step {i}")
            assert_bit_equal(qm.delta, g[f"{tag}::delta"][i], f"{tag} delta step {i}")
            assert_bit_equal(qm.zero_float, g[f"{tag}::zero_float"][i], f"{tag} zf step {i}")
            assert_bit_equal(y, ys[i], f"{tag} y step {i}")
        qm.fix_ranges()
        assert_bit_equal(qm(g[f"{tag}::x_fixed"]), g[f"{tag}::y_fixed"], f"{tag} fixed")


def _kat_pairs():
    est = O.RunningMinMax(0.9)
    return [est(np.array(p, F32)) for p in ([-1, 2], [-3, 4], [0.5, 1])]


def test_manager_requires_range_before_fix():
    with pytest.raises(RuntimeError):
        O.QuantizationManager().fix_ranges()


# ---------------------------------------------------------------- golden: attention modules
def _sd(g, tag, base):
    sd = {}
    for k in g.files:
        for pre in (base + "::sd::", tag + "::sd::"):
            if k.startswith(pre):
                sd[k[len(pre):]] = g[k].astype(F32)
    return sd


def _gate_wb(sd, heads):
    if "alpha.0.weight" not in sd:
        return None, None
    w = np.stack([sd[f"alpha.{h}.weight"][0] for h in range(heads)])
    b = np.array([sd[f"alpha.{h}.bias"][0] for h in range(heads)], F32)
    return w, b


def test_bert_module_matches_reference(golden):
    g = golden("attention_modules.npz")
    for tag, key, gated, masked in g["bert_cases"]:
        sd = _sd(g, tag, "bert_base")
        kind, gamma, eta = O.SOFTMAX_TABLE[str(key)]
        gw, gb = _gate_wb(sd, 2) if int(gated) else (None, None)
        mask = g[f"{tag}::mask"] if int(masked) else None
        y = O.bert_self_attention_forward(g[f"{tag}::hidden"].astype(F32), sd["query.weight"], sd["query.bias"],
                                          sd["key.weight"], sd["key.bias"], sd["value.weight"], sd["value.bias"], 2,
                                          kind=kind, gamma=gamma, eta=eta, attention_mask=mask, gate_w=gw, gate_b=gb)
        np.testing.assert_allclose(y, g[f"{tag}::out"], rtol=2e-4, atol=2e-5, err_msg=f"{tag} {key}")


def test_bert_alpha_override_matches_reference(golden):
    g = golden("attention_modules.npz")
    sd = _sd(g, "bert_alpha", "bert_base")
    y = O.bert_self_attention_forward(g["bert_alpha::hidden"].astype(F32), sd["query.weight"], sd["query.bias"],
                                      sd["key.weight"], sd["key.bias"], sd["value.weight"], sd["value.bias"], 2,
                                      kind=O.KIND_CLIPPED, gamma=-4.0 / 64, eta=1.0)   # bert_attention.py:88-94
    np.testing.assert_allclose(y, g["bert_alpha::out"], rtol=2e-4, atol=2e-5)


def test_bert_mlp_gate_matches_reference(golden):
    g = golden("attention_modules.npz")
    for tag in ("bert_mlp", "bert_mlp2"):
        sd = _sd(g, tag, "bert_base")
        hidden = g[f"{tag}::hidden"].astype(F32)
        w1 = np.stack([sd[f"alpha.{h}.0.weight"] for h in range(2)])
        b1 = np.stack([sd[f"alpha.{h}.0.bias"] for h in range(2)])
        w2 = np.stack([sd[f"alpha.{h}.2.weight"][0] for h in range(2)])
        b2 = np.array([sd[f"alpha.{h}.2.bias"][0] for h in range(2)], F32)
        xh = O._split_heads(hidden, 2)
        gate = O.sigmoid(O.gate_logits_mlp(xh, w1, b1, w2, b2))
        q = O._split_heads(O._linear(hidden, sd["query.weight"], sd["query.bias"]), 2)
        k = O._split_heads(O._linear(hidden, sd["key.weight"], sd["key.bias"]), 2)
        v = O._split_heads(O._linear(hidden, sd["value.weight"], sd["value.bias"]), 2)
        ctx = O.attention_core(q, k, v, scale=0.125, kind=O.KIND_SOFTMAX1, gate=gate)
        y = ctx.transpose(0, 2, 1, 3).reshape(hidden.shape)
        np.testing.assert_allclose(y, g[f"{tag}::out"], rtol=2e-4, atol=2e-5, err_msg=tag)


def test_opt_module_matches_reference(golden):
    g = golden("attention_modules.npz")
    for tag, key, gated in g["opt_cases"]:
        sd = _sd(g, tag, "opt_base")
        kind, gamma, eta = O.SOFTMAX_TABLE[str(key)]
        gw, gb = _gate_wb(sd, 2) if int(gated) else (None, None)
        y = O.opt_attention_forward(g[f"{tag}::hidden"].astype(F32), sd["q_proj.weight"], sd["q_proj.bias"],
                                    sd["k_proj.weight"], sd["k_proj.bias"], sd["v_proj.weight"], sd["v_proj.bias"],
                                    sd["out_proj.weight"], sd["out_proj.bias"], 2, kind=kind, gamma=gamma, eta=eta,
                                    causal=True, gate_w=gw, gate_b=gb)
        np.testing.assert_allclose(y, g[f"{tag}::out"], rtol=2e-4, atol=2e-5, err_msg=f"{tag} {key}")
