"""The C restatement (oracle/oeh_oracle.c) must agree with the numpy restatement and the goldens."""
import numpy as np

from oracle import c_oracle as CO
from oracle import oeh_oracle as O
from test_oracle_golden import assert_bit_equal

F32 = np.float32


def test_c_softmax_rows_vs_golden(golden):
    g = golden("softmax.npz")
    x = g["x"]
    for key, kind, gamma, eta in zip(g["keys"], g["kinds"], g["gammas"], g["etas"]):
        y = CO.softmax_rows(x, int(kind), float(gamma), float(eta))
        np.testing.assert_allclose(y, g["y::" + str(key)], rtol=2e-6, atol=6e-8, err_msg=str(key))


def test_c_quantizer_bit_exact_vs_golden(golden):
    g = golden("quantizer.npz")
    for tag, n_bits, lo, hi in g["qcases"]:
        n_bits = int(n_bits)
        d, zf, scale, zp = CO.set_quant_range(float(lo), float(hi), n_bits)
        assert_bit_equal(d, g[f"{tag}::delta"].reshape(()), tag)
        assert_bit_equal(zf, g[f"{tag}::zero_float"].reshape(()), tag)
        y, codes = CO.fake_quant(g[f"{tag}::x"], scale, zp, n_bits, want_codes=True)
        assert_bit_equal(codes, g[f"{tag}::x_int"], tag + " codes")
        assert_bit_equal(y, g[f"{tag}::x_q"], tag + " x_q")


def test_c_running_minmax_vs_golden(golden):
    g = golden("quantizer.npz")
    for tag in g["rm_cases"]:
        xs = g[f"{tag}::x"]
        cur = np.zeros(2, F32)
        for i in range(xs.shape[0]):
            mm = CO.minmax(xs[i])
            assert mm[0] == xs[i].min() and mm[1] == xs[i].max()
            cur = CO.running_minmax_update(cur, mm, 0.9, i > 0)
            assert_bit_equal(cur[0], g[f"{tag}::cur_min"][i], f"{tag} {i}")
            assert_bit_equal(cur[1], g[f"{tag}::cur_max"][i], f"{tag} {i}")


def test_c_attention_vs_numpy_oracle():
    rng = np.random.default_rng(7)
    B, H, Tq, Tk, d = 2, 3, 37, 53, 64
    q = rng.standard_normal((B, H, Tq, d), F32)
    k = rng.standard_normal((B, H, Tk, d), F32)
    v = rng.standard_normal((B, H, Tk, d), F32)
    gate = rng.uniform(0, 1, (B, H, Tq)).astype(F32)
    mask = np.zeros((B, 1, 1, Tk), F32)
    mask[1, ..., 40:] = np.finfo(F32).min
    for kind, gamma, eta in ((0, 0, 1), (1, 0, 1), (2, -0.025, 1.0), (3, -0.025, 1.1)):
        ref = O.attention_core(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=mask, gate=gate)
        got = CO.attention(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=mask, gate=gate)
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-6)
    qs = rng.standard_normal((B, H, Tk, d), F32)
    ref = O.attention_core(qs, k, v, scale=1.0, kind=1, causal=True, clamp_min=float(np.finfo(F32).min))
    got = CO.attention(qs, k, v, scale=1.0, kind=1, causal=True, clamp_min=True, out_bthd=True)
    np.testing.assert_allclose(got, ref.transpose(0, 2, 1, 3).reshape(B, Tk, H * d), rtol=1e-4, atol=2e-6)


def test_c_gate_logits_vs_numpy():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 17, 128), F32)
    w = rng.standard_normal((2, 64), F32)
    b = rng.standard_normal(2).astype(F32)
    ref = O.gate_logits_per_token(O._split_heads(x, 2), w, b)
    np.testing.assert_allclose(CO.gate_logits(x, w, b, 2), ref, rtol=1e-5, atol=1e-5)
