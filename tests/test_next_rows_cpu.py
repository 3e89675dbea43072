"""CPU pins for the SURVEY section 8f rows: the oracle restatements against the live-reference fixtures
(tests/golden/next_rows.npz, written by tests/golden/make_golden_next.py) and against numpy itself, plus the host
side of the percentile C ABI (index plan) which needs no GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import oeh_oracle as O

F32 = np.float32


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_percentile_oracle_equals_numpy(dtype):
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 100, 4097, 65537):
        x = (rng.standard_normal(n) * 3).astype(dtype)
        if n > 50:
            x[::37] *= 20
        for p in (99.999, 99.99, 99.0, 75.0, 50.0, 100.0):
            got = O.percentile_linear(x, 100 - p, p)
            want = np.percentile(x, (100 - p, p))
            assert got[0] == want[0] and got[1] == want[1], (dtype, n, p, got, want)


def test_percentile_oracle_nan_and_ties():
    x = np.ones(1000, F32)
    x[10] = np.nan
    assert all(np.isnan(v) for v in O.percentile_linear(x, 0.1, 99.9))
    x = np.zeros(1000, F32)
    x[:3] = -5
    x[-2:] = 7
    assert O.percentile_linear(x, 1, 99) == tuple(np.percentile(x, (1, 99)))


def test_percentile_plan_matches_numpy_indexing():
    from outeffhop_b200 import _lib
    h = _lib.lib()
    for n in (1, 2, 5, 1000, 12582912, (1 << 31) + 7):
        for q_lo, q_hi in ((100 - 99.999, 99.999), (0.0, 100.0), (25.0, 75.0), (100 - 99.9, 99.9)):
            r, g = (C.c_int64 * 4)(), (C.c_double * 2)()
            assert h.oeh_percentile_plan(n, q_lo, q_hi, r, g) == 0
            for j, q in enumerate((q_lo, q_hi)):
                virt = (n - 1) * (np.float64(q) / 100)
                prev = int(np.floor(virt))
                if virt >= n - 1:
                    assert (r[2 * j], r[2 * j + 1]) == (n - 1, n - 1) and g[j] == virt + 1
                else:
                    assert (r[2 * j], r[2 * j + 1]) == (prev, prev + 1) and g[j] == virt - prev
    r, g = (C.c_int64 * 4)(), (C.c_double * 2)()
    assert h.oeh_percentile_plan(0, 1.0, 99.0, r, g) == _lib.OEH_ERR_BAD_ARG
    assert h.oeh_percentile_plan(10, -1.0, 99.0, r, g) == _lib.OEH_ERR_BAD_ARG


def test_percentile_estimator_oracle_vs_reference(golden):
    g = golden("next_rows.npz")
    for ci in range(3):
        est = O.RunningMinMaxPercentile(float(g[f"pct{ci}::percentile"]))
        for step in range(3):
            x = g[f"pct{ci}::x"][step]
            lo, hi = est.update(x)
            d, zf, scale, zp = O.asym_set_quant_range_f64(lo, hi)
            want = g[f"pct{ci}::ranges"][step]
            assert (lo, hi, d, zf) == tuple(want), (ci, step)
            y = O.asym_quantize(x, F32(d), F32(zf))         # fp32(scale), zp: what ATen applies to an fp32 tensor
            # the oracle's fp32 helper recomputes scale/zp from fp32(delta), fp32(zero_float); rounding zero_float to
            # fp32 cannot move round() across .5 unless it sat within 1 ulp of it
            assert np.array_equal(y, g[f"pct{ci}::y"][step]), (ci, step)
        y = O.asym_quantize(g[f"pct{ci}::x_fixed"], F32(d), F32(zf))
        assert np.array_equal(y, g[f"pct{ci}::y_fixed"])
    lo, hi = O.percentile_linear(g["pct_cur::x"], 99.0, 1.0)          # CurrentMinMax order (:65-67)
    assert (lo, hi) == tuple(g["pct_cur::range"]) and lo > hi


def test_outlier_oracle_vs_reference(golden):
    g = golden("next_rows.npz")
    inf, kurt = O.outlier_stats(g["outlier::x"])
    assert np.array_equal(inf.astype(F32), g["outlier::inf"])
    np.testing.assert_allclose(kurt, g["outlier::kurt"], rtol=2e-5)


def test_residual_block_oracle_vs_reference(golden):
    """QuantizedBertSelfOutput: dense(QuantLinear) -> +input -> tap -> LayerNorm(QuantLayerNorm) re-stated with the
    oracle's quantizer pieces; the residual tap (the fused kernel's job) is checked bit-exactly through its recorded
    range, the block output with a tolerance (numpy vs ATen GEMM / LayerNorm summation order)."""
    g = golden("next_rows.npz")
    sd = {k[len("res::sd::"):]: g[k] for k in g.files if k.startswith("res::sd::")}
    d_res, zf_res = sd["res_act_quantizer.activation_quantizer.quantizer._delta"], \
        sd["res_act_quantizer.activation_quantizer.quantizer._zero_float"]
    w, b = g["res::fp::dense.weight"], g["res::fp::dense.bias"]
    # symmetric weight quantizer (uniform_quantizers.py:243-310) from the recorded delta
    dw = sd["dense.weight_quantizer.quantizer._delta"]
    signed = bool(sd["dense.weight_quantizer.quantizer._signed"])
    int_min, int_max = (-128.0, 127.0) if signed else (0.0, 255.0)
    wq = (np.clip(np.round(w / np.maximum(dw, F32(1e-8))), int_min, int_max) * np.maximum(dw, F32(1e-8))).astype(F32)
    h, inp = g["res::h_fixed"], g["res::in_fixed"]
    dense = (h.reshape(-1, 128) @ wq.T + b).astype(F32).reshape(h.shape)
    dense = O.asym_quantize(dense, sd["dense.activation_quantizer.quantizer._delta"],
                            sd["dense.activation_quantizer.quantizer._zero_float"])
    tapped = O.add_fake_quant(dense, inp, d_res, zf_res)
    lw, lb = g["res::fp::LayerNorm.weight"], g["res::fp::LayerNorm.bias"]
    dl = sd["LayerNorm.weight_quantizer.quantizer._delta"]
    lsigned = bool(sd["LayerNorm.weight_quantizer.quantizer._signed"])
    lmin, lmax = (-128.0, 127.0) if lsigned else (0.0, 255.0)
    lwq = (np.clip(np.round(lw / dl), lmin, lmax) * dl).astype(F32)
    mu = tapped.mean(-1, keepdims=True)
    var = tapped.var(-1, keepdims=True)
    ln = ((tapped - mu) / np.sqrt(var + 1e-12) * lwq + lb).astype(F32)
    y = O.asym_quantize(ln, sd["LayerNorm.activation_quantizer.quantizer._delta"],
                        sd["LayerNorm.activation_quantizer.quantizer._zero_float"])
    step = float(sd["LayerNorm.activation_quantizer.quantizer._delta"])
    diff = np.abs(y - g["res::y_fixed"])
    assert diff.max() <= step * 1.001 and (diff > 0).mean() < 0.02    # at most a few codes flip by one bin


def test_new_symbols_are_declared_in_header():
    import os
    from outeffhop_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "oeh.h")).read()
    for name in ("oeh_percentile", "oeh_percentile_plan", "oeh_percentile_hist", "oeh_percentile_scan",
                 "oeh_percentile_finish", "oeh_percentile_begin", "oeh_range_update_f64", "oeh_add_fake_quant",
                 "oeh_outlier_stats", "oeh_outlier_stats_workspace_bytes"):
        assert name + "(" in hdr and name in _lib.SYMBOLS
        getattr(_lib.lib(), name)
