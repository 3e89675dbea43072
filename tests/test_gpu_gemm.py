"""GPU parity of the tcgen05 Linear kernel with fused epilogue (oeh_linear_fwd), through the C ABI.

Reference semantics: QuantLinear.forward = F.linear -> [activation] -> activation_quantizer (quantization/hijacker.py:78-134)
and HF BertIntermediate = dense -> erf GELU.  Checked against an fp32 evaluation of the same chain on the same 16-bit
operands (fp32 accumulate, the result rounded to the tensor dtype after the GEMM, after the activation and after the tap,
like the unfused kernels), the tap through the oracle's quantizer.  Tolerances: GEMM 1 ulp of the 16-bit output (accumulation
order) ; tapped outputs may differ by ONE quantisation step where the accumulator sits on a bin edge (same contract as the
fused attention taps), on a small fraction of the elements."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oeh_oracle as O  # noqa: E402  (checker only)

SHAPES = [(4096, 2304, 768), (4096, 768, 768), (4096, 3072, 768), (4096, 768, 3072), (2048, 768, 768), (1000, 776, 520),
          (5, 8, 8), (129, 264, 72), (300, 128, 64)]


def _ops():
    from outeffhop_b200 import ops
    return ops


def _mk(M, N, K, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(M, K, generator=g)).to(dtype).cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(dtype).cuda()
    b = (torch.randn(N, generator=g) * 0.5).to(dtype).cuda()
    return x, w, b


def _ref(x, w, b, dtype, act="none"):
    y = (x.double() @ w.double().t() + (b.double() if b is not None else 0)).float().to(dtype)   # what the GEMM stores
    if act == "relu":
        y = torch.relu(y)
    elif act == "gelu":
        y = torch.nn.functional.gelu(y.float()).to(dtype)
    return y


def _ulp(ref, dtype):
    bits = 10 if dtype == torch.float16 else 7
    e = torch.floor(torch.log2(ref.float().abs().clamp_min(2.0 ** -14)))
    return 2.0 ** (e - bits)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("tile_n", [0, 128, 192, 256, 2128, 2192, 2256])   # 2xxx: CTA pairs (cta_group::2)
@pytest.mark.parametrize("shape", SHAPES)
def test_linear_matches_fp64_reference(shape, tile_n, dtype):
    ops = _ops()
    M, N, K = shape
    x, w, b = _mk(M, N, K, dtype, seed=M + N + K)
    y = ops.linear_fwd(x, w, b, tile_n=tile_n)
    ref = _ref(x, w, b, dtype)
    err = (y.float() - ref.float()).abs()
    assert y.shape == (M, N) and y.dtype == dtype
    # one ulp of the 16-bit store + the fp32 accumulation error of a K-term dot product of O(1) terms (matters only where
    # the result itself is tiny)
    acc_tol = 4e-6 * K ** 0.5
    assert bool((err <= _ulp(ref, dtype) + acc_tol).all()), float((err - _ulp(ref, dtype)).max())
    assert float((err > 0).float().mean()) < 0.02                       # fp32 accumulation: almost always the same rounding
    y2 = ops.linear_fwd(x, w, None, tile_n=tile_n)                      # no bias
    ref2 = _ref(x, w, None, dtype)
    assert bool(((y2.float() - ref2.float()).abs() <= _ulp(ref2, dtype) + acc_tol).all())


@pytest.mark.parametrize("tile_n", [0, 192, 2192, 2256])
@pytest.mark.parametrize("act", ["relu", "gelu"])
@pytest.mark.parametrize("shape", [(4096, 3072, 768), (2048, 3072, 768), (777, 520, 264)])
def test_linear_fused_activation(shape, act, tile_n):
    ops = _ops()
    M, N, K = shape
    x, w, b = _mk(M, N, K, torch.float16, seed=7 + M)
    y = ops.linear_fwd(x, w, b, act=act, tile_n=tile_n)
    ref = _ref(x, w, b, torch.float16, act)
    lin = _ref(x, w, b, torch.float16)
    err = (y.float() - ref.float()).abs()
    # the GEMM output may differ by one ulp (accumulation order); the activation's slope is <= 1.13, then the result's own
    # rounding; + 3e-5 (the library erf GELU's difference from ATen's in the negative tail, tests/test_gpu_gelu.py)
    tol = 1.13 * _ulp(lin, torch.float16) + _ulp(ref, torch.float16) + 3e-5 + 4e-6 * K ** 0.5
    assert bool((err <= tol).all()), float((err - tol).max())
    assert float((err > 0).float().mean()) < 0.03


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("act", ["none", "relu", "gelu"])
@pytest.mark.parametrize("tile_n", [0, 2192])
@pytest.mark.parametrize("n_seg", [1, 3])
def test_linear_fused_tap_vs_oracle_quantizer(n_seg, act, dtype, tile_n):
    """Fixed-range output tap(s) in the epilogue vs oracle.asym_quantize of the untapped 16-bit result (the unfused chain)."""
    ops = _ops()
    M, K = 2048, 768
    N = 768 * n_seg
    x, w, b = _mk(M, N, K, dtype, seed=31 + n_seg)
    plain = ops.linear_fwd(x, w, b, act=act).float().cpu().numpy()
    ranges = [(-3.0, 4.0), (-1.5, 1.0), (-5.0, 5.5)][:n_seg]
    qs = [O.asym_set_quant_range(lo, hi) for lo, hi in ranges]
    dev = [torch.tensor([float(v) for v in O.asym_scale_zero_point(*q)], device="cuda") for q in qs]
    y = ops.linear_fwd(x, w, b, act=act, qparams=dev if n_seg > 1 else dev[0], tile_n=tile_n).float().cpu().numpy()
    for s, q in enumerate(qs):
        sl = slice(768 * s, 768 * (s + 1))
        ref = O.asym_quantize(plain[:, sl], *q)
        step = float(q[0])
        diff = np.abs(y[:, sl] - ref)
        tol16 = 2.0 ** (-10 if dtype == torch.float16 else -7) * np.abs(ref)           # the 16-bit store of scale * code
        assert (diff <= 1.001 * step + tol16).all(), (float(diff.max()), step)
        assert (diff > 0.5 * step).mean() < 5e-3, float((diff > 0.5 * step).mean())    # flips only at bin edges
        codes = np.unique(np.round(y[:, sl] / step))
        assert len(codes) <= 256


def test_linear_error_conventions_and_views():
    ops = _ops()
    from outeffhop_b200._lib import OehUnsupported
    x, w, b = _mk(64, 128, 64, torch.float16, seed=1)
    with pytest.raises(TypeError):
        ops.linear_fwd(x.float(), w, b)
    with pytest.raises(ValueError):
        ops.linear_fwd(x, w[:, :32], b)
    with pytest.raises(OehUnsupported):
        ops.linear_fwd(x, w[:100], b[:100])                              # N % 8 != 0
    x3 = torch.randn(4, 16, 64, device="cuda", dtype=torch.float16)     # leading dims are flattened like F.linear
    y3 = ops.linear_fwd(x3, w, b)
    assert y3.shape == (4, 16, 128)
    assert torch.allclose(y3.float(), torch.nn.functional.linear(x3, w, b).float(), atol=2e-3, rtol=2e-3)
    assert ops.linear_fwd(x[:0], w, b).shape == (0, 128)
