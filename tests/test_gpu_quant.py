"""GPU parity of the HBM-bound quantizer / range kernels, through the C ABI, vs the oracle and goldens.
fp32: bit-exact.  fp16/bf16: bit-exact vs the oracle's ATen-CUDA semantics and vs a torch op chain on device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oeh_oracle as O  # noqa: E402  (checker only)
from test_oracle_golden import assert_bit_equal  # noqa: E402

F32 = np.float32


def _ops():
    from outeffhop_b200 import ops
    return ops


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _qparams(scale, zp):
    return torch.tensor([float(scale), float(zp)], dtype=torch.float32, device="cuda")


def test_library_loads_and_device_supported():
    from outeffhop_b200 import _lib
    assert _lib.lib().oeh_version() == 100
    assert _lib.lib().oeh_device_supported() == 1, "tests must run on an sm_100 device"


def test_fake_quant_golden_bit_exact(golden):
    ops = _ops()
    g = golden("quantizer.npz")
    for tag, n_bits, lo, hi in g["qcases"]:
        n_bits = int(n_bits)
        d, zf, scale, zp = ops.set_quant_range_host(float(lo), float(hi), n_bits)
        assert_bit_equal(F32(d), g[f"{tag}::delta"].reshape(()), tag)
        assert_bit_equal(F32(zf), g[f"{tag}::zero_float"].reshape(()), tag)
        x = _dev(g[f"{tag}::x"])
        qp = _qparams(scale, zp)
        assert_bit_equal(ops.fake_quant_asym(x, qp, n_bits).cpu().numpy(), g[f"{tag}::x_q"], tag + " x_q")
        assert_bit_equal(ops.fake_quant_asym(x, qp, n_bits, emit_codes=True).cpu().numpy(), g[f"{tag}::x_int"], tag + " codes")


@pytest.mark.parametrize("n", [1, 7, 4099, (1 << 20) + 5, 1 << 24])
def test_fake_quant_fp32_vs_oracle_sizes(n):
    ops = _ops()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(F32) * 3
    x[rng.integers(0, n, max(1, n // 10000))] *= 50      # outliers
    delta, zf = O.asym_set_quant_range(x.min() * 0.7, x.max() * 0.7)
    scale, zp = O.asym_scale_zero_point(delta, zf)
    k = np.arange(0, 256, dtype=F32)[: min(256, n)]
    x[: k.size] = (k + F32(0.5) - zp) * scale              # exact half-way points
    y = ops.fake_quant_asym(_dev(x), _qparams(scale, zp)).cpu().numpy()
    assert_bit_equal(y, O.asym_quantize(x, delta, zf), f"n={n}")


def test_fake_quant_unaligned_and_noncontiguous():
    ops = _ops()
    rng = np.random.default_rng(5)
    base = rng.standard_normal(10007).astype(F32)
    delta, zf = O.asym_set_quant_range(-2.0, 2.5)
    scale, zp = O.asym_scale_zero_point(delta, zf)
    xb = _dev(base)
    x = xb[1:]                                              # 4-byte aligned only
    out = torch.empty(10007, device="cuda")[1:]
    y = ops.fake_quant_asym(x, _qparams(scale, zp), out=out)
    assert_bit_equal(y.cpu().numpy(), O.asym_quantize(base[1:], delta, zf), "unaligned")
    x2 = xb[: 10000].view(100, 100).t()                     # non-contiguous view
    assert_bit_equal(ops.fake_quant_asym(x2, _qparams(scale, zp)).cpu().numpy(),
                     O.asym_quantize(np.ascontiguousarray(base[:10000].reshape(100, 100).T), delta, zf), "transposed")


def test_fake_quant_nan_inf_and_empty():
    ops = _ops()
    qp = _qparams(0.1, 12.0)
    x = torch.tensor([float("nan"), float("inf"), float("-inf"), 0.0], device="cuda")
    y = ops.fake_quant_asym(x, qp).cpu().numpy()
    ref = torch.tensor(0.1) * (torch.clamp(torch.round(x.cpu() / 0.1) + 12.0, 0, 255) - 12.0)
    assert np.isnan(y[0]) and np.array_equal(y[1:], ref.numpy()[1:])
    assert ops.fake_quant_asym(torch.empty(0, device="cuda"), qp).numel() == 0


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fake_quant_low_precision_matches_aten_cuda(dtype):
    """fp16/bf16 contract: the same per-op rounding ATen's CUDA kernels apply to the reference's chain
    (uniform_quantizers.py:114-115,146) with 0-dim fp32 CUDA buffers for scale / zero_point."""
    ops = _ops()
    torch.manual_seed(3)
    x = (torch.randn(1 << 18, device="cuda") * 2).to(dtype)
    for lo, hi in ((-1.0, 3.0), (0.0, 1.0), (-7.3, 0.2)):
        d, zf, scale, zp = ops.set_quant_range_host(lo, hi, 8)
        scale_t = torch.tensor(scale, device="cuda")        # 0-dim fp32 CUDA tensors, like the module buffers
        zp_t = torch.tensor(zp, device="cuda")
        x_int = torch.clamp(torch.round(x / scale_t) + zp_t, 0.0, 255.0)
        ref = scale_t * (x_int - zp_t)
        assert ref.dtype == dtype
        y = ops.fake_quant_asym(x, _qparams(scale, zp))
        assert torch.equal(y, ref), f"{dtype} range ({lo},{hi}): {(y != ref).sum().item()} mismatches"
        if dtype == torch.float16:
            yo = O.asym_quantize(x.cpu().numpy(), F32(d), F32(zf), lowp_mode="cuda")
            assert_bit_equal(y.cpu().numpy(), yo, "oracle cuda-mode")


@pytest.mark.parametrize("n", [1, 33, 100003, 1 << 22])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_minmax_exact(n, dtype):
    ops = _ops()
    torch.manual_seed(n)
    x = (torch.randn(n, device="cuda") * 4).to(dtype)
    if n > 10:
        x[n // 3] = 77.0
        x[n // 2] = -91.0
    mm = ops.minmax(x).cpu().numpy()
    assert mm[0] == x.float().min().item() and mm[1] == x.float().max().item()
    if n > 1:   # unaligned start
        mm2 = ops.minmax(x[1:]).cpu().numpy()
        assert mm2[0] == x[1:].float().min().item() and mm2[1] == x[1:].float().max().item()


def test_minmax_nan_propagates_and_empty_raises():
    ops = _ops()
    x = torch.randn(5000, device="cuda")
    x[1234] = float("nan")
    assert np.isnan(ops.minmax(x).cpu().numpy()).all()      # torch.min / torch.max propagate NaN
    with pytest.raises(RuntimeError):
        ops.minmax(torch.empty(0, device="cuda"))


def test_running_minmax_and_manager_golden_bit_exact(golden):
    """minmax -> EMA -> set_quant_range -> qdq on the device equals the reference's
    QuantizationManager in estimate_ranges state, step by step, then in fix_ranges state."""
    ops = _ops()
    from outeffhop_b200 import _lib as L
    g = golden("quantizer.npz")
    for tag in g["rm_cases"]:
        state = ops.new_qstate("cuda")
        xs, ys = g[f"{tag}::x"], g[f"{tag}::y"]
        for i in range(xs.shape[0]):
            y = ops.estimate_and_quantize(_dev(xs[i]), state)
            st = state.cpu().numpy()
            assert_bit_equal(st[L.QS_CUR_MIN], g[f"{tag}::cur_min"][i], f"{tag} min {i}")
            assert_bit_equal(st[L.QS_CUR_MAX], g[f"{tag}::cur_max"][i], f"{tag} max {i}")
            assert_bit_equal(st[L.QS_DELTA], g[f"{tag}::delta"][i], f"{tag} delta {i}")
            assert_bit_equal(st[L.QS_ZERO_FLOAT], g[f"{tag}::zero_float"][i], f"{tag} zf {i}")
            assert_bit_equal(y.cpu().numpy(), ys[i], f"{tag} y {i}")
        yf = ops.fake_quant_asym(_dev(g[f"{tag}::x_fixed"]), state[L.QS_SCALE:L.QS_SCALE + 2])
        assert_bit_equal(yf.cpu().numpy(), g[f"{tag}::y_fixed"], f"{tag} fixed")


def test_full_size_properties_fp32():
    """BASELINE-size tensor (2^28 fp32 = 1 GiB): exact equality with the reference's op chain evaluated by
    ATen on the device, idempotence, and code range -- size-independent properties."""
    ops = _ops()
    n = 1 << 28
    torch.manual_seed(11)
    x = torch.randn(n, device="cuda")
    x[::100003] *= 50
    state = ops.new_qstate("cuda")
    y = ops.estimate_and_quantize(x, state)
    st = state.cpu()
    assert st[0].item() == x.min().item() and st[1].item() == x.max().item()
    scale, zp = st[4], st[5]
    ref = scale.cuda() * (torch.clamp(torch.round(x / scale.cuda()) + zp.cuda(), 0, 255) - zp.cuda())
    assert torch.equal(y, ref)
    del ref
    assert torch.equal(ops.fake_quant_asym(y, state[4:6]), y)                     # idempotent
    codes = ops.fake_quant_asym(x, state[4:6], emit_codes=True)
    assert codes.min().item() >= 0 and codes.max().item() <= 255 and torch.equal(codes, codes.round())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n_bits", [4, 8, 11, 12])
def test_fake_quant_low_precision_exhaustive_vs_aten_cuda(dtype, n_bits):
    """EVERY 16-bit pattern (all exponents, denormals, inf, NaN) through the packed half2 / bfloat162 path and the scalar
    tail, for many ranges, against ATen's per-op-rounded chain on the GPU: bit for bit."""
    if n_bits > (8 if dtype == torch.bfloat16 else 11):
        pytest.skip("2^n - 1 is not representable in the tensor dtype: ATen rounds the clamp bound itself; the reference "
                    "quantises 16-bit activations to 8 bits (16-bit quantizers only on fp32 LayerNorm outputs)")
    ops = _ops()
    bits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16)
    x = torch.cat([bits, bits[:37]]).cuda().view(dtype)            # 65573 elements: vector body + unaligned scalar tail
    int_max = float(2 ** n_bits - 1)
    g = torch.Generator().manual_seed(n_bits)
    ranges = [(-1.0, 3.0), (0.0, 1.0), (-7.3, 0.2), (-0.01, 0.02), (-300.0, 500.0), (0.0, 1e-3), (-60000.0, 60000.0)]
    ranges += [(-float(a), float(b)) for a, b in (torch.rand(12, 2, generator=g) * 20.0)]
    for lo, hi in ranges:
        d, zf, scale, zp = ops.set_quant_range_host(lo, hi, n_bits)
        scale_t, zp_t = torch.tensor(scale, device="cuda"), torch.tensor(zp, device="cuda")
        x_int = torch.clamp(torch.round(x / scale_t) + zp_t, 0.0, int_max)
        ref = scale_t * (x_int - zp_t)
        y = ops.fake_quant_asym(x, _qparams(scale, zp), n_bits=n_bits)
        same = (y.view(torch.int16) == ref.view(torch.int16)) | (torch.isnan(y) & torch.isnan(ref))
        assert bool(same.all()), (dtype, n_bits, lo, hi, int((~same).sum()), x[~same][:4], y[~same][:4], ref[~same][:4])
        codes = ops.fake_quant_asym(x, _qparams(scale, zp), n_bits=n_bits, emit_codes=True)
        same = (codes.view(torch.int16) == x_int.view(torch.int16)) | (torch.isnan(codes) & torch.isnan(x_int))
        assert bool(same.all()), ("codes", dtype, n_bits, lo, hi, int((~same).sum()))
        # the fused consumers use the same packed chain: residual add + tap, producer tap + add + tap
        b = x.flip(0).contiguous()
        s_ref = x + b
        s_int = torch.clamp(torch.round(s_ref / scale_t) + zp_t, 0.0, int_max)
        ref2 = scale_t * (s_int - zp_t)
        y2 = ops.add_fake_quant(x, b, _qparams(scale, zp), n_bits=n_bits)
        same = (y2.view(torch.int16) == ref2.view(torch.int16)) | (torch.isnan(y2) & torch.isnan(ref2))
        assert bool(same.all()), ("add", dtype, n_bits, lo, hi, int((~same).sum()))
