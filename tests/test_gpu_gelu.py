"""erf GELU kernel (oeh_gelu_erf, csrc/gelu.cu) -- the activation HF's BertIntermediate applies between the FFN GEMMs.
Reference: float64 ``x * 0.5 * (1 + erf(x / sqrt 2))`` rounded once to the tensor dtype (what ATen's fp32-opmath kernel
approximates), and torch.nn.functional.gelu itself on the same device."""
import math

import numpy as np
import pytest
import torch

from outeffhop_b200 import hf_compat


def _ops():
    from outeffhop_b200 import ops
    return ops


def _exact(x64: torch.Tensor) -> torch.Tensor:
    return x64 * 0.5 * torch.special.erfc(-x64 / math.sqrt(2.0))      # = 0.5 (1 + erf), without the cancellation


def _ulp_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """distance in representable 16-bit values (sign-magnitude -> linear)"""
    def lin(t):
        i = t.view(torch.int16).to(torch.int32)
        return torch.where(i < 0, -(i & 0x7FFF), i)
    return (lin(a) - lin(b)).abs()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gelu_all_16bit_patterns(dtype):
    """Every finite 16-bit input: at most one ulp from the exactly rounded value, on a handful of inputs; NaN / inf as
    ATen (NaN -> NaN, +inf -> +inf)."""
    ops = _ops()
    bits = torch.arange(65536, dtype=torch.int32).to(torch.int16).cuda()
    x = bits.view(dtype)
    y = ops.gelu(x)
    fin = torch.isfinite(x)
    ref64 = _exact(x[fin].double())
    ref = ref64.to(dtype)
    d = _ulp_diff(y[fin], ref)
    # -0.0 vs +0.0 (x * Phi underflow) is not a difference; bf16 only: results below fp32's normal range (x < -13,
    # |gelu| < 1e-37) flush to zero in the fp32 arithmetic, here as in ATen
    d = torch.where(((y[fin] == 0) & (ref == 0)) | ((ref64.abs() < 1e-36) & (y[fin].float().abs() < 1e-36)),
                    torch.zeros_like(d), d)
    assert int(d.max()) <= 1
    assert int((d > 0).sum()) <= 16, int((d > 0).sum())
    assert bool(torch.isnan(y[torch.isnan(x)]).all())
    assert bool((y[x == float("inf")] == float("inf")).all())
    # ATen's own fp32 chain (1 + erf cancels in the negative tail) stays within a few 16-bit ulps of the same values
    t = torch.nn.functional.gelu(x[fin])
    assert float((y[fin].float() - t.float()).abs().max()) <= 2e-4 if dtype == torch.float16 else True
    near = _ulp_diff(y[fin], t)
    near = torch.where((y[fin] == 0) & (t == 0), torch.zeros_like(near), near)
    assert float((near > 0).float().mean()) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 8, 4099, 3 * 1024 * 1024 + 5])
def test_gelu_sizes_alignment_inplace(n):
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    base = (torch.randn(n + 3, generator=g) * 3).half().cuda()
    for off in (0, 1):                                   # 16-byte aligned and not
        x = base[off:off + n]
        ref = _exact(x.double()).half()
        y = ops.gelu(x)
        assert int(_ulp_diff(y, ref).max()) <= 1
        z = x.clone()                                    # in place; the clone is 16-byte aligned even when x is not:
        assert ops.gelu(z, out=z) is z                   # the vector and scalar paths give the same bits
        assert torch.equal(z, y)


@pytest.mark.gpu
def test_gelu_errors():
    ops = _ops()
    with pytest.raises(NotImplementedError):
        ops.gelu(torch.zeros(8, device="cuda", dtype=torch.float32))
    with pytest.raises(ValueError):
        ops.gelu(torch.zeros(8, device="cuda", dtype=torch.float16), out=torch.zeros(4, device="cuda", dtype=torch.float16))
    with pytest.raises((ValueError, RuntimeError)):
        ops.gelu(torch.zeros(8, dtype=torch.float16))   # CPU tensor: no fallback
    assert ops.gelu(torch.zeros(0, device="cuda", dtype=torch.float16)).numel() == 0


@pytest.mark.gpu
def test_gelu_module_in_bert_step():
    """build_bert_base swaps the FFN activation; logits agree with the ATen-GELU model within fp16 noise, and a forward
    hook on the module (somebody observing it) falls back to ATen."""
    from outeffhop_b200 import eval_harness as EH
    model = EH.build_bert_base(seed=5, num_hidden_layers=2, attn_softmax="softmax1")
    acts = [l.intermediate.intermediate_act_fn for l in model.bert.encoder.layer]
    assert all(isinstance(a, hf_compat.GELUFused) for a in acts)
    batch = {k: v.cuda() for k, v in EH.synthetic_mlm_batch(4, 128, seed=1).items()}
    with torch.no_grad():
        a = model(**batch).logits.float()
        handles = [m.register_forward_hook(lambda *_: None) for m in acts]
        b = model(**batch).logits.float()
        for h in handles:
            h.remove()
    assert float((a - b).abs().max()) <= 2e-2 * float(b.abs().max())
    assert float((a - b).abs().mean()) <= 1e-3 * float(b.abs().max())


def test_gelu_module_cpu_fallback_and_detection():
    """Host logic (no GPU): fp32 / CPU tensors take ATen; only the erf form is swapped."""
    m = hf_compat.GELUFused().eval()
    x = torch.linspace(-4, 4, 33)
    assert torch.equal(m(x), torch.nn.functional.gelu(x))
    assert hf_compat._is_erf_gelu(torch.nn.functional.gelu)
    assert hf_compat._is_erf_gelu(torch.nn.GELU())
    assert not hf_compat._is_erf_gelu(torch.nn.GELU(approximate="tanh"))
    assert not hf_compat._is_erf_gelu(torch.nn.ReLU())
    from transformers.activations import ACT2FN
    assert hf_compat._is_erf_gelu(ACT2FN["gelu"])
    assert not hf_compat._is_erf_gelu(ACT2FN["gelu_new"])
