"""oeh_linear_add_layer_norm (csrc/gemm_ln_sm100.cu): Linear -> + residual -> LayerNorm on a thread-block cluster, against the
two-launch path it replaces (library GEMM -> oeh_add_layer_norm: same element arithmetic, different fp32 summation orders) and
against an fp64 restatement of HF BertSelfOutput / BertOutput.forward (transformers modeling_bert; quantized_bert.py:541-606)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from outeffhop_b200 import ops
    return ops


def _mk(M, N, K, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g)
    gamma = 1.0 + 0.2 * torch.randn(N, generator=g)
    beta = 0.1 * torch.randn(N, generator=g)
    return [t.to(dtype).cuda() for t in (x, w, b, res, gamma, beta)]


def _ulp(ref, dtype):
    bits = 10 if dtype == torch.float16 else 7
    return np.maximum(np.abs(ref), 2.0 ** -14) * 2.0 ** -bits


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(4096, 768, 768), (2048, 768, 3072), (300, 768, 200), (128, 1024, 512), (77, 2048, 264),
                                   (129, 384, 64), (256, 1536, 768)])
def test_linear_add_layer_norm_vs_two_launches_and_fp64(dtype, M, N, K):
    ops = _ops()
    x, w, b, res, gamma, beta = _mk(M, N, K, dtype, seed=M + N + K)
    eps = 1e-12
    y = ops.linear_add_layer_norm(x, w, b, res, gamma, beta, eps)
    lin = torch.nn.functional.linear(x, w, b)
    two = ops.add_layer_norm(lin, res, gamma, beta, eps)
    # fp64 restatement with the roundings of the chain: Linear output and the sum are stored in the tensor dtype
    v = (x.double() @ w.double().t() + b.double()).to(dtype).double()
    s = (v + res.double()).to(dtype).double()
    mu = s.mean(-1, keepdim=True)
    var = ((s - mu) ** 2).mean(-1, keepdim=True)
    ref = ((s - mu) / torch.sqrt(var + eps) * gamma.double() + beta.double()).cpu().numpy()
    got, two_np = y.double().cpu().numpy(), two.double().cpu().numpy()
    # the Linear output may round the other way at fp32 accumulation noise, and take the rounding of the sum with it: one ulp of
    # max(|v|, |s|) (two where the sum crosses into the next binade), amplified by rstd * gamma
    amp = (gamma.double().abs().cpu().numpy()[None, :] / torch.sqrt(var + eps).cpu().numpy())
    tol = 1.01 * _ulp(ref, dtype) + amp * np.maximum(_ulp(v.cpu().numpy(), dtype), _ulp(s.cpu().numpy(), dtype)) * 2.02 + 1e-6
    assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) - tol))
    assert np.all(np.abs(two_np - ref) <= tol)
    # and the two implementations agree on almost every element exactly
    assert (got == two_np).mean() > 0.995


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("taps", ["res_out", "a_res_out", "a_only"])
def test_linear_add_layer_norm_taps_and_sum_out(dtype, taps):
    """INT8 joints: q_a (the Linear's tap), q_res (the residual tap), q_out (the LayerNorm tap) and sum_out (the OPT residual
    stream) -- compared with library GEMM -> oeh_add_layer_norm, which applies the same taps with the same arithmetic."""
    ops = _ops()
    M, N, K = 2048, 768, 768
    x, w, b, res, gamma, beta = _mk(M, N, K, dtype, seed=3)
    qa = torch.tensor([0.04, 131.0], device="cuda") if "a" in taps.split("_") else None
    qr = torch.tensor([0.05, 127.0], device="cuda") if "res" in taps else None
    qo = torch.tensor([0.03, 120.0], device="cuda") if "out" in taps else None
    y, s = ops.linear_add_layer_norm(x, w, b, res, gamma, beta, 1e-5, qparams_a=qa, qparams_res=qr, qparams_out=qo,
                                     return_sum=True)
    lin = torch.nn.functional.linear(x, w, b)
    y2, s2 = ops.add_layer_norm(lin, res, gamma, beta, 1e-5, qparams_a=qa, qparams_res=qr, qparams_out=qo, return_sum=True)
    # the sum: identical wherever the Linear output rounded identically (a tap may move a value by one bin at an edge)
    step = 0.05 if qr is not None else (0.04 if qa is not None else 0.0)
    ds = (s.float() - s2.float()).abs()
    assert (ds == 0).float().mean() > 0.97
    assert float(ds.max()) <= step + 2.0 ** (-8 if dtype == torch.float16 else -5)
    dy = (y.float() - y2.float()).abs()
    assert (dy == 0).float().mean() > 0.9
    bound = (0.03 if qo is not None else 0.0) + 0.08
    assert float(dy.max()) <= bound


def test_linear_add_layer_norm_without_residual_and_errors():
    ops = _ops()
    x, w, b, res, gamma, beta = _mk(256, 768, 768, torch.float16, seed=1)
    y = ops.linear_add_layer_norm(x, w, b, None, gamma, beta, 1e-5)
    ref = torch.nn.functional.layer_norm(torch.nn.functional.linear(x, w, b).float(), (768,), gamma.float(), beta.float(), 1e-5)
    assert float((y.float() - ref).abs().max()) < 2e-2
    assert not ops.linear_add_layer_norm_supported(x, torch.zeros(800, 768, device="cuda", dtype=torch.float16))
    with pytest.raises(NotImplementedError):
        ops.linear_add_layer_norm(x, torch.zeros(800, 768, device="cuda", dtype=torch.float16), None, None,
                                  torch.ones(800, device="cuda", dtype=torch.float16), None)
    with pytest.raises(ValueError):
        ops.linear_add_layer_norm(x, w, b, None, gamma, beta, 1e-5, return_sum=True)


def test_linear_add_layer_norm_graph_capture():
    ops = _ops()
    x, w, b, res, gamma, beta = _mk(512, 768, 3072, torch.float16, seed=2)
    out = torch.empty(512, 768, device="cuda", dtype=torch.float16)
    want = ops.linear_add_layer_norm(x, w, b, res, gamma, beta, 1e-5)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.linear_add_layer_norm(x, w, b, res, gamma, beta, 1e-5, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
