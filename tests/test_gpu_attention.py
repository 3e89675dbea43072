"""GPU parity of the fused tcgen05 attention kernel, through the C ABI, vs the oracle.
Tolerance: max-abs 2e-3 on the attention output (fp16/bf16 inputs, fp32 accumulate) -- BASELINE.json."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oeh_oracle as O  # noqa: E402  (checker only)

F32 = np.float32
TOL = 2e-3


def _ops():
    from outeffhop_b200 import ops
    return ops


def _mk(B, H, Tq, Tk, dtype, seed, layout="bhtd", bias=0.0):
    g = torch.Generator().manual_seed(seed)
    def t(T):
        x = torch.randn(B, H, T, 64, generator=g)
        return x
    q, k, v = t(Tq), t(Tk), t(Tk)
    q = q + bias / 8.0     # shifts the scores: exercises the "+1" term regimes of softmax_1
    q, k, v = (x.to(dtype) for x in (q, k, v))
    def place(x):
        if layout == "bthd":   # BERT: (B,T,H,d) memory viewed as (B,H,T,d)
            return x.permute(0, 2, 1, 3).contiguous().cuda().permute(0, 2, 1, 3)
        return x.cuda()
    return place(q), place(k), place(v)


def _oracle(q, k, v, **kw):
    f = lambda x: x.float().cpu().numpy()
    return O.attention_core(f(q), f(k), f(v), **kw)


KINDS = [(0, 0.0, 1.0), (1, 0.0, 1.0), (2, -0.025, 1.0), (3, -0.025, 1.1), (2, 0.0, 1.03)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("kind,gamma,eta", KINDS)
@pytest.mark.parametrize("shape", [(2, 3, 128, 128), (1, 2, 64, 64), (2, 2, 50, 77), (1, 2, 256, 384), (1, 1, 130, 1),
                                   (1, 2, 100, 300), (1, 2, 300, 100), (1, 1, 5, 33)])
def test_attention_vs_oracle(dtype, kind, gamma, eta, shape):
    ops = _ops()
    B, H, Tq, Tk = shape
    q, k, v = _mk(B, H, Tq, Tk, dtype, seed=100 + Tq + kind)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta)
    tol = TOL if dtype == torch.float16 else 8e-3   # bf16 output rounding alone is 2^-9 relative
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=tol, rtol=0)


@pytest.mark.parametrize("kind,gamma,eta", KINDS)
@pytest.mark.parametrize("bias", [-8.0, 0.0, 8.0])
def test_attention_score_regimes_bert_layout(kind, gamma, eta, bias):
    """C-BERT-like layout ((B,T,H,d) strided views in, (B,T,H*d) out) in the three score regimes."""
    ops = _ops()
    B, H, T = 2, 4, 128
    q, k, v = _mk(B, H, T, T, torch.float16, seed=7 + kind, layout="bthd", bias=bias)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, out_layout="bthd")
    assert out.shape == (B, T, H * 64)
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta)
    ref = ref.transpose(0, 2, 1, 3).reshape(B, T, H * 64)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("kind,gamma,eta", KINDS)
@pytest.mark.parametrize("T", [72, 128, 300, 512])
def test_attention_causal(kind, gamma, eta, T):
    ops = _ops()
    q, k, v = _mk(2, 2, T, T, torch.float16, seed=T + kind)
    out = ops.attention_fwd(q, k, v, scale=1.0 / 8, kind=kind, gamma=gamma, eta=eta, causal=True, out_layout="bhtd")
    ref = _oracle(q, k, v, scale=1.0 / 8, kind=kind, gamma=gamma, eta=eta, causal=True)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("kernel", ["tile", "ws", "ws:1", "ws:2"])
@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
def test_attention_kernel_override_same_result(kernel, kind, gamma, eta, override):
    """The three kernels (one-tile / multi-tile / warp-specialised) are interchangeable on a shape all of them accept."""
    ops = _ops()
    q, k, v = _mk(2, 2, 128, 128, torch.float16, seed=11 + kind, layout="bthd")
    logits = torch.randn(2, 2, 128).cuda()
    kw = dict(scale=0.125, kind=kind, gamma=gamma, eta=eta, gate_logits=logits, causal=(kind % 2 == 1))
    base = ops.attention_fwd(q, k, v, **kw).float()
    override("OEH_ATTN_KERNEL", kernel.split(":")[0])
    if ":" in kernel:
        override("OEH_WS_TILES", kernel.split(":")[1])      # Q tiles per CTA of the warp-specialised kernel
    other = ops.attention_fwd(q, k, v, **kw).float()
    assert (base - other).abs().max().item() <= 1e-3
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=(kind % 2 == 1), gate=O.sigmoid(logits.cpu().numpy()))
    ref = ref.transpose(0, 2, 1, 3).reshape(2, 128, 128)
    np.testing.assert_allclose(other.cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
def test_attention_dense_masks(kind, gamma, eta):
    """HF-4.31 additive masks: (B,1,1,Tk) padding mask (BERT) and (B,1,Tq,Tk) causal mask with clamp (OPT)."""
    ops = _ops()
    B, H, T = 2, 2, 200
    q, k, v = _mk(B, H, T, T, torch.float16, seed=31 + kind)
    fmin = torch.finfo(torch.float16).min
    pad = torch.zeros(B, 1, 1, T, dtype=torch.float16)
    pad[1, ..., 150:] = fmin
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=pad.cuda(), out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=pad.float().numpy())
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)
    cm = torch.full((T, T), fmin, dtype=torch.float16).triu(1)[None, None].expand(B, 1, T, T).contiguous()
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=cm.cuda(), clamp_min=True,
                            out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=True)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


def test_attention_all_masked_rows_match_reference_semantics():
    """Row masked entirely with finfo.min: vanilla -> uniform probs, softmax_1 -> zeros."""
    ops = _ops()
    B, H, T = 1, 1, 128
    q, k, v = _mk(B, H, T, T, torch.bfloat16, seed=5)
    m = torch.zeros(B, 1, T, T)
    m[:, :, 3, :] = torch.finfo(torch.float32).min
    for kind in (0, 1):
        out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, mask=m.cuda(), out_layout="bhtd")
        ref = _oracle(q, k, v, scale=0.125, kind=kind, mask=m.numpy())
        np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=8e-3, rtol=0)


@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
def test_attention_gate_epilogue(kind, gamma, eta):
    ops = _ops()
    B, H, T = 2, 3, 160
    q, k, v = _mk(B, H, T, T, torch.float16, seed=77 + kind)
    logits = torch.randn(B, H, T) * 2
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, gate_logits=logits.cuda(),
                            gate_scale=4.0, out_layout="bhtd")
    gate = O.sigmoid(logits.numpy())
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, gate=gate, gate_scale=4.0)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=2 * TOL, rtol=0)   # x4 gate scale


def test_attention_error_conventions():
    ops = _ops()
    from outeffhop_b200._lib import OehUnsupported
    q = torch.randn(1, 1, 16, 32, device="cuda", dtype=torch.float16)
    with pytest.raises(OehUnsupported):
        ops.attention_fwd(q, q, q, scale=1.0, kind=0)           # head dim 32 is not built
    q = torch.randn(1, 1, 16, 64, device="cuda", dtype=torch.float16)
    with pytest.raises(ValueError):
        ops.attention_fwd(q, q, q, scale=1.0, kind=9)
    with pytest.raises(ValueError):
        ops.attention_fwd(q, q[:, :, :8], q, scale=1.0, kind=0)
    e = torch.empty(0, 2, 16, 64, device="cuda", dtype=torch.float16)
    assert ops.attention_fwd(e, e, e, scale=1.0, kind=1).shape == (0, 16, 128)


@pytest.mark.parametrize("T,B", [(2048, 4), (8192, 1)])
def test_attention_full_size_properties(T, B):
    """BASELINE config-5 sizes: checked against a torch fp32 restatement on the device for a sample of
    heads, plus size-independent properties (rows of vanilla sum to 1 via V=ones; softmax_1 rows <= 1;
    linearity in V)."""
    ops = _ops()
    H = 12
    torch.manual_seed(T)
    q = torch.randn(B, H, T, 64, device="cuda", dtype=torch.float16)
    k = torch.randn(B, H, T, 64, device="cuda", dtype=torch.float16)
    v = torch.randn(B, H, T, 64, device="cuda", dtype=torch.float16)
    ones = torch.ones_like(v)
    row_sum = ops.attention_fwd(q, k, ones, scale=0.125, kind=0, out_layout="bhtd").float()
    assert (row_sum - 1).abs().max().item() < 2e-3
    row_sum1 = ops.attention_fwd(q, k, ones, scale=0.125, kind=1, out_layout="bhtd").float()
    assert row_sum1.max().item() <= 1 + 1e-3 and row_sum1.min().item() > 0
    o1 = ops.attention_fwd(q, k, v, scale=0.125, kind=1, out_layout="bhtd").float()
    o2 = ops.attention_fwd(q, k, v * 2, scale=0.125, kind=1, out_layout="bhtd").float()
    assert (o2 - 2 * o1).abs().max().item() < 4e-3
    for (b, h) in ((0, 0), (B - 1, H - 1)):
        s = q[b, h].float() @ k[b, h].float().t() * 0.125
        m = s.max(-1, keepdim=True).values
        e = torch.exp(s - m)
        p = e / (e.sum(-1, keepdim=True) + torch.exp(-m))
        ref = p @ v[b, h].float()
        assert (o1[b, h] - ref).abs().max().item() < TOL


@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
@pytest.mark.parametrize("before_gate", [False, True])
@pytest.mark.parametrize("T", [96, 300])
def test_attention_fused_int8_taps_vs_oracle(kind, gamma, eta, before_gate, T):
    """Fixed-range fake-quant taps folded into the kernel (scores before the mask, probs, context before/after the
    gate: quantized_bert.py:363,374,434 / quantized_opt.py:154,182,210) against the oracle's tap chain on IDENTICAL
    q,k,v.  Only tensor-core accumulation order differs, so at most a handful of codes may flip at bin edges."""
    ops = _ops()
    B, H = 2, 2
    q, k, v = _mk(B, H, T, T, torch.float16, seed=900 + kind + T)
    logits = torch.randn(B, H, T)
    gate = O.sigmoid(logits.numpy())
    rng = {"s": (-6.0, 7.0), "p": (0.0, 1.0), "c": (-0.9, 1.1)}
    qs = {n: O.asym_set_quant_range(lo, hi) for n, (lo, hi) in rng.items()}
    dev = {n: torch.tensor([float(x) for x in O.asym_scale_zero_point(*qs[n])], device="cuda") for n in qs}
    fq = {n: (lambda x, n=n: O.asym_quantize(x, *qs[n])) for n in qs}
    pad = torch.zeros(B, 1, 1, T, dtype=torch.float16)
    pad[1, ..., T - 20:] = torch.finfo(torch.float16).min
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=pad.cuda(),
                            gate_logits=logits.cuda(), qdq_scores=dev["s"], qdq_probs=dev["p"], qdq_ctx=dev["c"],
                            n_bits=8, ctx_qdq_before_gate=before_gate, out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=pad.float().numpy(), gate=gate,
                  qdq_scores=fq["s"], qdq_probs=fq["p"], qdq_ctx=fq["c"], ctx_qdq_before_gate=before_gate)
    diff = np.abs(out.float().cpu().numpy() - ref)
    step = float(qs["c"][0])
    assert diff.max() <= 1.05 * step + 1e-3, (diff.max(), step)           # never more than one context code
    assert (diff > 2e-3).mean() < 0.02, (diff > 2e-3).mean()              # and rarely even that


@pytest.mark.parametrize("persistent", ["0", "1"])
@pytest.mark.parametrize("tiles", ["1", "2"])
@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
@pytest.mark.parametrize("T,causal", [(700, False), (700, True), (1024, True)])
def test_attention_ws_tiles_per_cta(persistent, tiles, kind, gamma, eta, T, causal, override):
    """Both CTA shapes of the warp-specialised kernel (two Q tiles per CTA / two one-tile CTAs per SM), launched per work
    item and as persistent CTAs, on ragged and causal long sequences."""
    ops = _ops()
    override("OEH_WS_TILES", tiles)
    override("OEH_WS_PERSISTENT", persistent)
    q, k, v = _mk(1, 3, T, T, torch.float16, seed=T + kind)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal, out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
@pytest.mark.parametrize("causal", [False, True])
def test_attention_persistent_many_items_per_cta(kind, gamma, eta, causal, override):
    """More work items than SM slots: every persistent CTA walks several (Q tile group, head, batch) items, so the ring /
    parity bookkeeping carries across items (odd and even step counts per item)."""
    ops = _ops()
    override("OEH_WS_PERSISTENT", "1")
    for T in (320, 448):                                   # 5 and 7 key steps of 64: odd -> the stream parities swap per item
        q, k, v = _mk(7, 12, T, T, torch.float16, seed=T + kind)       # 7*12*2 = 168 items > 148 SMs
        out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal, out_layout="bhtd")
        ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal)
        np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("persistent", ["0", "1"])
@pytest.mark.parametrize("chunk", ["0", "1", "5", "1000"])
@pytest.mark.parametrize("kind,gamma,eta", [KINDS[1], KINDS[3]])
def test_attention_causal_work_order(persistent, chunk, kind, gamma, eta, override):
    """Causal launches number their work items heaviest-first over chunks of (batch, head) pairs (causal_item in
    csrc/attn_common.cuh).  Any chunk size -- per-head order (0), chunks that do not divide B*H (5 of 14), one chunk for
    everything -- must visit every (Q tile, head, batch) exactly once: the result does not depend on it."""
    ops = _ops()
    override("OEH_WS_PERSISTENT", persistent)
    override("OEH_CAUSAL_CHUNK", chunk)
    q, k, v = _mk(2, 7, 600, 600, torch.float16, seed=11 + kind)        # 14 pairs x 5 Q tiles (3 tile groups of 2)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=True, out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=True)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("form", ["split", "regs", "chunked"])
@pytest.mark.parametrize("kind,gamma,eta", KINDS)
@pytest.mark.parametrize("shape,causal", [((2, 3, 128, 128), False), ((2, 2, 50, 77), False), ((1, 2, 100, 100), True),
                                          ((1, 1, 130, 1), False), ((2, 2, 300, 100), False), ((1, 2, 64, 64), True),
                                          ((1, 2, 97, 97), False), ((1, 1, 33, 65), False)])
def test_attention_one_tile_softmax_shapes(form, kind, gamma, eta, shape, causal, override):
    """One-tile kernels (Tk <= 128) against the oracle on full, ragged (rows of the last warp beyond Tq: the coalesced store must
    clip them), causal and single-key rows, and key counts that leave the upper half of the row with 0, 1 or 2 chunks:
    two threads per row (attn_fwd_one2.cuh, the default), and the one-thread kernel with the S row held in registers (one TMEM
    read, 3 CTAs / SM) or walked in chunked TMEM passes (4 CTAs / SM)."""
    ops = _ops()
    if form == "split":
        override("OEH_ONE_SPLIT", "1")
    else:
        override("OEH_ONE_SPLIT", "0")
        override("OEH_ONE_REGS", "1" if form == "regs" else "0")
    B, H, Tq, Tk = shape
    q, k, v = _mk(B, H, Tq, Tk, torch.float16, seed=Tq + Tk + kind)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal, out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("kind,gamma,eta", KINDS)
def test_attention_one_tile_store_does_not_touch_neighbours(dtype, kind, gamma, eta):
    """The one-tile kernel writes the context as full 128-byte lines staged per warp (store_rows_coalesced).  With the
    reference's (B, T, H*d) layout and Tq = 100 the last warp of every tile holds rows beyond Tq, and a head's row is
    flanked by the other heads' bytes: a canary-filled output must come back with exactly the rows of the result written."""
    ops = _ops()
    B, H, T = 3, 5, 100
    q, k, v = _mk(B, H, T, T, dtype, seed=9 + kind, layout="bthd")
    full = torch.full((B, T + 4, H * 64), 7.0, device="cuda", dtype=dtype)       # 4 canary rows behind every batch
    want = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, out_layout="bhtd")
    args = dict(scale=0.125, kind=kind, gamma=gamma, eta=eta)
    # write through a strided view: out rows of batch b start at full[b, 0]
    from outeffhop_b200 import _lib as L
    import ctypes as C
    a = L.OehAttnArgs()
    view = full[:, :T].view(B, T, H, 64).permute(0, 2, 1, 3)
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), view.data_ptr()
    for name, t in (("q_strides", q), ("k_strides", k), ("v_strides", v), ("out_strides", view)):
        setattr(a, name, (C.c_int64 * 3)(t.stride(0), t.stride(1), t.stride(2)))
    a.B, a.H, a.Tq, a.Tk, a.d = B, H, T, T, 64
    a.dtype = 1 if dtype == torch.float16 else 2
    a.scale, a.kind, a.gamma, a.eta = args["scale"], kind, gamma, eta
    a.gate_scale = 1.0
    a.stream = torch.cuda.current_stream().cuda_stream
    L.check(L.lib().oeh_attn_fwd(C.byref(a)))
    torch.cuda.synchronize()
    assert torch.equal(view.contiguous(), want)
    assert bool((full[:, T:] == 7.0).all())


@pytest.mark.parametrize("mid", ["0", "1", "3"])      # 0: two-pass / persistent kernels; 1: single pass, 4 threads per row; 3: 2 per row
@pytest.mark.parametrize("kind,gamma,eta", KINDS)
@pytest.mark.parametrize("shape,causal", [((2, 3, 512, 512), True), ((2, 3, 512, 512), False), ((1, 2, 130, 130), True),
                                          ((2, 2, 200, 333), False), ((1, 2, 100, 300), True), ((1, 3, 384, 384), True),
                                          ((1, 2, 257, 511), True), ((1, 1, 40, 129), False)])
def test_attention_mid_kernel_single_pass(mid, kind, gamma, eta, shape, causal, override):
    """128 < Tk <= 512: the single-pass kernel (whole S row block in TMEM, two threads per row; attn_fwd_mid.cuh) forced on
    for every kind (OEH_MID_KERNEL=1) and off (=0: the two-pass / persistent kernels) -- same oracle, same tolerance; full,
    ragged, causal with Tq != Tk, one to four K/V tiles per Q tile."""
    ops = _ops()
    override("OEH_MID_KERNEL", mid)
    B, H, Tq, Tk = shape
    q, k, v = _mk(B, H, Tq, Tk, torch.float16, seed=Tq + Tk + kind, bias=4.0 if kind >= 2 else 0.0)
    logits = torch.randn(B, H, Tq)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal,
                            gate_logits=logits.cuda(), out_layout="bhtd")
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal, gate=O.sigmoid(logits.numpy()))
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=TOL, rtol=0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
def test_attention_mid_kernel_taps_and_masks(kind, gamma, eta, dtype, override):
    """The single-pass kernel with everything on: the three fixed-range taps (OPT order), an additive (B,1,Tq,Tk) mask with
    the finfo.min clamp, a per-token gate -- against the oracle's chain and against the two-pass kernel on the same inputs."""
    ops = _ops()
    B, H, T = 2, 3, 448
    q, k, v = _mk(B, H, T, T, dtype, seed=4400 + kind)
    logits = torch.randn(B, H, T)
    rng = {"s": (-6.0, 7.0), "p": (0.0, 1.0), "c": (-0.9, 1.1)}
    qs = {n: O.asym_set_quant_range(lo, hi) for n, (lo, hi) in rng.items()}
    dev = {n: torch.tensor([float(x) for x in O.asym_scale_zero_point(*qs[n])], device="cuda") for n in qs}
    fq = {n: (lambda x, n=n: O.asym_quantize(x, *qs[n])) for n in qs}
    fmin = torch.finfo(dtype).min
    cm = torch.full((T, T), fmin, dtype=dtype).triu(1)[None, None].expand(B, 1, T, T).contiguous()
    kw = dict(scale=0.125, kind=kind, gamma=gamma, eta=eta, mask=cm.cuda(), clamp_min=True, gate_logits=logits.cuda(),
              qdq_scores=dev["s"], qdq_probs=dev["p"], qdq_ctx=dev["c"], n_bits=8, ctx_qdq_before_gate=True,
              out_layout="bhtd")
    override("OEH_MID_KERNEL", "1")
    out = ops.attention_fwd(q, k, v, **kw).float().cpu().numpy()
    override("OEH_MID_KERNEL", "3")
    out2 = ops.attention_fwd(q, k, v, **kw).float().cpu().numpy()
    assert (np.abs(out - out2) > 2e-3).mean() < 0.02                    # 4 vs 2 threads per row: same arithmetic, other order
    override("OEH_MID_KERNEL", "0")
    two_pass = ops.attention_fwd(q, k, v, **kw).float().cpu().numpy()
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=True, gate=O.sigmoid(logits.numpy()),
                  qdq_scores=fq["s"], qdq_probs=fq["p"], qdq_ctx=fq["c"], ctx_qdq_before_gate=True)
    step = float(qs["c"][0])
    for got in (out, two_pass):
        diff = np.abs(got - ref)
        assert diff.max() <= 1.05 * step + (1e-3 if dtype == torch.float16 else 8e-3), (diff.max(), step)
        assert (diff > 2e-3).mean() < (0.02 if dtype == torch.float16 else 0.2), (diff > 2e-3).mean()
    assert (np.abs(out - two_pass) > 2e-3).mean() < 0.02


@pytest.mark.parametrize("kind,gamma,eta", KINDS[:4])
@pytest.mark.parametrize("shape,causal,layout", [((2, 3, 128, 128), False, "bthd"), ((1, 2, 300, 300), True, "bhtd"),
                                                 ((2, 2, 77, 130), False, "bthd"), ((1, 2, 1024, 1024), True, "bhtd")])
def test_attention_fp32_tensors_fp16_operands(kind, gamma, eta, shape, causal, layout):
    """fp32 q / k / v / out (the reference's eval precision) through the fused kernels: Q, K, V are rounded to fp16 OPERANDS
    inside oeh_attn_fwd (fp32 accumulation, fp32 softmax), so the result carries the fp16 path's error against the fp32
    oracle -- the same 2e-3 max-abs contract -- not an fp32 GEMM's."""
    ops = _ops()
    B, H, Tq, Tk = shape
    q, k, v = _mk(B, H, Tq, Tk, torch.float32, seed=Tq + kind, layout=layout)
    logits = torch.randn(B, H, Tq)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal, gate_logits=logits.cuda(),
                            out_layout=layout)
    assert out.dtype == torch.float32
    ref = _oracle(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal, gate=O.sigmoid(logits.numpy()))
    if layout == "bthd":
        ref = ref.transpose(0, 2, 1, 3).reshape(B, Tq, H * 64)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=TOL, rtol=0)
