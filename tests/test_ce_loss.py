"""Eval-loss row: the fused mean cross-entropy kernel (oeh_ce_loss) against the oracle and against ATen's
CrossEntropyLoss -- the primitive behind HF's `outputs.loss` that the reference's eval loops aggregate
(validate_mlm_config.py:489-496, validate_clm.py:555-561).  Tolerance: fp32 evaluation of the same half-precision
logits, |diff| <= 2e-6 * max(1, |loss|) (perplexity contract: 0.1 %)."""
import numpy as np
import pytest
import torch

from oracle import oeh_oracle as O  # checker only


def _case(rows, vocab, seed, p_ignore=0.85, dtype=np.float32, spread=3.0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((rows, vocab)) * spread).astype(dtype)
    lab = rng.integers(0, vocab, rows).astype(np.int64)
    lab[rng.random(rows) < p_ignore] = -100
    lab[0] = rng.integers(0, vocab)
    return x, lab


@pytest.mark.parametrize("rows,vocab", [(7, 5), (64, 30522), (33, 1001)])
def test_oracle_equals_aten_cross_entropy(rows, vocab):
    x, lab = _case(rows, vocab, 11)
    mean, per_row = O.ce_loss_mean(x, lab)
    ref = torch.nn.functional.cross_entropy(torch.from_numpy(x).double(), torch.from_numpy(lab))
    assert abs(mean - ref.item()) < 1e-12 * max(1.0, abs(ref.item()))
    ref_rows = torch.nn.functional.cross_entropy(torch.from_numpy(x).double(), torch.from_numpy(lab), reduction="none")
    np.testing.assert_allclose(per_row, ref_rows.numpy(), rtol=1e-12, atol=1e-12)


def test_oracle_all_ignored_is_nan_like_aten():
    x, lab = _case(4, 9, 3)
    lab[:] = -100
    assert np.isnan(O.ce_loss_mean(x, lab)[0])
    assert torch.isnan(torch.nn.functional.cross_entropy(torch.from_numpy(x), torch.from_numpy(lab)))


def test_symbols_declared():
    import os
    from outeffhop_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "oeh.h")).read()
    for name in ("oeh_ce_loss", "oeh_ce_loss_workspace_bytes", "oeh_ce_loss_shifted", "oeh_embed_sum"):
        assert name + "(" in hdr and name in _lib.SYMBOLS
        getattr(_lib.lib(), name)


# ------------------------------------------------------------------------------------------------ GPU parity
_T = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("rows,vocab,pad", [(1, 1, 0), (5, 7, 0), (5, 7, 1), (64, 30522, 6), (64, 30522, 0), (257, 50272, 0),
                                            (4096, 777, 7)])
def test_ce_loss_vs_oracle_and_aten(dtype, rows, vocab, pad):
    from outeffhop_b200 import ops
    x, lab = _case(rows, vocab, 1000 + rows + vocab)
    xt = torch.from_numpy(x).cuda().to(_T[dtype])
    buf = torch.full((rows, vocab + pad), 1e4, device="cuda", dtype=_T[dtype])   # poison in the pad columns
    buf[:, :vocab] = xt
    logits = buf[:, :vocab] if pad == 0 else buf                                 # padded: width > vocab
    labels = torch.from_numpy(lab).cuda()
    loss, per_row = ops.ce_loss(logits, labels, vocab=vocab, return_row_losses=True)
    want, want_rows = O.ce_loss_mean(xt.float().cpu().numpy(), lab)
    assert abs(loss.item() - want) <= 2e-6 * max(1.0, abs(want))
    np.testing.assert_allclose(per_row.cpu().numpy(), want_rows, rtol=2e-6, atol=2e-6)
    aten = torch.nn.functional.cross_entropy(xt.float(), labels)
    assert abs(loss.item() - aten.item()) <= 2e-6 * max(1.0, abs(aten.item()))


@pytest.mark.gpu
def test_ce_loss_edge_cases():
    from outeffhop_b200 import ops
    x = torch.randn(8, 40, device="cuda", dtype=torch.float16)
    lab = torch.full((8,), -100, dtype=torch.int64, device="cuda")
    assert torch.isnan(ops.ce_loss(x, lab))                                       # nothing labelled: NaN like ATen
    lab[3] = 5
    x[3, 7] = float("-inf")                                                       # -inf logits are legal
    want = torch.nn.functional.cross_entropy(x.float(), lab)
    assert abs(ops.ce_loss(x, lab).item() - want.item()) < 1e-5
    x[3] = float("-inf")
    x[3, 5] = 0.0                                                                 # one finite logit: loss 0
    assert ops.ce_loss(x, lab).item() == 0.0
    big = torch.full((2, 16), 60000.0, device="cuda", dtype=torch.float16)        # no overflow in the max frame
    lab2 = torch.tensor([1, 2], device="cuda")
    assert abs(ops.ce_loss(big, lab2).item() - float(np.log(16.0))) < 1e-5
    with pytest.raises(ValueError):
        ops.ce_loss(x, lab[:4])
    with pytest.raises(RuntimeError):
        ops.ce_loss(x.cpu(), lab.cpu())


@pytest.mark.gpu
def test_ce_loss_non_contiguous_rows_and_graph_capture():
    from outeffhop_b200 import ops
    base = torch.randn(50, 2, 300, device="cuda", dtype=torch.float16)
    x = base[:, 1, :297]                                                          # row stride 600, unaligned width
    lab = torch.randint(0, 297, (50,), device="cuda")
    want = torch.nn.functional.cross_entropy(x.float(), lab).item()
    assert abs(ops.ce_loss(x, lab).item() - want) < 2e-6 * max(1, abs(want))
    out = torch.zeros((), device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out.copy_(ops.ce_loss(x, lab))
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out.copy_(ops.ce_loss(x, lab))
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert abs(out.item() - want) < 2e-6 * max(1, abs(want))


@pytest.mark.gpu
def test_bert_step_fused_loss_equals_aten_loss():
    """The harness step (padded decoder GEMM + oeh_ce_loss) against the same model through ATen only."""
    from outeffhop_b200 import eval_harness as EH
    model = EH.build_bert_base(seed=3000, num_hidden_layers=2)
    b = EH.synthetic_mlm_batch(8, 128, seed=5)
    ids, lab = b["input_ids"].cuda(), b["labels"].cuda()
    with torch.no_grad():
        ours = EH.bert_mlm_loss(model, ids, lab).item()
        aten = EH.bert_mlm_loss_aten(model, ids, lab).item()
    assert abs(ours - aten) < 2e-3 * abs(aten)          # two different GEMM kernels produce the fp16 logits


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16"])
@pytest.mark.parametrize("B,T,vocab", [(1, 2, 5), (4, 512, 1000), (3, 7, 50272)])
def test_ce_loss_shifted_equals_hf_shift(dtype, B, T, vocab):
    """HF's causal-LM loss (validate_clm.py:555-561): CrossEntropyLoss(logits[:, :-1], labels[:, 1:]); the kernel applies the
    shift to the label index.  Same B*(T-1) terms: bit-identical to oeh_ce_loss on a shifted copy of the labels, and equal to
    ATen within the fp32 tolerance of this file."""
    from outeffhop_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + T)
    logits = (torch.randn(B, T, vocab, generator=g) * 3).to(_T[dtype]).cuda()
    labels = torch.randint(0, vocab, (B, T), generator=g).cuda()
    labels[0, -1] = -100                                         # an ignored label that the shift moves onto row T-2
    loss, rows = ops.ce_loss(logits.view(-1, vocab), labels.view(-1), vocab=vocab, shift_block=T, return_row_losses=True)
    shifted = torch.full_like(labels, -100)
    shifted[:, :-1] = labels[:, 1:]
    loss2, rows2 = ops.ce_loss(logits.view(-1, vocab), shifted.view(-1), vocab=vocab, return_row_losses=True)
    assert torch.equal(torch.nan_to_num(loss, nan=-1.0), torch.nan_to_num(loss2, nan=-1.0)) and torch.equal(rows, rows2)   # (B=1, T=2: no labelled row, NaN like ATen)
    aten = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, vocab).float(), labels[:, 1:].reshape(-1))
    if T > 1 and bool((labels[:, 1:] != -100).any()):
        assert abs(loss.item() - aten.item()) <= 2e-6 * max(1.0, abs(aten.item()))
    with pytest.raises(ValueError):
        ops.ce_loss(logits.view(-1, vocab), labels.view(-1), vocab=vocab, shift_block=T + 1 if (B * T) % (T + 1) else B * T + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
def test_embed_sum_bit_exact_vs_aten(dtype):
    """oeh_embed_sum == embed_tokens(ids) + embed_positions(position_ids) of HF's OPT decoder (offset 2 added inside the
    positional module): gathers and one add in the tensor dtype -- bit-identical."""
    from outeffhop_b200 import ops
    g = torch.Generator().manual_seed(5)
    dt = _T[dtype]
    word = torch.randn(1000, 768, generator=g).to(dt).cuda()
    pos = torch.randn(130, 768, generator=g).to(dt).cuda()
    ids = torch.randint(0, 1000, (3, 128), generator=g).cuda()
    pids = torch.arange(128).expand(3, 128).cuda()
    want = torch.nn.functional.embedding(ids, word) + torch.nn.functional.embedding(pids + 2, pos)
    assert torch.equal(ops.embed_sum(ids, word, pos, pids, 2), want)
    assert torch.equal(ops.embed_sum(ids, word, pos, None, 2), want)          # default positions 0..T-1
    pids2 = torch.randint(0, 128, (3, 128), generator=g).cuda()
    want2 = torch.nn.functional.embedding(ids, word) + torch.nn.functional.embedding(pids2 + 2, pos)
    assert torch.equal(ops.embed_sum(ids, word, pos, pids2, 2), want2)
    with pytest.raises(ValueError):
        ops.embed_sum(ids, word, pos[:, :64].contiguous(), pids, 2)
