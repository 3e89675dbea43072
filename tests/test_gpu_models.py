"""Model-level parity on the GPU: BERT-base / OPT-125m shaped random-init models with the drop-in attention
(fp16, fused kernel, CUDA graph) against the same weights evaluated in fp32 with a plain-torch restatement of the
reference attention (bert_attention.py:222-337 / opt_attention.py:204-322 + vutils/softmax_1.py).
north_star tolerance: perplexity within 0.1 %."""
import copy
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_softmax(s, kind, gamma, eta):
    if kind in (0, 2):
        p = torch.softmax(s, dim=-1)
    else:                                                     # vutils/softmax_1.py:11-21
        m = s.max(dim=-1, keepdim=True).values
        e = torch.exp(s - m)
        p = e / (e.sum(dim=-1, keepdim=True) + torch.exp(-m))
    if kind >= 2:                                             # softmax.py:12-13
        p = torch.clip(p * (eta - gamma) + gamma, 0, 1)
    return p


class RefBertSelfAttention(torch.nn.Module):
    """Plain-torch fp32 restatement of BertSelfAttentionWithExtras.forward (eval, self-attention)."""

    def __init__(self, ours):
        super().__init__()
        self.q, self.k, self.v = (copy.deepcopy(m).float() for m in (ours.query, ours.key, ours.value))
        self.H, self.d = ours.num_attention_heads, ours.attention_head_size
        self.fn = ours.softmax_fn
        self.alpha = copy.deepcopy(ours.alpha).float() if ours.alpha is not None else None

    def forward(self, hidden_states, attention_mask=None, **kw):
        B, T, _ = hidden_states.shape
        sp = lambda x: x.view(B, T, self.H, self.d).permute(0, 2, 1, 3)
        q, k, v = sp(self.q(hidden_states)), sp(self.k(hidden_states)), sp(self.v(hidden_states))
        s = q @ k.transpose(-1, -2) / math.sqrt(self.d)
        if attention_mask is not None:
            s = s + attention_mask
        ctx = _ref_softmax(s, self.fn.kind, self.fn.gamma, self.fn.eta) @ v
        if self.alpha is not None:
            x = sp(hidden_states)
            gate = torch.sigmoid(torch.stack([self.alpha[h](x[:, h]) for h in range(self.H)], dim=1))
            ctx = ctx * gate
        return ctx.permute(0, 2, 1, 3).reshape(B, T, self.H * self.d), None


@pytest.mark.parametrize("key,gated", [("clippedsoftmax1(-.025:1)", False), ("softmax1", True)])
def test_bert_base_eval_perplexity_matches_fp32_reference(key, gated):
    from outeffhop_b200 import eval_harness as EH
    kw = dict(attn_gate_type="conditional_per_token", attn_gate_init=0.25) if gated else {}
    model = EH.build_bert_base(seed=3000, dtype=torch.float16, attn_softmax=key, fuse_residual=False, **kw)
    ref = copy.deepcopy(model).float()                                  # reference keeps HF's ATen add + LayerNorm
    EH.fuse_bert_residual_blocks(model)
    for lo, lr in zip(model.bert.encoder.layer, ref.bert.encoder.layer):
        lr.attention.self = RefBertSelfAttention(lo.attention.self).cuda()
    batches = [EH.synthetic_mlm_batch(32, 128, seed=3000 + i) for i in range(3)]
    step = EH.make_bert_eval_step(model, batches[0], use_graph=True)
    eager = EH.make_bert_eval_step(model, batches[0], use_graph=False)
    ours, refl = [], []
    with torch.no_grad():
        for b in batches:
            step.load(b)
            ours.append(step.run().item())
            eager.load(b)
            assert abs(eager.run().item() - ours[-1]) < 2e-3           # graph replay ~ eager launch (cuBLAS may pick other algos)
            refl.append(EH.bert_mlm_loss_aten(ref, b["input_ids"].cuda(), b["labels"].cuda()).item())
    ppl_ours = math.exp(sum(ours) / len(ours))
    ppl_ref = math.exp(sum(refl) / len(refl))
    assert abs(ppl_ours / ppl_ref - 1) < 1e-3, (ppl_ours, ppl_ref)
    assert EH.aggregate_perplexity(torch.tensor(sum(ours), device="cuda"), len(ours)) == pytest.approx(ppl_ours, rel=1e-6)


class RefOPTAttention(torch.nn.Module):
    """Plain-torch fp32 restatement of OPTAttentionWithExtras.forward (eval, self-attention)."""

    def __init__(self, ours):
        super().__init__()
        self.q, self.k, self.v, self.o = (copy.deepcopy(m).float() for m in (ours.q_proj, ours.k_proj, ours.v_proj, ours.out_proj))
        self.H, self.d, self.scaling = ours.num_heads, ours.head_dim, ours.scaling
        self.fn = ours.softmax_fn
        self.alpha = copy.deepcopy(ours.alpha).float() if ours.alpha is not None else None

    def forward(self, hidden_states, attention_mask=None, **kw):
        B, T, _ = hidden_states.shape
        sp = lambda x: x.view(B, T, self.H, self.d).permute(0, 2, 1, 3)
        q, k, v = sp(self.q(hidden_states) * self.scaling), sp(self.k(hidden_states)), sp(self.v(hidden_states))
        s = q @ k.transpose(-1, -2)
        if attention_mask is not None:
            s = torch.max(s + attention_mask, torch.tensor(torch.finfo(s.dtype).min, device=s.device))
        ctx = _ref_softmax(s, self.fn.kind, self.fn.gamma, self.fn.eta) @ v
        if self.alpha is not None:
            x = sp(hidden_states)
            ctx = ctx * torch.sigmoid(torch.stack([self.alpha[h](x[:, h]) for h in range(self.H)], dim=1))
        return self.o(ctx.permute(0, 2, 1, 3).reshape(B, T, self.H * self.d)), None


@pytest.mark.parametrize("key,gated,causal_hint", [("vanilla", True, True), ("softmax1", False, False)])
def test_opt_125m_eval_perplexity_matches_fp32_reference(key, gated, causal_hint):
    """BASELINE config 3: OPT-125m causal fp16, conditional_per_token gate (init 0.25), block 512, batch 4."""
    from outeffhop_b200 import eval_harness as EH
    kw = dict(attn_gate_type="conditional_per_token", attn_gate_init=0.25) if gated else {}
    model = EH.build_opt_125m(seed=5678, dtype=torch.float16, attn_softmax=key, causal_hint=causal_hint,
                              fuse_residual=False, **kw)
    ref = copy.deepcopy(model).float()                                  # reference keeps HF's ATen decoder layers
    EH.fuse_opt_decoder_layers(model)
    for lo, lr in zip(model.model.decoder.layers, ref.model.decoder.layers):
        lr.self_attn = RefOPTAttention(lo.self_attn).cuda()
    batches = [EH.synthetic_clm_batch(4, 512, seed=5678 + i) for i in range(2)]
    step = EH.make_opt_eval_step(model, batches[0], use_graph=True)
    mask = EH.make_causal_mask(4, 512, torch.float32, "cuda")
    pos = torch.arange(512, device="cuda").unsqueeze(0).expand(4, 512)
    ours, refl = [], []
    with torch.no_grad():
        for b in batches:
            step.load(b)
            ours.append(step.run().item())
            refl.append(EH.opt_clm_loss(ref, b["input_ids"].cuda(), b["labels"].cuda(), mask, pos, aten_loss=True).item())
    ppl_ours, ppl_ref = math.exp(sum(ours) / 2), math.exp(sum(refl) / 2)
    assert abs(ppl_ours / ppl_ref - 1) < 1e-3, (ppl_ours, ppl_ref)


def test_swapped_modules_inherit_eval_mode_and_use_the_fused_kernel():
    """The module swap must keep the model's train/eval state (a fresh nn.Module defaults to training=True, which
    would silently route through the unfused dropout path)."""
    from outeffhop_b200 import eval_harness as EH, ops
    model = EH.build_bert_base(num_hidden_layers=1)
    att = model.bert.encoder.layer[0].attention.self
    assert not att.training
    calls = []
    orig = ops.attention_fwd
    ops.attention_fwd = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        b = EH.synthetic_mlm_batch(2, 128)
        with torch.no_grad():
            EH.bert_mlm_loss(model, b["input_ids"].cuda(), b["labels"].cuda())
    finally:
        ops.attention_fwd = orig
    assert calls, "fused attention kernel was not used"
