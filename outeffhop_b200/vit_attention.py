"""Drop-in ``ViTSelfAttentionWithExtras`` (transformers_language/models/vit_attention.py:77-275): the timm-style
attention block of the reference's ViT scripts (fused ``qkv`` Linear, optional q/k LayerNorm, ``proj`` inside), with
the same ``softmax_fn`` / gate options as the BERT module.  Non-causal, no mask, T = 197 for ViT-S/B at 224 px.

The core ``softmax_fn(q k^T * scale) v`` (vit_attention.py:54-75 and :221-229) runs as one fused ``oeh_attn_fwd``
launch on the strided q/k/v views of the qkv projection (no unbind copies); gates are folded into its epilogue
(:233-270).  SURVEY section 8f rank 4."""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import ops
from .bert_attention import FUSED_HEAD_DIM, AttentionGateType, GateMixin, _check_no_grad
from .softmax import as_softmax_fn


class ViTSelfAttentionWithExtras(GateMixin, nn.Module):
    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False, qk_norm: bool = False,
                 attn_drop: float = 0., proj_drop: float = 0., norm_layer=nn.LayerNorm,
                 softmax_fn=torch.nn.functional.softmax, gamma=None, ssm_eps=None, tau=None, skip_attn=False,
                 attn_gate_type=AttentionGateType.none, attn_gate_init=None, attn_gate_mlp=False,
                 attn_gate_mlp2=False, attn_gate_linear_all_features=False, fine_tuning=False,
                 max_seq_length=None) -> None:
        super().__init__()
        assert dim % num_heads == 0, 'dim should be divisible by num_heads'
        self.num_attention_heads = num_heads
        self.attention_head_size = dim // num_heads
        self.all_head_size = dim          # the reference reads it for linear_all_features without defining it (:142)
        self.scale = self.attention_head_size ** -0.5
        self.fused_attn = True
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = norm_layer(self.attention_head_size) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.attention_head_size) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.attn_scores = nn.Identity()
        self.attn_probs_before_dropout = nn.Identity()
        self.attn_probs_after_dropout = nn.Identity()
        self.gamma, self.ssm_eps, self.tau, self.max_seq_length = gamma, ssm_eps, tau, max_seq_length
        self.softmax_fn = softmax_fn
        self.skip_attn = skip_attn
        self.materialize_probs = False
        self._init_gate(num_heads, self.attention_head_size, self.all_head_size, attn_gate_type, attn_gate_init,
                        attn_gate_mlp, attn_gate_mlp2, attn_gate_linear_all_features, fine_tuning, fine_tuning_std=0.01)

    def transpose_for_scores(self, x: torch.Tensor) -> torch.Tensor:
        new_x_shape = x.size()[:-1] + (self.num_attention_heads, self.attention_head_size)
        return x.view(new_x_shape).permute(0, 2, 1, 3)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("outeffhop_b200.ViTSelfAttentionWithExtras runs on CUDA only (no CPU fallback)")
        _check_no_grad(x, self.qkv.weight)
        hidden_states = x
        B, N, C = x.shape
        H, d = self.num_attention_heads, self.attention_head_size
        qkv = self.qkv(x).reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)                       # :205
        q, k, v = qkv.unbind(0)                                                                # strided (B,H,N,d) views
        q, k = self.q_norm(q), self.k_norm(k)
        gate_logits = self._gate_logits(hidden_states, H, d)                                   # :233-263
        fn = as_softmax_fn(self.softmax_fn)
        training_dropout = self.training and self.attn_drop.p > 0
        if (fn is not None and d == FUSED_HEAD_DIM and x.dtype in (torch.float16, torch.bfloat16)
                and not training_dropout and not self.materialize_probs and not self._probes_hooked()):
            ctx = ops.attention_fwd(q, k, v, scale=self.scale, kind=fn.kind, gamma=fn.gamma, eta=fn.eta,
                                    gate_logits=gate_logits, gate_scale=self.effective_gate_scale, out_layout="bthd")
        else:                                                                                  # :221-229, S/P materialised
            attn = (q * self.scale) @ k.transpose(-2, -1)
            attn = fn(attn, dim=-1) if fn is not None else self.softmax_fn(attn, dim=-1)
            attn = self.attn_probs_after_dropout(self.attn_drop(self.attn_probs_before_dropout(attn)))
            ctx = attn @ v
            if gate_logits is not None:
                ctx = ctx * (torch.sigmoid(gate_logits).to(ctx.dtype).unsqueeze(-1) * self.effective_gate_scale)
            ctx = ctx.transpose(1, 2).reshape(B, N, C)                                         # :272
        return self.proj_drop(self.proj(ctx))                                                  # :273-274
