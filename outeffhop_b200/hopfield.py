"""Drop-in ``Association`` / ``Hopfield`` / ``HopfieldPooling`` of the reference's STanHop time-series models
(STanHop_time_seeries/cross_models/hopfield.py:19-51, :53-90, :92-146): einsum attention
``softmax_fn(scale * Q K^T) V`` over (B, L, H, E) tensors with ``mode`` in {softmax, softmax1, clip, clip_softmax1}
(cross_models/softmax_1.py, clip_softmax.py).  Same math as the BERT path, so the core is one ``oeh_attn_fwd`` launch
(head dim 64, fp16/bf16); other head dims / fp32 use the row-softmax kernel between two library GEMMs.
``entmax`` / ``sparsemax`` are different operators and stay out of scope.  The reference's ``ClipSoftmax_1.__init__``
raises TypeError (it calls ``super(ClipSoftmax, self)``, clip_softmax.py:52), so ``mode='clip_softmax1'`` never ran
there; here it computes what the module was written to compute.  SURVEY section 8f rank 4."""
from __future__ import annotations

from math import sqrt

import torch
from torch import nn

from . import ops
from .bert_attention import FUSED_HEAD_DIM, _check_no_grad
from .softmax import KIND_CLIPPED, KIND_CLIPPED1, KIND_SOFTMAX1, KIND_VANILLA, SoftmaxFn


class Association(nn.Module):
    def __init__(self, scale=None, attention_dropout=0.1, eta=1.1, gamma=-0.1, mode='entmax'):
        super().__init__()
        self.scale = scale
        self.dropout = nn.Dropout(attention_dropout)
        if mode == 'softmax1':
            self.softmax = SoftmaxFn(KIND_SOFTMAX1)
        elif mode == 'softmax':
            self.softmax = SoftmaxFn(KIND_VANILLA)
        elif mode == 'clip':
            self.softmax = SoftmaxFn(KIND_CLIPPED, gamma, eta)
        elif mode == 'clip_softmax1':
            self.softmax = SoftmaxFn(KIND_CLIPPED1, gamma, eta)
        elif mode in ('entmax', 'sparsemax'):
            raise NotImplementedError(f"mode={mode!r}: entmax / sparsemax are not softmax-family kernels of this library")
        else:
            raise ValueError(f"unknown mode {mode!r}")
        self.mode = mode

    def forward(self, queries, keys, values):
        B, L, H, E = queries.shape
        _, S, _, D = values.shape
        if not queries.is_cuda:
            raise RuntimeError("outeffhop_b200.Association runs on CUDA only (no CPU fallback)")
        _check_no_grad(queries, keys, values)
        scale = self.scale or 1. / sqrt(E)
        fn = self.softmax
        q, k, v = queries.permute(0, 2, 1, 3), keys.permute(0, 2, 1, 3), values.permute(0, 2, 1, 3)   # (B,H,*,E) views
        training_dropout = self.training and self.dropout.p > 0
        if (E == FUSED_HEAD_DIM and D == FUSED_HEAD_DIM and queries.dtype in (torch.float16, torch.bfloat16)
                and not training_dropout and all(t.stride(-1) == 1 for t in (q, k, v))):
            out = ops.attention_fwd(q, k, v, scale=scale, kind=fn.kind, gamma=fn.gamma, eta=fn.eta, out_layout="bthd")
            return out.view(B, L, H, D)                                                        # "blhd", contiguous
        scores = torch.einsum("blhe,bshe->bhls", queries, keys)                                # :46
        A = self.dropout(ops.softmax_rows(scores, fn.kind, fn.gamma, fn.eta, in_scale=scale).view(scores.shape))   # :47
        return torch.einsum("bhls,bshd->blhd", A, values).contiguous()                         # :48-50


class Hopfield(nn.Module):
    """hopfield.py:53-90 (note the reference projects VALUES from the projected keys, :77)."""

    def __init__(self, d_model, n_heads, d_keys=None, d_values=None, mix=True, dropout=0.1, eta=1.1, gamma=-0.1,
                 mode='entmax'):
        super().__init__()
        d_keys = d_keys or (d_model // n_heads)
        d_values = d_values or (d_model // n_heads)
        self.inner_attention = Association(scale=None, attention_dropout=dropout, eta=eta, gamma=gamma, mode=mode)
        self.query_projection = nn.Linear(d_model, d_keys * n_heads)
        self.key_projection = nn.Linear(d_model, d_keys * n_heads)
        self.value_projection = nn.Linear(d_keys * n_heads, d_values * n_heads)
        self.out_projection = nn.Linear(d_values * n_heads, d_model)
        self.n_heads = n_heads
        self.mix = mix

    def forward(self, queries, keys, values):
        B, L, _ = queries.shape
        _, S, _ = keys.shape
        H = self.n_heads
        queries = self.query_projection(queries).view(B, L, H, -1)
        k = self.key_projection(keys)
        values = self.value_projection(k).view(B, S, H, -1)
        keys = k.view(B, S, H, -1)
        out = self.inner_attention(queries, keys, values)
        if self.mix:
            out = out.transpose(2, 1).contiguous()
        out = out.view(B, L, -1)
        return self.out_projection(out)


class HopfieldPooling(nn.Module):
    """hopfield.py:92-146: learned pattern queries... in the reference the learned ``key`` parameter plays the stored
    patterns and the input the queries."""

    def __init__(self, d_model, n_heads, num_pattern=1, d_keys=None, d_values=None, mix=True, dropout=0.1, eta=1.1,
                 gamma=-0.1, mode='entmax'):
        super().__init__()
        d_keys = d_keys or (d_model // n_heads)
        d_values = d_values or (d_model // n_heads)
        self.inner_attention = Association(scale=None, attention_dropout=dropout, eta=eta, gamma=gamma, mode=mode)
        self.query_projection = nn.Linear(d_model, d_keys * n_heads)
        self.key_projection = nn.Linear(d_model, d_keys * n_heads)
        self.value_projection = nn.Linear(d_keys * n_heads, d_values * n_heads)
        self.out_projection = nn.Linear(d_values * n_heads, d_model)
        self.n_heads = n_heads
        self.mix = mix
        self.key = nn.Parameter(torch.empty(size=(1, num_pattern, d_model)), requires_grad=True)
        nn.init.normal_(self.key, mean=0.0, std=0.02)

    def forward(self, query):
        B, L, _ = query.shape
        _, S, _ = self.key.shape
        H = self.n_heads
        k = self.key.repeat(B, 1, 1)
        queries = self.query_projection(query).view(B, L, H, -1)
        keys = self.key_projection(k)
        values = self.value_projection(keys).view(B, S, H, -1)
        keys = keys.view(B, S, H, -1)
        out = self.inner_attention(queries, keys, values)
        if self.mix:
            out = out.transpose(2, 1).contiguous()
        out = out.view(B, L, -1)
        return self.out_projection(out)
