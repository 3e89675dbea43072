"""CPU baseline of the model-level workload (TEST / BENCH INFRASTRUCTURE ONLY).

The reference's own path for BASELINE config 1 is: HF BERT-base (fp32, CPU) whose self-attention modules
are the reference's BertSelfAttentionWithExtras.  /root/reference does not travel to the GPU box, so this
file re-states that module on top of the C oracle (oracle/oeh_oracle.c, OpenMP): Q/K/V projections by
torch.nn.Linear as in the reference (bert_attention.py:183,205-208), the attention core
(bert_attention.py:222-337) by oeh_ref_attention.  Used by bench.py's cpu_baseline and --impl reference.
"""
from __future__ import annotations

import time

import numpy as np
import torch
from torch import nn

from . import c_oracle as CO
from .oeh_oracle import SOFTMAX_TABLE


class OracleBertSelfAttention(nn.Module):
    """BertSelfAttentionWithExtras.forward (bert_attention.py:169-343), eval, self-attention, CPU fp32."""

    def __init__(self, hf_self_attn, attn_softmax: str):
        super().__init__()
        self.query, self.key, self.value = hf_self_attn.query, hf_self_attn.key, hf_self_attn.value
        self.num_attention_heads = hf_self_attn.num_attention_heads
        self.attention_head_size = hf_self_attn.attention_head_size
        self.kind, self.gamma, self.eta = SOFTMAX_TABLE[attn_softmax]

    def forward(self, hidden_states, attention_mask=None, **kw):
        B, T, _ = hidden_states.shape
        H, d = self.num_attention_heads, self.attention_head_size

        def heads(x):   # (B,T,H*d) -> (B,H,T,d) strided numpy view (transpose_for_scores, :164-167)
            return x.detach().view(B, T, H, d).permute(0, 2, 1, 3).numpy()

        q, k, v = heads(self.query(hidden_states)), heads(self.key(hidden_states)), heads(self.value(hidden_states))
        mask = None if attention_mask is None else attention_mask.detach().numpy().astype(np.float32)
        ctx = CO.attention(q, k, v, scale=1.0 / np.sqrt(d), kind=self.kind, gamma=self.gamma, eta=self.eta,
                           mask=mask, out_bthd=True)
        return torch.from_numpy(ctx), None


def build_bert_base_cpu(attn_softmax: str, seed: int = 3000, num_hidden_layers: int = 12):
    from transformers import BertConfig, BertForMaskedLM
    cfg = BertConfig(num_hidden_layers=num_hidden_layers)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = BertForMaskedLM(cfg).eval()
    for layer in model.bert.encoder.layer:
        layer.attention.self = OracleBertSelfAttention(layer.attention.self, attn_softmax)
    return model


def time_bert_eval_cpu(attn_softmax: str, batch: int, seq: int, steps: int, warmup: int, seed: int = 3000,
                       threads: int | None = None):
    """tokens/s of forward+loss on the host cores.  Returns (tokens_per_s, ms_per_step, threads, losses)."""
    import os
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    CO.set_threads(threads)            # OpenMP side too: torchrun exports OMP_NUM_THREADS=1
    model = build_bert_base_cpu(attn_softmax, seed)
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 30522, (batch, seq), generator=g)
    masked = torch.rand(batch, seq, generator=g) < 0.15
    masked[:, 0] = True
    labels = torch.where(masked, ids, torch.full_like(ids, -100))
    losses = []
    with torch.no_grad():
        for _ in range(warmup):
            model(input_ids=ids, labels=labels)
        t0 = time.perf_counter()
        for _ in range(steps):
            losses.append(float(model(input_ids=ids, labels=labels).loss))
        dt = time.perf_counter() - t0
    return batch * seq * steps / dt, dt / steps * 1e3, threads, losses
