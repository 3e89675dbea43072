"""``BertSelfAttentionWithExtras`` -- drop-in for the reference module
(transformers_language/models/bert_attention.py:29-343) whose attention core
``QK^T -> /sqrt(d) -> +mask -> softmax_fn -> PV -> gate -> permute`` (:222-337) runs as ONE fused
tcgen05 kernel (csrc/attn_fwd_sm100.cu) on strided views of the Linear outputs -- no S/P tensor, no
permute copy.  Same constructor kwargs, attributes, submodule names and state_dict keys.

Paths (all on the GPU, all through liboeh.so; there is no CPU fallback):
  fused    fp16/bf16 activations, head dim 64, a recognised softmax_fn, eval mode        -> oeh_attn_fwd
  unfused  fp32 activations, output_attentions=True, head_mask, relative position embeddings, other
           head dims or an unknown softmax callable: library GEMMs + oeh_softmax_rows + oeh_gate_logits.
           S and P are materialised, so the three nn.Identity probes (:78-81) see real data there.
Training / backward is out of scope: calling with autograd enabled on tensors that
require grad raises.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import ops
from .quantization import BaseEnumOptions
from .softmax import SoftmaxFn, as_softmax_fn, clipped_softmax, KIND_CLIPPED


def logit(p, eps=1e-16):
    """bert_attention.py:16-18."""
    p = np.clip(p, eps, 1 - eps)
    return -np.log(1 / p - 1)


class AttentionGateType(BaseEnumOptions):
    """bert_attention.py:21-25."""
    none = 0
    unconditional_per_head = 1
    conditional_per_head = 2
    conditional_per_token = 3


FUSED_HEAD_DIM = 64


def _fusable_dtype(module, q: torch.Tensor) -> bool:
    """16-bit activations always take the fused kernel.  fp32 activations (the reference's eval precision,
    accelerate_configs/1gpu_no_mp.yaml:14) take it only when the module's ``fused_fp32`` flag is set: the kernel then rounds
    Q / K / V to fp16 operands (fp32 accumulation and softmax), i.e. it has the fp16 path's 2e-3 accuracy, not an fp32 GEMM's
    -- the default keeps fp32 on the exact, unfused library path."""
    return q.dtype in (torch.float16, torch.bfloat16) or (q.dtype == torch.float32 and bool(getattr(module, "fused_fp32", False)))


class GateMixin:
    """Gate construction (bert_attention.py:98-162; opt_attention.py:81-144) and evaluation
    (bert_attention.py:294-331; opt_attention.py:276-313) shared by the BERT and OPT modules.
    All gate flavours are reduced to one (B,H,T) fp32 *logit* tensor; the sigmoid and the multiply by
    ``gate_scaling_factor`` happen in the attention epilogue."""

    def _init_gate(self, num_heads, head_dim, all_head_size, attn_gate_type, attn_gate_init, attn_gate_mlp,
                   attn_gate_mlp2, attn_gate_linear_all_features, fine_tuning, fine_tuning_std):
        self.last_gate_avg_prob_ = None
        self._last_gate_logits = None
        self.attn_gate_type = attn_gate_type
        self.attn_gate_init = attn_gate_init
        self.attn_gate_mlp = attn_gate_mlp
        self.attn_gate_mlp2 = attn_gate_mlp2
        self.attn_gate_linear_all_features = attn_gate_linear_all_features
        self.alpha = None
        self.gate_fn = torch.sigmoid
        self.pooling_fn = partial(torch.mean, dim=1, keepdims=True)
        self.fine_tuning = fine_tuning
        self.gate_scaling_factor = 1.0
        if self.fine_tuning and self.attn_gate_init is not None:
            self.gate_scaling_factor = 1.0 / self.attn_gate_init
        self._gate_pack = None   # (key, packed fp32 weights) cache for the gate kernel

        if self.attn_gate_type == AttentionGateType.unconditional_per_head:
            self.alpha = nn.Parameter(torch.zeros(size=(num_heads,)), requires_grad=True)
        elif self.attn_gate_type in (AttentionGateType.conditional_per_head, AttentionGateType.conditional_per_token):
            if self.attn_gate_linear_all_features:
                self.alpha = nn.Linear(all_head_size, num_heads, bias=True)
            else:
                module_list = []
                for _ in range(num_heads):
                    if self.attn_gate_mlp:
                        fc = nn.Sequential(nn.Linear(head_dim, head_dim // 4, bias=True), nn.ReLU(),
                                           nn.Linear(head_dim // 4, 1, bias=True))
                    elif self.attn_gate_mlp2:
                        fc = nn.Sequential(nn.Linear(head_dim, head_dim, bias=True), nn.ReLU(),
                                           nn.Linear(head_dim, 1, bias=True))
                    else:
                        fc = nn.Linear(head_dim, 1, bias=True)
                        if self.attn_gate_init is not None:
                            torch.nn.init.constant_(fc.bias, logit(self.attn_gate_init))
                        if self.fine_tuning:
                            torch.nn.init.normal_(fc.weight, mean=0.0, std=fine_tuning_std)
                    module_list.append(fc)
                self.alpha = nn.ModuleList(module_list)

    @property
    def effective_gate_scale(self) -> float:
        """``gate_scaling_factor`` multiplies the CONDITIONAL gates only (bert_attention.py:327, opt_attention.py:309);
        the unconditional per-head branch is ``context_layer *= gate.view(-1, 1, 1)`` with no factor (:295-297 / :277-279)."""
        if self.attn_gate_type == AttentionGateType.unconditional_per_head:
            return 1.0
        return self.gate_scaling_factor

    def _probes_hooked(self) -> bool:
        """A forward hook sits on one of the S / P probe Identities (attach_act_hooks hooks every named module): the
        fused kernel never materialises S / P, so such a call takes the unfused path where the probes see real data."""
        return any(m._forward_hooks or m._forward_pre_hooks
                   for m in (self.attn_scores, self.attn_probs_before_dropout, self.attn_probs_after_dropout))

    # gate statistics the reference stores eagerly; computed on demand here (read only by training TB code)
    @property
    def last_gate_all_probs(self):
        if self._last_gate_logits is None:
            return None
        return torch.sigmoid(self._last_gate_logits).unsqueeze(-1)      # (B,H,T,1)

    @last_gate_all_probs.setter
    def last_gate_all_probs(self, v):
        self._last_gate_logits = None if v is None else v

    @property
    def last_gate_avg_prob(self):
        if self._last_gate_logits is None:
            return self.last_gate_avg_prob_
        g = torch.sigmoid(self._last_gate_logits)                       # (B,H,T)
        return g.mean(dim=0).mean(dim=1)

    @last_gate_avg_prob.setter
    def last_gate_avg_prob(self, v):
        self.last_gate_avg_prob_ = v

    def _packed_gate_weights(self, device):
        """Per-head gate predictors packed once into contiguous fp32 tensors for oeh_gate_logits."""
        params = list(self.alpha.parameters())
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if self._gate_pack is not None and self._gate_pack[0] == key:
            return self._gate_pack[1]
        with torch.no_grad():
            if isinstance(self.alpha[0], nn.Linear):
                w1 = torch.stack([fc.weight[0] for fc in self.alpha]).float().contiguous().to(device)
                b1 = torch.stack([fc.bias[0] for fc in self.alpha]).float().contiguous().to(device)
                pack = (w1, b1, None, None)
            else:
                w1 = torch.stack([fc[0].weight for fc in self.alpha]).float().contiguous().to(device)
                b1 = torch.stack([fc[0].bias for fc in self.alpha]).float().contiguous().to(device)
                w2 = torch.stack([fc[2].weight[0] for fc in self.alpha]).float().contiguous().to(device)
                b2 = torch.stack([fc[2].bias[0] for fc in self.alpha]).float().contiguous().to(device)
                pack = (w1, b1, w2, b2)
        self._gate_pack = (key, pack)
        return pack

    def fusable_gate(self, device):
        """(w (H,64), b (H)) fp32 of Linear per-token gate predictors, for a producer kernel that can emit the gate logits
        together with the hidden states it writes (``ops.add_layer_norm(..., gate=...)``); None for every other gate flavour."""
        if (self.attn_gate_type != AttentionGateType.conditional_per_token or self.attn_gate_linear_all_features
                or not isinstance(self.alpha, nn.ModuleList) or not isinstance(self.alpha[0], nn.Linear)
                or self.alpha[0].in_features != FUSED_HEAD_DIM):
            return None
        w1, b1, _, _ = self._packed_gate_weights(device)
        return w1, b1

    def _gate_logits(self, hidden_states: torch.Tensor, num_heads: int, head_dim: int) -> Optional[torch.Tensor]:
        """(B,H,T) fp32 gate logits, or None when the module has no gate."""
        gt = self.attn_gate_type
        if gt == AttentionGateType.none or gt is None:
            return None
        B, T, _ = hidden_states.shape
        if gt == AttentionGateType.unconditional_per_head:
            lg = self.alpha.detach().float().view(1, num_heads, 1).expand(B, num_heads, T).contiguous()
        elif self.attn_gate_linear_all_features:
            lg = self.alpha(hidden_states).float().permute(0, 2, 1).contiguous()          # (B,H,T)
        elif head_dim == FUSED_HEAD_DIM:
            w1, b1, w2, b2 = self._packed_gate_weights(hidden_states.device)
            lg = ops.gate_logits(hidden_states, num_heads, w1, b1, w2, b2)
        else:   # head dims the gate kernel is not built for: the reference's per-head loop
            x = hidden_states.view(B, T, num_heads, head_dim).permute(0, 2, 1, 3)
            lg = torch.stack([self.alpha[h](x[:, h]).squeeze(-1) for h in range(num_heads)], dim=1).float().contiguous()
        if gt == AttentionGateType.conditional_per_head and not self.attn_gate_linear_all_features:
            lg = lg.mean(dim=2, keepdim=True).expand(B, num_heads, T).contiguous()       # pooling_fn over tokens
        self._last_gate_logits = lg
        return lg


def _check_no_grad(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "outeffhop_b200 implements the evaluation forward only; wrap the call in torch.no_grad() "
            "(training / backward is outside the hot-path scope, SURVEY section 2 #15)")


def fused_qkv_projection(owner: nn.Module, mods, hidden_states: torch.Tensor, num_heads: int, head_dim: int):
    """Plain self-attention in eval: ONE (B*T, D) x (D, 3D) library GEMM for the three projections instead of three (a
    768-wide output is < 1 wave of GEMM tiles on 148 SMs; 2304 wide fills them).  The concatenated weight is a cache
    keyed on the three Linear parameters (the q/k/v Linear modules stay the modules of record: state_dict, hooks,
    quantized wrappers); q, k, v come back as (B,H,T,d) strided views of the (B,T,3,H,d) result, which the fused
    kernel's tensor maps take as they are.  Returns None whenever the three modules must run themselves."""
    if owner.training or any(type(m) is not nn.Linear or m._forward_hooks or m._forward_pre_hooks for m in mods) \
            or any((m.bias is None) != (mods[0].bias is None) for m in mods):
        return None
    key = tuple((m.weight.data_ptr(), m.weight._version, 0 if m.bias is None else m.bias.data_ptr(),
                 0 if m.bias is None else m.bias._version) for m in mods)
    cache = owner.__dict__.get("_qkv_cache")
    if cache is None or cache[0] != key:
        with torch.no_grad():
            w = torch.cat([m.weight for m in mods], dim=0)
            b = None if mods[0].bias is None else torch.cat([m.bias for m in mods], dim=0)
        cache = (key, w, b)
        owner.__dict__["_qkv_cache"] = cache
    y = nn.functional.linear(hidden_states, cache[1], cache[2])
    B, T = y.shape[0], y.shape[1]
    y = y.view(B, T, 3, num_heads, head_dim)
    return tuple(y[:, :, j].permute(0, 2, 1, 3) for j in range(3))


class BertSelfAttentionWithExtras(GateMixin, nn.Module):
    def __init__(self, config, position_embedding_type=None, softmax_fn=torch.nn.functional.softmax, alpha=None,
                 ssm_eps=None, tau=None, max_seq_length=None, skip_attn=False,
                 attn_gate_type=AttentionGateType.none, attn_gate_init=None, attn_gate_mlp=False,
                 attn_gate_mlp2=False, attn_gate_linear_all_features=False, fine_tuning=False):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0 and not hasattr(config, "embedding_size"):
            raise ValueError(
                f"The hidden size ({config.hidden_size}) is not a multiple of the number of attention "
                f"heads ({config.num_attention_heads})")
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size

        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)

        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)
        self.position_embedding_type = position_embedding_type or getattr(config, "position_embedding_type", "absolute")
        if self.position_embedding_type in ("relative_key", "relative_key_query"):
            self.max_position_embeddings = config.max_position_embeddings
            self.distance_embedding = nn.Embedding(2 * config.max_position_embeddings - 1, self.attention_head_size)
        self.is_decoder = config.is_decoder

        # names matter: the validate scripts attach forward hooks to every named module (:78-81)
        self.attn_scores = nn.Identity()
        self.attn_probs_before_dropout = nn.Identity()
        self.attn_probs_after_dropout = nn.Identity()

        self.alpha = alpha
        self.ssm_eps = ssm_eps
        self.tau = tau
        self.max_seq_length = max_seq_length
        if self.alpha is not None:                                   # --alpha override (:88-94)
            assert self.max_seq_length is not None
            gamma = -self.alpha / self.max_seq_length
            self.softmax_fn = SoftmaxFn(KIND_CLIPPED, gamma, 1.0)
        else:
            self.softmax_fn = softmax_fn
        self.skip_attn = skip_attn
        self.materialize_probs = False    # force the unfused path (e.g. to feed the Identity probes)
        self.fused_fp32 = False           # fp32 activations on the fused kernel with fp16 operands (see _fusable_dtype)
        self._init_gate(self.num_attention_heads, self.attention_head_size, self.all_head_size, attn_gate_type,
                        attn_gate_init, attn_gate_mlp, attn_gate_mlp2, attn_gate_linear_all_features, fine_tuning,
                        fine_tuning_std=0.01)

    # hooks overridden by the quantized wrapper (quantized_attention.py)
    def _taps(self):
        return None

    def _gate_scale(self) -> float:
        return self.effective_gate_scale

    def _fused_qkv(self, hidden_states: torch.Tensor):
        return fused_qkv_projection(self, (self.query, self.key, self.value), hidden_states, self.num_attention_heads,
                                    self.attention_head_size)

    def transpose_for_scores(self, x: torch.Tensor) -> torch.Tensor:
        new_x_shape = x.size()[:-1] + (self.num_attention_heads, self.attention_head_size)
        return x.view(new_x_shape).permute(0, 2, 1, 3)

    # ------------------------------------------------------------------------------------------------
    def _core(self, q, k, v, attention_mask, gate_logits, gate_scale, head_mask, output_attentions, taps=None):
        """(context (B,T,H*d), probs-or-None).  ``taps`` = optional (scores, probs, ctx) QuantizedActivation
        modules of the quantized wrapper (quantized_bert.py:363,374,434)."""
        fn = as_softmax_fn(self.softmax_fn)
        d = self.attention_head_size
        scale = 1.0 / math.sqrt(d)
        training_dropout = self.training and self.dropout.p > 0
        active = [t for t in (taps or ()) if t is not None and t.active]
        fusable_taps = all(t.fixed for t in active)
        fused = (fn is not None and _fusable_dtype(self, q) and d == FUSED_HEAD_DIM
                 and head_mask is None and not output_attentions and not self.materialize_probs
                 and not self._probes_hooked() and not training_dropout and fusable_taps
                 and self.position_embedding_type not in ("relative_key", "relative_key_query"))
        if fused:
            kw = {}
            if active:
                ts, tp, tc = taps
                kw = dict(qdq_scores=ts.kernel_qparams() if ts.active else None,
                          qdq_probs=tp.kernel_qparams() if tp.active else None,
                          qdq_ctx=tc.kernel_qparams() if tc.active else None,
                          n_bits=active[0].activation_quantizer.n_bits, ctx_qdq_before_gate=False)
            ctx = ops.attention_fwd(q, k, v, scale=scale, kind=fn.kind, gamma=fn.gamma, eta=fn.eta,
                                    mask=attention_mask, gate_logits=gate_logits, gate_scale=gate_scale,
                                    out_layout="bthd", **kw)
            return ctx, None
        return self._core_unfused(q, k, v, attention_mask, gate_logits, gate_scale, head_mask, taps)

    def _core_unfused(self, q, k, v, attention_mask, gate_logits, gate_scale, head_mask, taps):
        """The reference's op sequence with S/P materialised: library GEMMs + the CUDA softmax kernel."""
        ts, tp, tc = taps if taps is not None else (None, None, None)
        fn = as_softmax_fn(self.softmax_fn)
        attention_scores = torch.matmul(q, k.transpose(-1, -2))                               # :222
        if self.position_embedding_type in ("relative_key", "relative_key_query"):           # :224-263
            attention_scores = attention_scores + self._relative_scores(q, k)
        attention_scores = attention_scores / math.sqrt(self.attention_head_size)             # :265
        attention_scores = self.attn_scores(attention_scores)                                 # :268 (probe)
        if ts is not None:
            attention_scores = ts(attention_scores)                                           # quantized_bert.py:363
        if fn is not None:
            probs = ops.softmax_rows(attention_scores, fn.kind, fn.gamma, fn.eta, mask=attention_mask)   # :270-276
        else:
            if attention_mask is not None:
                attention_scores = attention_scores + attention_mask
            probs = self.softmax_fn(attention_scores, dim=-1)
        probs = self.attn_probs_before_dropout(probs)                                         # :279
        if tp is not None:
            probs = tp(probs)                                                                 # quantized_bert.py:374
        probs = self.dropout(probs)                                                           # :283
        probs = self.attn_probs_after_dropout(probs)                                          # :286
        if head_mask is not None:
            probs = probs * head_mask                                                         # :289-290
        ctx = torch.matmul(probs, v)                                                          # :292
        if gate_logits is not None:                                                           # :294-331
            ctx = ctx * (torch.sigmoid(gate_logits).to(ctx.dtype).unsqueeze(-1) * gate_scale)
        ctx = ctx.permute(0, 2, 1, 3).contiguous()                                            # :335
        ctx = ctx.view(ctx.size()[:-2] + (self.all_head_size,))                               # :336-337
        if tc is not None:
            ctx = tc(ctx)                                                                     # quantized_bert.py:434
        return ctx, probs

    def _relative_scores(self, query_layer, key_layer):
        query_length, key_length = query_layer.shape[2], key_layer.shape[2]
        dev = query_layer.device
        position_ids_l = torch.arange(query_length, dtype=torch.long, device=dev).view(-1, 1)
        position_ids_r = torch.arange(key_length, dtype=torch.long, device=dev).view(1, -1)
        distance = position_ids_l - position_ids_r
        pe = self.distance_embedding(distance + self.max_position_embeddings - 1).to(dtype=query_layer.dtype)
        rel = torch.einsum("bhld,lrd->bhlr", query_layer, pe)
        if self.position_embedding_type == "relative_key_query":
            rel = rel + torch.einsum("bhrd,lrd->bhlr", key_layer, pe)
        return rel

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.FloatTensor] = None,
                head_mask: Optional[torch.FloatTensor] = None,
                encoder_hidden_states: Optional[torch.FloatTensor] = None,
                encoder_attention_mask: Optional[torch.FloatTensor] = None,
                past_key_value: Optional[Tuple[Tuple[torch.FloatTensor]]] = None,
                output_attentions: Optional[bool] = False, **hf_kwargs) -> Tuple[torch.Tensor]:
        """Same contract as the reference (:169-343).  Extra keyword arguments of newer HF call conventions
        (``past_key_values=None``, ``cache_position`` ...) are accepted and ignored when None."""
        if hf_kwargs.get("past_key_values") is not None:
            raise NotImplementedError("HF>=4.36 Cache objects are not supported; pass past_key_value tuples")
        if self.skip_attn:
            return (torch.zeros_like(hidden_states),)                                          # :179-181
        if not hidden_states.is_cuda:
            raise RuntimeError("outeffhop_b200.BertSelfAttentionWithExtras runs on CUDA only (no CPU fallback)")
        _check_no_grad(hidden_states, self.query.weight)

        is_cross_attention = encoder_hidden_states is not None
        fused_qkv = self._fused_qkv(hidden_states) if not is_cross_attention and past_key_value is None else None
        if fused_qkv is not None:
            q, k, v = fused_qkv
        else:
            q = self.transpose_for_scores(self.query(hidden_states))                           # :183, :220
        if fused_qkv is not None:
            pass
        elif is_cross_attention and past_key_value is not None:                                # :190-194
            k, v = past_key_value[0], past_key_value[1]
            attention_mask = encoder_attention_mask
        elif is_cross_attention:                                                               # :195-198
            k = self.transpose_for_scores(self.key(encoder_hidden_states))
            v = self.transpose_for_scores(self.value(encoder_hidden_states))
            attention_mask = encoder_attention_mask
        elif past_key_value is not None:                                                       # :199-203
            k = torch.cat([past_key_value[0], self.transpose_for_scores(self.key(hidden_states))], dim=2)
            v = torch.cat([past_key_value[1], self.transpose_for_scores(self.value(hidden_states))], dim=2)
        else:                                                                                  # :204-206
            k = self.transpose_for_scores(self.key(hidden_states))
            v = self.transpose_for_scores(self.value(hidden_states))
        if self.is_decoder:
            past_key_value = (k, v)                                                            # :219

        gate_logits = self._gate_logits(hidden_states, self.num_attention_heads, self.attention_head_size)
        ctx, probs = self._core(q, k, v, attention_mask, gate_logits, self._gate_scale(), head_mask,
                                output_attentions, taps=self._taps())
        outputs = (ctx, probs) if output_attentions else (ctx,)
        if self.is_decoder:
            outputs = outputs + (past_key_value,)
        return outputs
