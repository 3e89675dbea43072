"""Synthetic evaluation harness: the eval loops of validate_mlm_config.py:489-548 and validate_clm.py:555-594
re-stated without datasets / Accelerate: random-init HF model of the named architecture,
attention modules swapped for the drop-ins, synthetic token batches, forward + loss per batch, perplexity =
exp(mean of per-batch mean losses).

One process per GPU.  Batches are sharded round-robin over ranks (rank r evaluates batches r, r+N, ...), there is no
data-path collective; the per-rank (sum of batch losses, #batches) pair is all-reduced once at the end
(replaces the per-step ``gather_for_metrics``, validate_mlm_config.py:496-498).  The eval step is captured in a
CUDA graph: at B=32,T=128 a BERT-base forward is ~1 ms of GPU work and Python launch overhead would dominate.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from .hf_compat import (fuse_bert_gelu, fuse_bert_residual_blocks, fuse_bert_shell, fuse_opt_decoder_layers,
                        swap_bert_attention, swap_opt_attention)


# ---- model / data builders ---------------------------------------------------------------------------------

def build_bert_base(seed: int = 3000, dtype=torch.float16, device="cuda", attn_softmax="clippedsoftmax1(-.025:1)",
                    num_hidden_layers: int = 12, fuse_residual: bool = True, **attn_kw):
    """BertConfig() defaults = BERT-base (12L/768/12H/3072, vocab 30522), random init (Cfg1/Cfg2).
    ``fuse_residual``: also swap BertSelfOutput / BertOutput for the fused add + LayerNorm modules and the FFN
    activation for the library's erf GELU (in the FFN1 GEMM epilogue), and the embeddings / MLM-head transform for their
    one-kernel forms."""
    from transformers import BertConfig, BertForMaskedLM
    cfg = BertConfig(num_hidden_layers=num_hidden_layers)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = BertForMaskedLM(cfg)
    model = model.to(device=device, dtype=dtype).eval()
    model = swap_bert_attention(model, attn_softmax=attn_softmax, max_seq_length=128, **attn_kw)
    return fuse_bert_shell(fuse_bert_gelu(fuse_bert_residual_blocks(model))) if fuse_residual else model


def build_opt_125m(seed: int = 5678, dtype=torch.float16, device="cuda", attn_softmax="vanilla",
                   num_hidden_layers: int = 12, causal_hint: bool = True, fuse_residual: bool = True, **attn_kw):
    """OPT-125m shape (Cfg3/Cfg4), random init.  ``fuse_residual``: also swap the decoder layers for the fused eval
    modules (hf_compat.OPTDecoderLayerFused)."""
    from transformers import OPTConfig, OPTForCausalLM
    cfg = OPTConfig(hidden_size=768, num_hidden_layers=num_hidden_layers, ffn_dim=3072, num_attention_heads=12,
                    word_embed_proj_dim=768, max_position_embeddings=2048, do_layer_norm_before=True,
                    vocab_size=50272)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = OPTForCausalLM(cfg)
    model = model.to(device=device, dtype=dtype).eval()
    model = swap_opt_attention(model, attn_softmax=attn_softmax, max_seq_length=512, causal_hint=causal_hint, **attn_kw)
    return fuse_opt_decoder_layers(model) if fuse_residual else model


def synthetic_mlm_batch(batch: int, seq: int, vocab: int = 30522, seed: int = 0, mlm_probability: float = 0.15,
                        pin: bool = True) -> Dict[str, torch.Tensor]:
    """ids ~ U{0..vocab-1}; labels = ids on a Bernoulli(0.15) mask else -100 (DataCollatorForLanguageModeling)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (batch, seq), generator=g)
    masked = torch.rand(batch, seq, generator=g) < mlm_probability
    masked[:, 0] = True                                   # never an empty label set
    labels = torch.where(masked, ids, torch.full_like(ids, -100))
    out = {"input_ids": ids, "labels": labels}
    return {k: (v.pin_memory() if pin and torch.cuda.is_available() else v) for k, v in out.items()}


def synthetic_clm_batch(batch: int, seq: int, vocab: int = 50272, seed: int = 0, pin: bool = True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (batch, seq), generator=g)
    out = {"input_ids": ids, "labels": ids.clone()}
    return {k: (v.pin_memory() if pin and torch.cuda.is_available() else v) for k, v in out.items()}


# ---- graph-capturable forwards ---------------------------------------------------------------------------------
# HF's top-level forward() builds its attention masks with host-side tensor constructions that cannot be captured in
# a CUDA graph, and (for all-ones masks) with data-dependent checks.  The eval data of the reference is chunked and
# un-padded (validate_mlm_config.py:344-357), so the BERT mask is all zeros and the OPT mask is the plain causal one:
# these forwards call the same HF submodules in the same order with that knowledge baked in.

def _padded_decoder(linear: torch.nn.Linear, multiple: int = 64):
    """(weight, bias, vocab) of an output projection with the vocabulary dimension zero-padded to a multiple of 64 rows.
    BERT's 30522-wide fp16 decoder gives logits rows that are only 4-byte aligned, which drops cuBLAS onto an
    `align2` sm80 kernel (736 us of a 3.1 ms step, profiles/r01_launches_bert_step.summary.txt); 30528 keeps the fast
    path.  The pad columns are never read by the loss kernel (``vocab`` < row stride).  Cached on the module; eval only."""
    V = linear.out_features
    Vp = (V + multiple - 1) // multiple * multiple
    if Vp == V:
        return linear.weight, linear.bias, V
    key = (linear.weight.data_ptr(), linear.weight._version, None if linear.bias is None else linear.bias.data_ptr())
    cached = getattr(linear, "_oeh_padded", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            w = linear.weight.new_zeros((Vp, linear.in_features))
            w[:V].copy_(linear.weight)
            b = None
            if linear.bias is not None:
                b = linear.bias.new_zeros((Vp,))
                b[:V].copy_(linear.bias)
        cached = (key, w, b)
        linear._oeh_padded = cached
    return cached[1], cached[2], V


def bert_mlm_logits(model, input_ids: torch.Tensor):
    """embeddings -> 12 x BertLayer (attention_mask=None) -> MLM head.  Returns ((B*T, V_padded) logits, V)."""
    bert = model.bert
    h = bert.embeddings(input_ids=input_ids)
    for layer in bert.encoder.layer:
        out = layer(h, attention_mask=None)
        h = out[0] if isinstance(out, tuple) else out
    pred = model.cls.predictions
    w, b, V = _padded_decoder(pred.decoder)
    logits = torch.nn.functional.linear(pred.transform(h), w, b)
    return logits.view(-1, logits.shape[-1]), V


def bert_mlm_loss(model, input_ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """BertForMaskedLM.forward for un-padded batches: mean CE over the labelled tokens, evaluated in fp32 by
    ``oeh_ce_loss`` directly on the fp16 logits (a fp16 mean would round the loss to 3 decimal digits; ATen needs a
    1 GB/step fp32 copy + log-softmax for the same number)."""
    from . import ops
    logits, V = bert_mlm_logits(model, input_ids)
    return ops.ce_loss(logits, labels.reshape(-1), vocab=V)


def bert_mlm_loss_aten(model, input_ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """The same loss through ATen only (unpadded decoder GEMM, fp32 copy, F.cross_entropy): the check for
    ``bert_mlm_loss`` in tests, and the pre-fusion baseline of bench.py --aten-loss."""
    bert = model.bert
    h = bert.embeddings(input_ids=input_ids)
    for layer in bert.encoder.layer:
        out = layer(h, attention_mask=None)
        h = out[0] if isinstance(out, tuple) else out
    logits = model.cls(h)
    return torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]).float(), labels.view(-1))


def opt_clm_loss(model, input_ids: torch.Tensor, labels: torch.Tensor, causal_mask: torch.Tensor,
                 position_ids: torch.Tensor, aten_loss: bool = False) -> torch.Tensor:
    """OPTForCausalLM.forward for un-padded blocks: token + learned position embeddings -> 12 x OPTDecoderLayer with the
    (B,1,T,T) causal mask -> final LayerNorm -> lm_head -> shifted mean CE in fp32."""
    dec = model.model.decoder
    tok, posm = dec.embed_tokens, dec.embed_positions
    if (type(tok) is torch.nn.Embedding and tok.max_norm is None and dec.project_in is None and position_ids is not None
            and tok.weight.is_cuda and tok.weight.dtype == posm.weight.dtype and not model.training
            and not tok._forward_hooks and not posm._forward_hooks and input_ids.dim() == 2):
        # token + learned position embeddings in ONE launch (two gathers, the int64 offset add and the sum in ATen)
        from . import ops
        h = ops.embed_sum(input_ids, tok.weight, posm.weight, position_ids, getattr(posm, "offset", 2))
    else:
        h = tok(input_ids)
        pos = posm(None, 0, position_ids=position_ids)
        if dec.project_in is not None:
            h = dec.project_in(h)
        h = h + pos
    from .hf_compat import OPTDecoderLayerFused
    layers = list(dec.layers)
    if (layers and dec.final_layer_norm is not None and dec.final_layer_norm.bias is not None
            and all(isinstance(l, OPTDecoderLayerFused) and l._fusable(h) for l in layers)):
        # chained fused layers: every residual add shares a kernel with the LayerNorm that follows it, across layer
        # boundaries too (the last one is the decoder's final_layer_norm)
        from . import ops
        ln = layers[0].self_attn_layer_norm
        gates = [l.gate_predictors(h.device) if h.dim() == 3 else None for l in layers] + [None]
        x = ops.add_layer_norm(h, None, ln.weight, ln.bias, ln.eps, gate=gates[0])
        logits = None
        if gates[0] is not None:
            x, logits = x
        for i, layer in enumerate(layers):
            next_ln = layers[i + 1].self_attn_layer_norm if i + 1 < len(layers) else dec.final_layer_norm
            # the kernel that writes the next layer's input also emits that layer's per-token gate logits
            out = layer.forward_chained(x, h, next_ln, attention_mask=causal_mask, position_ids=position_ids,
                                        gate_logits=logits, next_gate=gates[i + 1])
            x, h = out[0], out[1]
            logits = out[2] if gates[i + 1] is not None else None
        h = x
    else:
        for layer in layers:
            out = layer(h, attention_mask=causal_mask, position_ids=position_ids, past_key_values=None, use_cache=False)
            h = out[0] if isinstance(out, tuple) else out
        if dec.final_layer_norm is not None:
            h = dec.final_layer_norm(h)
    if dec.project_out is not None:
        h = dec.project_out(h)
    w, b, V = _padded_decoder(model.lm_head)
    logits = torch.nn.functional.linear(h, w, b)
    if aten_loss:
        return torch.nn.functional.cross_entropy(logits[:, :-1, :V].reshape(-1, V).float(), labels[:, 1:].reshape(-1))
    # shifted CE without slicing / copying the logits: row (b, t) is scored against labels[b, t+1], the last row
    # of every block is ignored (same set of B*(T-1) terms as HF's shift, validate_clm.py:555-561)
    # (the shift is applied to the label index inside the kernel: no shifted copy of the labels either)
    from . import ops
    return ops.ce_loss(logits.view(-1, logits.shape[-1]), labels.reshape(-1), vocab=V, shift_block=labels.shape[-1])


def make_causal_mask(batch: int, seq: int, dtype, device) -> torch.Tensor:
    """HF-4.31 ``_make_causal_mask``: (B,1,T,T), 0 on/below the diagonal, finfo(dtype).min above."""
    m = torch.full((seq, seq), torch.finfo(dtype).min, dtype=dtype, device=device).triu(1)
    return m[None, None].expand(batch, 1, seq, seq).contiguous()


# ---- eval step ----------------------------------------------------------------------------------------------

@dataclass
class EvalStep:
    """One captured eval step: static device inputs -> static loss (fp32 scalar on the device)."""
    loss_fn: Callable[..., torch.Tensor]
    static_inputs: Dict[str, torch.Tensor]
    loss: torch.Tensor
    graph: Optional[torch.cuda.CUDAGraph]

    def load(self, host_batch: Dict[str, torch.Tensor]):
        """H2D of one batch into the static buffers (async on the current stream; pinned host memory)."""
        for k, v in host_batch.items():
            self.static_inputs[k].copy_(v, non_blocking=True)

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self.loss.copy_(self.loss_fn(**self.static_inputs))
        return self.loss

    def h2d_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.static_inputs.values())


def make_eval_step(loss_fn: Callable[..., torch.Tensor], example_batch: Dict[str, torch.Tensor], device="cuda",
                   use_graph: bool = True, warmup: int = 3) -> EvalStep:
    """``loss_fn(**batch) -> scalar loss`` is warmed up on a side stream, then captured into one CUDA graph."""
    static = {k: v.to(device) for k, v in example_batch.items()}
    loss = torch.zeros((), dtype=torch.float32, device=device)
    step = EvalStep(loss_fn, static, loss, None)
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # warm every lazy init (cuBLAS handles, gate packs, attrs)
                loss.copy_(loss_fn(**static))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss.copy_(loss_fn(**static))
            step.graph = g
    return step


def make_bert_eval_step(model, example_batch, use_graph: bool = True) -> EvalStep:
    return make_eval_step(lambda input_ids, labels: bert_mlm_loss(model, input_ids, labels), example_batch,
                          device=next(model.parameters()).device, use_graph=use_graph)


def make_opt_eval_step(model, example_batch, use_graph: bool = True) -> EvalStep:
    p = next(model.parameters())
    B, T = example_batch["input_ids"].shape
    mask = make_causal_mask(B, T, p.dtype, p.device)
    pos = torch.arange(T, device=p.device).unsqueeze(0).expand(B, T).contiguous()
    return make_eval_step(lambda input_ids, labels: opt_clm_loss(model, input_ids, labels, mask, pos), example_batch,
                          device=p.device, use_graph=use_graph)


def aggregate_perplexity(loss_sum: torch.Tensor, n_batches: int) -> float:
    """exp(mean over per-batch mean losses), every batch weighted equally (validate_mlm_config.py:542-548);
    with torch.distributed initialised the (sum, count) pair is all-reduced once."""
    import torch.distributed as dist
    buf = torch.stack((loss_sum.double(), torch.tensor(float(n_batches), dtype=torch.float64, device=loss_sum.device)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    mean = (buf[0] / buf[1]).item()
    try:
        return math.exp(mean)
    except OverflowError:
        return float("inf")


def shard_indices(n_batches: int, rank: int, world: int):
    """Round-robin batch sharding: rank r takes batches r, r+N, ...."""
    return list(range(rank, n_batches, world))


# ---- INT8 (--quantize) evaluation ---------------------------------------------------------------------------------

def make_bert_int8_eval_step(qmodel, example_batch, estimation_batches, use_graph: bool = True) -> EvalStep:
    """validate_mlm_config.py:466-476 then the eval loop: estimate the activation ranges on ``estimation_batches``
    (``pass_data_for_range_estimation``, every tap minmax -> EMA -> qdq on the device), freeze them, and capture the
    fixed-range step -- fused attention with its three taps, one-kernel residual blocks, fused loss -- in a CUDA graph.
    ``qmodel``: ``quantized_models.quantize_bert_for_masked_lm(...)``."""
    from .quantized_models import pass_data_for_range_estimation
    qmodel.estimate_ranges()
    pass_data_for_range_estimation(estimation_batches, qmodel, act_quant=True, max_num_batches=len(estimation_batches))
    qmodel.fix_ranges()
    return make_eval_step(lambda input_ids, labels: qmodel(input_ids=input_ids, labels=labels).loss, example_batch,
                          device=next(qmodel.parameters()).device, use_graph=use_graph)


def make_opt_int8_eval_step(qmodel, example_batch, estimation_batches, use_graph: bool = True) -> EvalStep:
    """validate_clm.py:487-500 then the eval loop, as ``make_bert_int8_eval_step`` (``qmodel``:
    ``quantized_models.quantize_opt_for_causal_lm(...)``; the reference estimates OPT ranges on the TRAIN loader)."""
    from .quantized_models import pass_data_for_range_estimation
    p = next(qmodel.parameters())
    B, T = example_batch["input_ids"].shape
    mask = make_causal_mask(B, T, p.dtype, p.device)
    pos = torch.arange(T, device=p.device).unsqueeze(0).expand(B, T).contiguous()
    qmodel.estimate_ranges()
    pass_data_for_range_estimation(estimation_batches, qmodel, act_quant=True, max_num_batches=len(estimation_batches))
    qmodel.fix_ranges()
    return make_eval_step(lambda input_ids, labels: qmodel(input_ids=input_ids, attention_mask=mask, position_ids=pos,
                                                           labels=labels).loss,
                          example_batch, device=p.device, use_graph=use_graph)
