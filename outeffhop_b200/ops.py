"""Tensor-level wrappers over the C ABI (include/oeh.h).  PyTorch here is plumbing only: device
memory, the current CUDA stream and strides.  Every function enqueues on
``torch.cuda.current_stream()`` and performs no host synchronisation."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

_DT = {torch.float32: L.OEH_F32, torch.float16: L.OEH_F16, torch.bfloat16: L.OEH_BF16}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise NotImplementedError(f"dtype {t.dtype} is not supported by liboeh") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("outeffhop_b200 ops run on CUDA tensors only (there is no CPU fallback)")


_ws_cache = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand (never shrinks, never synchronises)."""
    key = (device.index, _stream())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _bht_strides(t: torch.Tensor):
    if t.dim() != 4 or t.stride(3) != 1:
        raise ValueError("expected a (B,H,T,d) view whose last dimension is contiguous")
    return (C.c_int64 * 3)(t.stride(0), t.stride(1), t.stride(2))


def _mask_args(mask: Optional[torch.Tensor], B: int, H: int, Tq: int, Tk: int):
    if mask is None:
        return None, 0, None, None
    m = mask
    while m.dim() < 4:
        m = m.unsqueeze(0)
    m = m.expand(B, H, Tq, Tk)               # broadcast dims get stride 0
    if Tk > 1 and m.stride(3) != 1:
        m = m.contiguous()
    return m, _dtype_code(m), (C.c_int64 * 3)(m.stride(0), m.stride(1), m.stride(2)), m


def attention_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: float, kind: int,
                  gamma: float = 0.0, eta: float = 1.0, causal: bool = False,
                  mask: Optional[torch.Tensor] = None, clamp_min: bool = False,
                  gate_logits: Optional[torch.Tensor] = None, gate_scale: float = 1.0,
                  qdq_scores: Optional[torch.Tensor] = None, qdq_probs: Optional[torch.Tensor] = None,
                  qdq_ctx: Optional[torch.Tensor] = None, n_bits: int = 8, ctx_qdq_before_gate: bool = False,
                  out_layout: str = "bthd", out: Optional[torch.Tensor] = None, _debug=None) -> torch.Tensor:
    """Fused attention core (oeh_attn_fwd).  q,k,v: (B,H,T,d) strided views, fp16/bf16.
    Returns (B,Tq,H*d) for out_layout="bthd" (the reference's permute+view folded into the store,
    bert_attention.py:335-337 / opt_attention.py:318-322) or (B,H,Tq,d) for "bhtd".
    qdq_* are device tensors [scale, zero_point] (e.g. ``state[4:6]`` of a quantizer state).
    ``out``: optional preallocated contiguous result of exactly that shape and dtype."""
    _require_cuda(q, k, v, mask, gate_logits)
    B, H, Tq, d = q.shape
    Tk = k.shape[2]
    if k.shape != (B, H, Tk, d) or v.shape != (B, H, Tk, d):
        raise ValueError(f"q/k/v shape mismatch: {tuple(q.shape)} {tuple(k.shape)} {tuple(v.shape)}")
    if not (q.dtype == k.dtype == v.dtype):
        raise ValueError("q, k, v must share a dtype")
    if out_layout not in ("bthd", "bhtd"):
        raise ValueError(f"unknown out_layout {out_layout!r}")
    if out is not None:
        want = (B, Tq, H * d) if out_layout == "bthd" else (B, H, Tq, d)
        if tuple(out.shape) != want or out.dtype != q.dtype or out.device != q.device or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous {want} tensor of q's dtype on q's device")
    if out_layout == "bthd":
        base = out.view(B, Tq, H, d) if out is not None else torch.empty((B, Tq, H, d), dtype=q.dtype, device=q.device)
        out_view = base.permute(0, 2, 1, 3)
    else:
        base = out_view = out if out is not None else torch.empty((B, H, Tq, d), dtype=q.dtype, device=q.device)
    a = L.OehAttnArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out_view.data_ptr()
    a.q_strides, a.k_strides, a.v_strides = _bht_strides(q), _bht_strides(k), _bht_strides(v)
    a.out_strides = _bht_strides(out_view)
    a.B, a.H, a.Tq, a.Tk, a.d = B, H, Tq, Tk, d
    a.dtype = _dtype_code(q)
    a.scale = float(scale)
    a.kind, a.gamma, a.eta = int(kind), float(gamma), float(eta)
    a.causal = int(bool(causal))
    m, mdt, mstr, _keep = _mask_args(mask, B, H, Tq, Tk)
    if m is not None:
        a.mask, a.mask_dtype, a.mask_strides = m.data_ptr(), mdt, mstr
    a.clamp_min = int(bool(clamp_min))
    if gate_logits is not None:
        if gate_logits.dtype != torch.float32 or gate_logits.numel() != B * H * Tq or not gate_logits.is_contiguous():
            raise ValueError("gate_logits must be a contiguous fp32 (B,H,Tq) tensor")
        a.gate_logits = gate_logits.data_ptr()
    a.gate_scale = float(gate_scale)
    for name, t in (("qdq_scores", qdq_scores), ("qdq_probs", qdq_probs), ("qdq_ctx", qdq_ctx)):
        if t is not None:
            if t.dtype != torch.float32 or t.numel() < 2 or not t.is_cuda:
                raise ValueError(f"{name} must be a CUDA fp32 tensor [scale, zero_point]")
            setattr(a, name, t.data_ptr())
    a.n_bits = int(n_bits)
    a.ctx_qdq_before_gate = int(bool(ctx_qdq_before_gate))
    a.stream = _stream()
    lib = L.lib()
    need = lib.oeh_attn_workspace_bytes(C.byref(a))
    if need:
        ws = _workspace(q.device, need)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    if _debug is not None:
        dbg_s, dbg_o = _debug
        L.check(lib.oeh_attn_fwd_debug(C.byref(a), dbg_s.data_ptr() if dbg_s is not None else None,
                                       dbg_o.data_ptr() if dbg_o is not None else None))
    else:
        L.check(lib.oeh_attn_fwd(C.byref(a)))
    return base.view(B, Tq, H * d) if out_layout == "bthd" else base


def softmax_rows(x: torch.Tensor, kind: int, gamma: float = 0.0, eta: float = 1.0, *, in_scale: float = 1.0,
                 mask: Optional[torch.Tensor] = None, causal: bool = False, clamp_min: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row softmax family on a materialised (B,H,Tq,Tk) (or (...,Tk)) score tensor (oeh_softmax_rows)."""
    _require_cuda(x, mask)
    if x.dim() == 4:
        B, H, Tq, Tk = x.shape
    else:
        if mask is not None or causal:
            raise ValueError("mask / causal need a 4-D (B,H,Tq,Tk) input")
        Tk = x.shape[-1]
        B, H, Tq = 1, 1, x.numel() // max(Tk, 1)
    xc = x.contiguous()
    y = out if out is not None else torch.empty_like(xc)
    m, mdt, mstr, _keep = _mask_args(mask, B, H, Tq, Tk)
    L.check(L.lib().oeh_softmax_rows(xc.data_ptr(), y.data_ptr(), B, H, Tq, Tk, _dtype_code(xc), float(in_scale),
                                     int(kind), float(gamma), float(eta), m.data_ptr() if m is not None else None,
                                     mdt, mstr, int(bool(causal)), int(bool(clamp_min)), _stream()))
    return y


def gate_logits(x: torch.Tensor, num_heads: int, w1: torch.Tensor, b1: torch.Tensor,
                w2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-token gate logits (oeh_gate_logits).  x (B,T,H*d); Linear gate: w1 (H,d), b1 (H);
    MLP gate: w1 (H,m,d), b1 (H,m), w2 (H,m), b2 (H).  Weights fp32 contiguous.  Returns (B,H,T) fp32."""
    _require_cuda(x, w1, b1, w2, b2)
    B, T, HD = x.shape
    d = HD // num_heads
    if x.stride(2) != 1 or x.stride(0) != T * x.stride(1):
        x = x.contiguous()
    hidden = 0 if w1.dim() == 2 else w1.shape[1]
    for t in (w1, b1, w2, b2):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("gate weights must be contiguous fp32 tensors")
    out = torch.empty((B, num_heads, T), dtype=torch.float32, device=x.device)
    L.check(L.lib().oeh_gate_logits(x.data_ptr(), _dtype_code(x), B, T, num_heads, d, x.stride(1), w1.data_ptr(),
                                    b1.data_ptr(), w2.data_ptr() if w2 is not None else None,
                                    b2.data_ptr() if b2 is not None else None, hidden, out.data_ptr(), _stream()))
    return out


def fake_quant_asym(x: torch.Tensor, qparams: torch.Tensor, n_bits: int = 8, emit_codes: bool = False,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Asymmetric uniform quantize-dequantize (oeh_fake_quant_asym).  qparams: CUDA fp32 [scale, zero_point]."""
    _require_cuda(x, qparams)
    xc = x.contiguous()
    y = out if out is not None else torch.empty_like(xc)
    if qparams.dtype != torch.float32 or qparams.numel() < 2:
        raise ValueError("qparams must be a CUDA fp32 tensor [scale, zero_point]")
    L.check(L.lib().oeh_fake_quant_asym(xc.data_ptr(), y.data_ptr(), xc.numel(), _dtype_code(xc), qparams.data_ptr(),
                                        int(n_bits), int(bool(emit_codes)), _stream()))
    return y


def minmax(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Global [min, max] of x as a 2-element CUDA fp32 tensor (oeh_minmax); no host sync."""
    _require_cuda(x)
    xc = x.contiguous()
    if xc.numel() == 0:
        raise RuntimeError("min(): Expected reduction dim to be specified for input.numel() == 0.")
    res = out if out is not None else torch.empty(2, dtype=torch.float32, device=x.device)
    lib = L.lib()
    need = lib.oeh_minmax_workspace_bytes(xc.numel())
    ws = _workspace(x.device, need)
    L.check(lib.oeh_minmax(xc.data_ptr(), xc.numel(), _dtype_code(xc), res.data_ptr(), ws.data_ptr(), ws.numel(),
                           _stream()))
    return res


def new_qstate(device) -> torch.Tensor:
    """Zeroed device-resident quantizer state (oeh_qstate_slot layout, 8 floats)."""
    return torch.zeros(L.QS_FLOATS, dtype=torch.float32, device=device)


def range_update(state: torch.Tensor, new_minmax: torch.Tensor, momentum: float = 0.9, n_bits: int = 8,
                 eps: float = 1e-8) -> None:
    """EMA + set_quant_range on the device (oeh_range_update)."""
    _require_cuda(state, new_minmax)
    L.check(L.lib().oeh_range_update(state.data_ptr(), new_minmax.data_ptr(), float(momentum), int(n_bits), float(eps),
                                     _stream()))


def set_quant_range_host(x_min: float, x_max: float, n_bits: int = 8, eps: float = 1e-8):
    """(_delta, _zero_float, scale, zero_point) from host scalars (oeh_set_quant_range_host)."""
    out = (C.c_float * 4)()
    L.check(L.lib().oeh_set_quant_range_host(float(x_min), float(x_max), int(n_bits), float(eps), out))
    return tuple(out)


def estimate_and_quantize(x: torch.Tensor, state: torch.Tensor, momentum: float = 0.9, n_bits: int = 8,
                          eps: float = 1e-8) -> torch.Tensor:
    """QuantizationManager.forward in estimate_ranges state on one GPU (oeh_estimate_and_quantize)."""
    _require_cuda(x, state)
    xc = x.contiguous()
    y = torch.empty_like(xc)
    lib = L.lib()
    need = lib.oeh_minmax_workspace_bytes(xc.numel()) + 256
    ws = _workspace(x.device, need)
    L.check(lib.oeh_estimate_and_quantize(xc.data_ptr(), y.data_ptr(), xc.numel(), _dtype_code(xc), state.data_ptr(),
                                          float(momentum), int(n_bits), float(eps), ws.data_ptr(), ws.numel(),
                                          _stream()))
    return y


# ---- percentile ranges, fused residual tap, outlier metrics ----------------------------

PCT_HIST_WORDS = 4 * 2048       # leading uint32 words of the percentile workspace = the histograms (all-reducible)


def percentile(x: torch.Tensor, q_lo: float, q_hi: float, group=None, sync: bool = False) -> torch.Tensor:
    """np.percentile(x, (q_lo, q_hi)) over the whole tensor, exact, as a 2-element CUDA float64 tensor
    (oeh_percentile; range_estimators.py:89-94).  With ``sync`` and an initialised process group the histograms are
    SUM-all-reduced between the select passes, so every rank gets the percentile of the union of all ranks' tensors
    (equal-sized shards), bit-identical to one process holding the concatenated batch.  No host sync."""
    _require_cuda(x)
    if x.dtype not in (torch.float32, torch.float16):
        raise TypeError(f"Got unsupported ScalarType {str(x.dtype).replace('torch.', '').capitalize()}")   # Tensor.numpy()
    xc = x.contiguous()
    if xc.numel() == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")                               # np.percentile([])
    lib = L.lib()
    out = torch.empty(2, dtype=torch.float64, device=x.device)
    need = lib.oeh_percentile_workspace_bytes()
    ws = _workspace(x.device, need)
    code, n = _dtype_code(xc), xc.numel()
    world = 1
    if sync:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(group)
    if world == 1:
        L.check(lib.oeh_percentile(xc.data_ptr(), n, code, float(q_lo), float(q_hi), out.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream()))
        return out
    import torch.distributed as dist
    ranks, gamma = (C.c_int64 * 4)(), (C.c_double * 2)()
    L.check(lib.oeh_percentile_plan(n * world, float(q_lo), float(q_hi), ranks, gamma))
    L.check(lib.oeh_percentile_begin(ws.data_ptr(), ws.numel(), ranks, _stream()))
    hist = ws[:PCT_HIST_WORDS * 4].view(torch.int32)
    for p in range(lib.oeh_percentile_passes(code)):
        L.check(lib.oeh_percentile_hist(xc.data_ptr(), n, code, p, ws.data_ptr(), _stream()))
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        L.check(lib.oeh_percentile_scan(code, p, ws.data_ptr(), _stream()))
    L.check(lib.oeh_percentile_finish(code, gamma, out.data_ptr(), ws.data_ptr(), _stream()))
    return out


def new_qstate64(device) -> torch.Tensor:
    """float64 side of a percentile-estimated range: [cur_min, cur_max, _delta, _zero_float, initialised]."""
    return torch.zeros(5, dtype=torch.float64, device=device)


def range_update_f64(state64: torch.Tensor, state: torch.Tensor, new_minmax: torch.Tensor, momentum: float = 0.9,
                     n_bits: int = 8, eps: float = 1e-8) -> None:
    """float64 EMA + set_quant_range, mirrored into the fp32 kernel state (oeh_range_update_f64)."""
    _require_cuda(state64, state, new_minmax)
    if state64.dtype != torch.float64 or new_minmax.dtype != torch.float64 or state.dtype != torch.float32:
        raise ValueError("range_update_f64: state64/new_minmax must be float64, state float32")
    L.check(L.lib().oeh_range_update_f64(state64.data_ptr(), state.data_ptr(), new_minmax.data_ptr(), float(momentum),
                                         int(n_bits), float(eps), _stream()))


def add_fake_quant(a: torch.Tensor, b: torch.Tensor, qparams: torch.Tensor, n_bits: int = 8,
                   out: Optional[torch.Tensor] = None, qparams_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qdq(a + [qdq_b] b) in one pass (oeh_add_fake_quant / oeh_tap_add_fake_quant): the residual-sum activation tap with
    fixed ranges, optionally with the output tap of the layer that produced ``b`` folded in."""
    _require_cuda(a, b, qparams, qparams_b)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("add_fake_quant: a and b must have the same shape and dtype")
    ac, bc = a.contiguous(), b.contiguous()
    y = out if out is not None else torch.empty_like(ac)
    if qparams_b is not None:
        L.check(L.lib().oeh_tap_add_fake_quant(ac.data_ptr(), bc.data_ptr(), y.data_ptr(), ac.numel(), _dtype_code(ac),
                                               qparams_b.data_ptr(), qparams.data_ptr(), int(n_bits), _stream()))
        return y
    L.check(L.lib().oeh_add_fake_quant(ac.data_ptr(), bc.data_ptr(), y.data_ptr(), ac.numel(), _dtype_code(ac),
                                       qparams.data_ptr(), int(n_bits), _stream()))
    return y


def outlier_stats(x: torch.Tensor, eps: float = 1e-6, meter: Optional[torch.Tensor] = None):
    """Per-sample (inf_norm, kurtosis) of x viewed as (x.size(0), -1) (oeh_outlier_stats;
    validate_mlm_config.py:513-528, transformers_language/utils.py:9-20).  ``meter``: optional CUDA float64[3]
    running [sum inf_norm, sum kurtosis, samples], updated in place on the device."""
    _require_cuda(x, meter)
    xc = x.contiguous()
    rows = xc.size(0)
    cols = xc.numel() // max(rows, 1)
    inf = torch.empty(rows, dtype=torch.float32, device=x.device)
    kurt = torch.empty(rows, dtype=torch.float32, device=x.device)
    lib = L.lib()
    need = lib.oeh_outlier_stats_workspace_bytes(rows, cols)
    ws = _workspace(x.device, need)
    if meter is not None and (meter.dtype != torch.float64 or meter.numel() < 3):
        raise ValueError("meter must be a CUDA float64 tensor of 3 elements")
    L.check(lib.oeh_outlier_stats(xc.data_ptr(), rows, cols, _dtype_code(xc), float(eps), inf.data_ptr(), kurt.data_ptr(),
                                  meter.data_ptr() if meter is not None else None, ws.data_ptr(), ws.numel(),
                                  _stream()))
    return inf, kurt


def ce_loss(logits: torch.Tensor, labels: torch.Tensor, vocab: Optional[int] = None, ignore_index: int = -100,
            return_row_losses: bool = False, shift_block: int = 0):
    """Mean cross-entropy in fp32 straight from (rows, >=vocab) logits (oeh_ce_loss): F.cross_entropy(
    logits[:, :vocab].float(), labels) without the fp32 copy or a materialised log-softmax.  ``logits`` may be a
    column-padded 2-D view (last dimension contiguous, any row stride); ``labels`` int64 (rows).  Returns a 0-dim
    fp32 device tensor (and the per-row fp32 losses when asked).  ``shift_block`` = T > 0: HF's causal-LM shift over blocks of T
    rows -- row (b, t) is scored against labels[b, t + 1], the last row of each block is ignored (oeh_ce_loss_shifted)."""
    _require_cuda(logits, labels)
    if logits.dim() != 2 or logits.stride(1) != 1:
        raise ValueError("ce_loss: expected 2-D logits whose last dimension is contiguous")
    if labels.dtype != torch.int64 or labels.numel() != logits.size(0):
        raise ValueError("ce_loss: labels must be int64 with one entry per logits row")
    rows = logits.size(0)
    vocab = logits.size(1) if vocab is None else int(vocab)
    if vocab > logits.size(1):
        raise ValueError("ce_loss: vocab exceeds the logits width")
    lab = labels.contiguous()
    lib = L.lib()
    out = torch.empty((), dtype=torch.float32, device=logits.device)
    if return_row_losses:
        ws = torch.empty(rows, dtype=torch.float32, device=logits.device)
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel() * 4
    else:
        ws = _workspace(logits.device, lib.oeh_ce_loss_workspace_bytes(rows))
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    if shift_block < 0 or (shift_block and rows % shift_block):
        raise ValueError("ce_loss: rows must be a whole number of blocks of shift_block tokens")
    L.check(lib.oeh_ce_loss_shifted(logits.data_ptr(), rows, vocab, logits.stride(0), _dtype_code(logits), lab.data_ptr(),
                                    int(ignore_index), int(shift_block), out.data_ptr(), None, ws_ptr, ws_bytes, _stream()))
    return (out, ws) if return_row_losses else out


def add_layer_norm(a: torch.Tensor, b: Optional[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor],
                   eps: float = 1e-12, qparams_res: Optional[torch.Tensor] = None,
                   qparams_out: Optional[torch.Tensor] = None, n_bits: int = 8,
                   out: Optional[torch.Tensor] = None, qparams_a: Optional[torch.Tensor] = None,
                   return_sum: bool = False, gate=None):
    """[qdq_out] LayerNorm([qdq_res]([qdq_a] a + b)) over the last dimension in one pass (oeh_add_layer_norm).  ``b``
    None = plain LayerNorm; ``qparams_*`` = device [scale, zero_point] of fixed-range taps; ``return_sum``: also return
    the (tapped) sum, which pre-LayerNorm blocks carry on as the residual.  ``gate`` = (w (H,64) fp32, b (H) fp32) of the
    Linear per-token gate predictors of the attention module that consumes the result: the (B,H,T) fp32 gate logits are
    produced by the same kernel and returned last (a (B,T,64 H) input is required)."""
    _require_cuda(a, b, weight, bias, qparams_a, qparams_res, qparams_out)
    if b is not None and (b.shape != a.shape or b.dtype != a.dtype):
        raise ValueError("add_layer_norm: a and b must have the same shape and dtype")
    cols = a.shape[-1]
    if weight.shape != (cols,) or weight.dtype != a.dtype or (bias is not None and (bias.shape != (cols,) or bias.dtype != a.dtype)):
        raise ValueError("add_layer_norm: weight / bias must be (cols,) tensors of the input dtype")
    ac = a.contiguous()
    bc = b.contiguous() if b is not None else None
    wc = weight.contiguous()
    bic = bias.contiguous() if bias is not None else None
    y = out if out is not None else torch.empty_like(ac)
    total = torch.empty_like(ac) if return_sum else None
    rows = ac.numel() // max(cols, 1)
    logits = gw = gb = None
    T_len, H = 1, 0
    if gate is not None:
        gw, gb = gate
        H = gw.shape[0]
        if ac.dim() != 3 or gw.shape != (H, 64) or gb.shape != (H,) or cols != 64 * H or gw.dtype != torch.float32 \
                or gb.dtype != torch.float32 or not gw.is_contiguous() or not gb.is_contiguous():
            raise ValueError("add_layer_norm: gate needs a (B,T,64*H) input and contiguous fp32 (H,64) / (H,) predictors")
        T_len = ac.shape[1]
        logits = torch.empty((ac.shape[0], H, T_len), dtype=torch.float32, device=ac.device)
    L.check(L.lib().oeh_add_layer_norm_gate(ac.data_ptr(), bc.data_ptr() if bc is not None else None, wc.data_ptr(),
                                            bic.data_ptr() if bic is not None else None, y.data_ptr(), rows, cols, float(eps),
                                            _dtype_code(ac), qparams_a.data_ptr() if qparams_a is not None else None,
                                            qparams_res.data_ptr() if qparams_res is not None else None,
                                            qparams_out.data_ptr() if qparams_out is not None else None, int(n_bits),
                                            total.data_ptr() if total is not None else None,
                                            gw.data_ptr() if gw is not None else None, gb.data_ptr() if gb is not None else None,
                                            logits.data_ptr() if logits is not None else None, int(T_len), int(H), _stream()))
    res = (y, total) if return_sum else (y,)
    if gate is not None:
        res = res + (logits,)
    return res if len(res) > 1 else res[0]


def gelu(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """erf GELU ``x * 0.5 * (1 + erf(x / sqrt 2))`` of a fp16 / bf16 tensor (oeh_gelu_erf); ``out`` may be ``x``."""
    _require_cuda(x)
    xc = x.contiguous()
    y = out if out is not None else torch.empty_like(xc)
    if y.shape != xc.shape or y.dtype != xc.dtype or not y.is_contiguous():
        raise ValueError("gelu: out must be a contiguous tensor of the input's shape and dtype")
    L.check(L.lib().oeh_gelu_erf(xc.data_ptr(), y.data_ptr(), xc.numel(), _dtype_code(xc), _stream()))
    return y


def fake_quant_asym_segments(x: torch.Tensor, qparams, n_bits: int = 8, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fake-quant of ``len(qparams)`` equal column segments of a (..., n_seg*seg_cols) tensor, each with its own device
    [scale, zero_point] (oeh_fake_quant_asym_segments): one launch for the output of a fused q|k|v projection."""
    _require_cuda(x, *qparams)
    n_seg = len(qparams)
    cols = x.shape[-1]
    if not 1 <= n_seg <= 4 or cols % n_seg:
        raise ValueError("fake_quant_asym_segments: 1..4 equal segments expected")
    xc = x.contiguous()
    y = out if out is not None else torch.empty_like(xc)
    rows = xc.numel() // max(cols, 1)
    ptrs = (C.c_void_p * n_seg)(*[q.data_ptr() for q in qparams])
    L.check(L.lib().oeh_fake_quant_asym_segments(xc.data_ptr(), y.data_ptr(), rows, cols // n_seg, n_seg, cols, cols,
                                                 _dtype_code(xc), ptrs, int(n_bits), _stream()))
    return y


def linear_fwd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: str = "none",
               qparams=None, n_bits: int = 8, out: Optional[torch.Tensor] = None, tile_n: int = 0) -> torch.Tensor:
    """``tap(act(x @ weight.T + bias))`` in ONE tcgen05 GEMM launch (oeh_linear_fwd).  x: (..., K) fp16 / bf16 with a contiguous
    last dim, weight: (N, K) (nn.Linear layout), bias: (N) or None.  ``act``: "none" | "relu" | "gelu" (erf form).
    ``qparams``: None, one device tensor [scale, zero_point] (a frozen output tap) or a list of up to 4 of them (equal column
    segments with their own ranges -- the fused q|k|v projection)."""
    _require_cuda(x, weight, bias)
    if x.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != x.dtype or (bias is not None and bias.dtype != x.dtype):
        raise TypeError("linear_fwd: x, weight and bias must share one 16-bit dtype")
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.dim() != 2 or weight.shape[1] != K:
        raise ValueError(f"linear_fwd: weight {tuple(weight.shape)} does not match x (..., {K})")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    w = weight if weight.stride(-1) == 1 else weight.contiguous()
    M = x2.shape[0]
    y = out if out is not None else torch.empty((M, N), dtype=x.dtype, device=x.device)
    if y.shape[-1] != N or y.numel() != M * N or y.stride(-1) != 1:
        raise ValueError("linear_fwd: out must be an (M, N) tensor with a contiguous last dim")
    y2 = y.view(-1, N) if y.dim() != 2 else y
    a = L.OehLinearArgs()
    a.x, a.w, a.y = x2.data_ptr(), w.data_ptr(), y2.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.ldx, a.ldw, a.ldy = x2.stride(0) if M > 1 else K, w.stride(0), y2.stride(0) if M > 1 else N
    a.M, a.N, a.K = M, N, K
    a.dtype = _dtype_code(x)
    a.act = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "gelu": L.ACT_GELU_ERF}[act]
    qs = [] if qparams is None else ([qparams] if torch.is_tensor(qparams) else list(qparams))
    for i, t in enumerate(qs):
        if t.dtype != torch.float32 or t.numel() < 2 or not t.is_cuda:
            raise ValueError("linear_fwd: qparams must be CUDA fp32 tensors [scale, zero_point]")
        a.qparams[i] = t.data_ptr()
    a.n_seg, a.n_bits, a.tile_n = len(qs), int(n_bits), int(tile_n)
    a.stream = _stream()
    L.check(L.lib().oeh_linear_fwd(C.byref(a)))
    return y if out is not None else y.view(*x.shape[:-1], N)


def linear_add_layer_norm_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """True if ``linear_add_layer_norm`` can take this Linear: 16-bit CUDA tensors and an output width one cluster can normalise
    (oeh_linear_add_layer_norm_supported: 192 c or 256 c columns, c <= 8), K a multiple of 8."""
    return (x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and weight.dtype == x.dtype and weight.dim() == 2
            and weight.shape[1] % 8 == 0 and bool(L.lib().oeh_linear_add_layer_norm_supported(int(weight.shape[0]))))


def linear_add_layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor],
                          ln_weight: torch.Tensor, ln_bias: Optional[torch.Tensor], eps: float = 1e-12, *,
                          qparams_a: Optional[torch.Tensor] = None, qparams_res: Optional[torch.Tensor] = None,
                          qparams_out: Optional[torch.Tensor] = None, n_bits: int = 8, return_sum: bool = False,
                          out: Optional[torch.Tensor] = None):
    """``[qdq_out] LayerNorm([qdq_res]([qdq_a](x @ weight.T + bias) + residual))`` in ONE launch (oeh_linear_add_layer_norm): the
    GEMM runs on a thread-block cluster that owns whole rows and normalises them in its epilogue.  x: (..., K), weight: (N, K),
    residual: (..., N) or None; ``return_sum``: also return the (tapped) sum, the residual a pre-LayerNorm block carries on.
    Same arithmetic as ``linear`` followed by ``add_layer_norm``."""
    _require_cuda(x, weight, bias, residual, ln_weight, ln_bias, qparams_a, qparams_res, qparams_out)
    K, N = x.shape[-1], weight.shape[0]
    if x.dtype not in (torch.float16, torch.bfloat16) or any(t is not None and t.dtype != x.dtype
                                                             for t in (weight, bias, residual, ln_weight, ln_bias)):
        raise TypeError("linear_add_layer_norm: every tensor must share one 16-bit dtype")
    if weight.dim() != 2 or weight.shape[1] != K:
        raise ValueError(f"linear_add_layer_norm: weight {tuple(weight.shape)} does not match x (..., {K})")
    if ln_weight.shape != (N,) or (ln_bias is not None and ln_bias.shape != (N,)) or (bias is not None and bias.shape != (N,)):
        raise ValueError("linear_add_layer_norm: bias / LayerNorm parameters must be (N,)")
    if return_sum and residual is None:
        raise ValueError("linear_add_layer_norm: return_sum needs a residual")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    w = weight if weight.stride(-1) == 1 else weight.contiguous()
    M = x2.shape[0]
    res2 = None
    if residual is not None:
        if residual.numel() != M * N or residual.shape[-1] != N:
            raise ValueError("linear_add_layer_norm: residual must be (..., N) with the rows of x")
        res2 = residual.contiguous()
    lead = x.shape[:-1]
    y = out if out is not None else torch.empty((*lead, N), dtype=x.dtype, device=x.device)
    if y.numel() != M * N or not y.is_contiguous():
        raise ValueError("linear_add_layer_norm: out must be a contiguous (..., N) tensor")
    total = torch.empty((*lead, N), dtype=x.dtype, device=x.device) if return_sum else None
    a = L.OehLinearLnArgs()
    a.x, a.w, a.y = x2.data_ptr(), w.data_ptr(), y.data_ptr()
    a.bias = bias.contiguous().data_ptr() if bias is not None else None
    a.residual = res2.data_ptr() if res2 is not None else None
    lw, lb = ln_weight.contiguous(), (ln_bias.contiguous() if ln_bias is not None else None)
    a.ln_weight, a.ln_bias = lw.data_ptr(), (lb.data_ptr() if lb is not None else None)
    a.sum_out = total.data_ptr() if total is not None else None
    a.ldx, a.ldw = (x2.stride(0) if M > 1 else K), w.stride(0)
    a.M, a.N, a.K = M, N, K
    a.dtype, a.n_bits, a.eps = _dtype_code(x), int(n_bits), float(eps)
    for name, t in (("q_a", qparams_a), ("q_res", qparams_res), ("q_out", qparams_out)):
        if t is not None:
            if t.dtype != torch.float32 or t.numel() < 2:
                raise ValueError("linear_add_layer_norm: qparams must be CUDA fp32 tensors [scale, zero_point]")
            setattr(a, name, t.data_ptr())
    a.stream = _stream()
    L.check(L.lib().oeh_linear_add_layer_norm(C.byref(a)))
    return (y, total) if return_sum else y


def embed_sum(input_ids: torch.Tensor, word: torch.Tensor, pos: torch.Tensor, position_ids: Optional[torch.Tensor] = None,
              pos_offset: int = 0) -> torch.Tensor:
    """``word[ids] + pos[position_ids + pos_offset]`` in one launch (oeh_embed_sum): the OPT decoder's input, HF
    ``embed_tokens(ids) + embed_positions(position_ids)`` (whose forward adds its offset of 2).  (B,T) int64 ids; missing
    ``position_ids`` = 0..T-1.  One rounding to the tensor dtype, like the ATen add."""
    _require_cuda(input_ids, word, pos, position_ids)
    B, T = input_ids.shape
    cols = word.shape[1]
    if pos.dtype != word.dtype or pos.shape[1] != cols or not pos.is_contiguous() or not word.is_contiguous():
        raise ValueError("embed_sum: tables must be contiguous, of one dtype and one width")
    ids = input_ids.contiguous()
    pp = None if position_ids is None else position_ids.expand(B, T).contiguous()
    y = torch.empty((B, T, cols), dtype=word.dtype, device=word.device)
    L.check(L.lib().oeh_embed_sum(ids.data_ptr(), pp.data_ptr() if pp is not None else None, word.data_ptr(), pos.data_ptr(),
                                  y.data_ptr(), B * T, cols, T, word.shape[0], pos.shape[0], int(pos_offset), _dtype_code(word),
                                  _stream()))
    return y


def embed_layer_norm(input_ids: torch.Tensor, word: torch.Tensor, pos: Optional[torch.Tensor], tok_type: Optional[torch.Tensor],
                     weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float, token_type_ids: Optional[torch.Tensor] = None,
                     position_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``LayerNorm((word[ids] + tok_type[token_type_ids]) + pos[position_ids])`` in one launch (oeh_embed_layer_norm): HF
    BertEmbeddings.forward for (B,T) int64 ids; missing ``token_type_ids`` = type 0, missing ``position_ids`` = 0..T-1."""
    _require_cuda(input_ids, word, pos, tok_type, weight, bias, token_type_ids, position_ids)
    B, T = input_ids.shape
    cols = word.shape[1]
    for t in (pos, tok_type):
        if t is not None and (t.dtype != word.dtype or t.shape[1] != cols or not t.is_contiguous()):
            raise ValueError("embed_layer_norm: tables must be contiguous, of one dtype and one width")
    ids = input_ids.contiguous()
    tt = None if token_type_ids is None else token_type_ids.expand(B, T).contiguous()
    pp = None if position_ids is None else position_ids.expand(B, T).contiguous()
    y = torch.empty((B, T, cols), dtype=word.dtype, device=word.device)
    L.check(L.lib().oeh_embed_layer_norm(ids.data_ptr(), tt.data_ptr() if tt is not None else None,
                                         pp.data_ptr() if pp is not None else None, word.contiguous().data_ptr(),
                                         tok_type.data_ptr() if tok_type is not None else None,
                                         pos.data_ptr() if pos is not None else None, weight.contiguous().data_ptr(),
                                         bias.contiguous().data_ptr() if bias is not None else None, y.data_ptr(), B * T, cols, T,
                                         word.shape[0], float(eps), _dtype_code(word), _stream()))
    return y
