"""Activation fake-quant surface of the reference (``quantization/`` package), re-hosted on the CUDA
kernels of liboeh.so.  Same class names, constructor kwargs, methods, state_dict keys and error
behaviour:

    AsymmetricUniformQuantizer   quantization/quantizers/uniform_quantizers.py:12-240
    SymmetricUniformQuantizer    quantization/quantizers/uniform_quantizers.py:243-310   (weights, cached)
    RunningMinMaxEstimator       quantization/range_estimators.py:77-106
    CurrentMinMaxEstimator       quantization/range_estimators.py:54-74
    QuantizationManager          quantization/quantization_manager.py:11-108
    QuantizedModule / QuantizedActivation   quantization/base_quantized_classes.py:38-199
    QuantLinear                  quantization/autoquant_utils.py:18-20 + hijacker.py:27-134

Design difference (B200-first): the running range and the derived (scale, zero_point) of an activation
quantizer live in ONE 8-float device tensor (``oeh_qstate_slot`` layout) shared by the estimator and
the quantizer of a QuantizationManager.  ``estimate_ranges`` forwards are min/max -> [NCCL MIN/MAX
all-reduce] -> EMA + set_quant_range -> quantize, all enqueued on the stream with no host round trip;
the reference's ``_delta`` / ``_zero_float`` / ``current_xmin`` / ``current_xmax`` buffers are exposed as
views of that tensor and saved / loaded under the reference's state_dict keys.
"""
from __future__ import annotations

from collections import namedtuple
from enum import Flag, auto
from functools import partial
from typing import Optional

import copy

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib as L
from . import ops
from ._lib import QuantizerNotInitializedError

# ---------------------------------------------------------------------------------------------
# enums (quantization/utils.py:35-59, qstates.py:7-13)
# ---------------------------------------------------------------------------------------------


class BaseEnumOptions(Flag):
    def __str__(self):
        return self.name

    @classmethod
    def list_names(cls):
        # __members__, not iteration: since Python 3.11 iterating a Flag skips zero-valued and composite
        # members (AttentionGateType.none = 0, conditional_per_token = 3), which the reference's
        # argparse `choices=` relied on under the Python it was written for.
        return list(cls.__members__)


class ClassEnumOptions(BaseEnumOptions):
    @property
    def cls(self):
        return self.value.cls

    def __call__(self, *args, **kwargs):
        return self.value.cls(*args, **kwargs)


MethodMap = partial(namedtuple("MethodMap", ["value", "cls"]), auto())


class Qstates(BaseEnumOptions):
    estimate_ranges = auto()        # ranges are updated in eval and train mode
    fix_ranges = auto()             # quantization ranges are fixed for train and eval
    learn_ranges = auto()           # quantization params are nn.Parameters (not supported: eval-only build)
    estimate_ranges_train = auto()  # ranges are updated during train and fixed for eval


# process group over which batch min/max are MIN/MAX-reduced BEFORE the EMA during range estimation, so that
# an N-GPU estimation equals a single-process run at the N-times larger global batch
_RANGE_SYNC_GROUP = None
_RANGE_SYNC_ENABLED = False


def set_range_sync(enabled: bool, group=None) -> None:
    """Enable / disable the cross-rank MIN/MAX all-reduce of estimation ranges (torch.distributed)."""
    global _RANGE_SYNC_GROUP, _RANGE_SYNC_ENABLED
    _RANGE_SYNC_ENABLED, _RANGE_SYNC_GROUP = bool(enabled), group


RANGE_SYNC_COLLECTIVES = 0      # all-reduces issued by _sync_minmax (instrumentation for tests / bench.py)


def _sync_minmax(mm: torch.Tensor) -> torch.Tensor:
    """[min, max] (or n stacked pairs, shape (n, 2)) -> ONE all-reduce as MAX over [-min, max] (exact, order independent)."""
    global RANGE_SYNC_COLLECTIVES
    if not _RANGE_SYNC_ENABLED:
        return mm
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(_RANGE_SYNC_GROUP) == 1:
        return mm
    buf = torch.stack((-mm[..., 0], mm[..., 1]), dim=-1).contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=_RANGE_SYNC_GROUP)
    RANGE_SYNC_COLLECTIVES += 1
    return torch.stack((-buf[..., 0], buf[..., 1]), dim=-1).contiguous()


class _StateHolder(nn.Module):
    """Mixin: owns (or borrows from a QuantizationManager) the 8-float device state."""

    def __init__(self):
        super().__init__()
        self._qstate: Optional[torch.Tensor] = None
        self._owns_state = True
        self._host_init = False        # host mirror of state[QS_INITIALIZED]: no device read in forward
        # percentile branch only: float64 [cur_min, cur_max, _delta, _zero_float, initialised] (the reference keeps
        # np.percentile's float64 through the EMA and set_quant_range); shared estimator <-> quantizer like _qstate
        self._qstate64: Optional[torch.Tensor] = None

    def _ensure_state(self, device) -> torch.Tensor:
        if self._qstate is None:
            self._qstate = ops.new_qstate(device)
        elif self._qstate.device != torch.device(device) and self._owns_state:
            self._qstate = self._qstate.to(device)
        return self._qstate

    def _bind_state(self, state: Optional[torch.Tensor]):
        self._qstate = state
        self._owns_state = False

    def _ensure_state64(self, device) -> torch.Tensor:
        if self._qstate64 is None:
            self._qstate64 = ops.new_qstate64(device)
        elif self._qstate64.device != torch.device(device):
            self._qstate64 = self._qstate64.to(device)
        return self._qstate64

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        if self._owns_state and self._qstate is not None:
            moved = fn(self._qstate)
            self._qstate = moved if moved.dtype == torch.float32 else self._qstate.to(moved.device)
        if self._qstate64 is not None and self._qstate is not None:
            self._qstate64 = self._qstate64.to(self._qstate.device)
        return self


# ---------------------------------------------------------------------------------------------
# quantizers
# ---------------------------------------------------------------------------------------------


class QuantizerBase(_StateHolder):
    """quantization/quantizers/base_quantizers.py:4-51."""

    def __init__(self, n_bits, *args, per_channel=False, act_quant=False, **kwargs):
        super().__init__()
        if per_channel:
            raise NotImplementedError("per-channel quantization is outside the accelerated hot path")
        self.n_bits = n_bits
        self.act_quant = act_quant
        self.per_channel = per_channel
        self.state = None
        self.x_min_fp32 = self.x_max_fp32 = None

    def extra_repr(self):
        return "n_bits={}, per_channel={}, is_initalized={}".format(self.n_bits, self.per_channel, self.is_initialized)

    def reset(self):
        self._host_init = False
        if self._qstate is not None:
            self._qstate[L.QS_DELTA:L.QS_INITIALIZED + 1].zero_()

    def make_range_trainable(self):
        raise NotImplementedError("learn_ranges (trainable quantizer ranges) is outside the eval hot path")


class AsymmetricUniformQuantizer(QuantizerBase):
    """Asymmetric uniform fake-quantizer (uniform_quantizers.py:12-240); per-tensor, linear scale domain.
    ``forward`` runs oeh_fake_quant_asym: fp32 tensors bit-exact with the reference."""

    def __init__(self, n_bits, scale_domain="linear", grad_scaling=False, eps=1e-8, **kwargs):
        super().__init__(n_bits=n_bits, **kwargs)
        assert scale_domain in ("linear", "log")
        if scale_domain != "linear" or grad_scaling:
            raise NotImplementedError("log scale domain / LSQ grad scaling are outside the accelerated hot path")
        self.scale_domain = scale_domain
        self.grad_scaling = grad_scaling
        self.eps = eps
        self._range_is_f64 = False

    # -- reference buffers as views of the device state --------------------------------------------
    @property
    def _f64(self) -> bool:
        """The current range came from float64 tensors (percentile estimation): buffers are float64 like the reference's."""
        return self._qstate64 is not None and self._range_is_f64

    @property
    def _delta(self):
        if not self._host_init:
            return None
        return self._qstate64[2] if self._f64 else self._qstate[L.QS_DELTA]

    @property
    def _zero_float(self):
        if not self._host_init:
            return None
        return self._qstate64[3] if self._f64 else self._qstate[L.QS_ZERO_FLOAT]

    @property
    def delta(self):
        if self._host_init:
            return self._delta
        raise QuantizerNotInitializedError()

    @property
    def zero_float(self):
        if self._host_init:
            return self._zero_float
        raise QuantizerNotInitializedError()

    @property
    def is_initialized(self):
        return self._host_init

    @property
    def symmetric(self):
        return False

    @property
    def int_min(self):
        return 0.0

    @property
    def int_max(self):
        return 2.0 ** self.n_bits - 1

    @property
    def scale(self):
        self.delta   # raises if not initialised
        return self._qstate[L.QS_SCALE]

    @property
    def zero_point(self):
        self.delta
        return self._qstate[L.QS_ZERO_POINT]

    @property
    def qparams(self) -> torch.Tensor:
        """Device tensor [scale, zero_point] consumed by the kernels."""
        self.delta
        return self._qstate[L.QS_SCALE:L.QS_SCALE + 2]

    @property
    def x_max(self):
        return self.scale * (self.int_max - self.zero_point)

    @property
    def x_min(self):
        return self.scale * (self.int_min - self.zero_point)

    def to_integer_forward(self, x_float):
        return ops.fake_quant_asym(x_float, self.qparams, self.n_bits, emit_codes=True).view(x_float.shape)

    def forward(self, x_float):
        return ops.fake_quant_asym(x_float, self.qparams, self.n_bits).view(x_float.shape)

    def set_quant_range(self, x_min, x_max):
        """uniform_quantizers.py:204-224.  Host scalars go through oeh_set_quant_range_host, device tensors
        through oeh_range_update semantics without the EMA (same fp32 op order either way)."""
        self.x_min_fp32, self.x_max_fp32 = x_min, x_max
        self._range_is_f64 = False
        if torch.is_tensor(x_min) and x_min.is_cuda and x_min.dtype == torch.float64:
            # float64 tensors (percentile estimators): the reference's arithmetic stays in float64 (:218-219)
            if x_min.dim() > 0 and x_min.numel() > 1:
                raise ValueError("x_min and x_max must be a float or 1-D Tensor for per-tensor quantization "
                                 "(per_channel=False)")
            st = self._ensure_state(x_min.device)
            st64 = self._ensure_state64(x_min.device)
            mm = torch.stack((x_min.reshape(()), x_max.reshape(())))
            keep32, keep64 = st[:2].clone(), st64[:2].clone()
            st64[4] = 0.0                                    # take the range as is (no EMA)
            ops.range_update_f64(st64, st, mm, 0.0, self.n_bits, self.eps)
            st[:2], st64[:2] = keep32, keep64
            self._range_is_f64 = True
        elif torch.is_tensor(x_min) and x_min.is_cuda:
            st = self._ensure_state(x_min.device)
            if x_min.dim() > 0 and x_min.numel() > 1:
                raise ValueError("x_min and x_max must be a float or 1-D Tensor for per-tensor quantization "
                                 "(per_channel=False)")
            mm = torch.stack((x_min.reshape(()).float(), x_max.reshape(()).float()))
            keep = st[:2].clone()
            st[L.QS_INITIALIZED] = 0.0                      # take the range as is (no EMA)
            ops.range_update(st, mm, 0.0, self.n_bits, self.eps)
            st[:2] = keep
        else:
            lo = float(x_min)
            hi = float(x_max)
            d, zf, s, zp = ops.set_quant_range_host(lo, hi, self.n_bits, self.eps)
            dev = self._qstate.device if self._qstate is not None else torch.device("cuda", torch.cuda.current_device())
            st = self._ensure_state(dev)
            st[L.QS_DELTA:L.QS_INITIALIZED + 1] = torch.tensor([d, zf, s, zp, 1.0], dtype=torch.float32).to(dev)
        self._host_init = True

    def fix_ranges(self):
        pass   # ranges are never nn.Parameters in this build

    # -- state_dict under the reference's keys ------------------------------------------------------
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if self._host_init:
            destination[prefix + "_delta"] = self._qstate[L.QS_DELTA].clone()
            destination[prefix + "_zero_float"] = self._qstate[L.QS_ZERO_FLOAT].clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        kd, kz = prefix + "_delta", prefix + "_zero_float"
        if kd in state_dict and kz in state_dict:
            d = state_dict.pop(kd).reshape(()).float()
            z = state_dict.pop(kz).reshape(()).float()
            dev = self._qstate.device if self._qstate is not None else (d.device if d.is_cuda else torch.device(
                "cuda", torch.cuda.current_device()))
            st = self._ensure_state(dev)
            d, z = d.to(dev), z.to(dev)
            st[L.QS_DELTA], st[L.QS_ZERO_FLOAT] = d, z
            st[L.QS_SCALE] = torch.clamp(d, min=self.eps)                          # :72-76
            st[L.QS_ZERO_POINT] = torch.clamp(torch.round(z), self.int_min, self.int_max)   # :79-82
            st[L.QS_INITIALIZED] = 1.0
            self._host_init = True
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class SymmetricUniformQuantizer(nn.Module):
    """Symmetric uniform quantizer (uniform_quantizers.py:243-310).  Used for WEIGHTS only: the result is
    computed once and cached by QuantLinear (hijacker.py:84-101), so it is evaluated with plain device tensor
    ops rather than a dedicated kernel (SURVEY section 8f rank 2 lists a fused version as 'next')."""

    def __init__(self, n_bits, scale_domain="linear", grad_scaling=False, eps=1e-8, per_channel=False,
                 act_quant=False, **kwargs):
        super().__init__()
        if per_channel or scale_domain != "linear" or grad_scaling:
            raise NotImplementedError("per-channel / log-domain / LSQ symmetric quantization is out of scope")
        self.n_bits, self.eps, self.per_channel, self.act_quant = n_bits, eps, per_channel, act_quant
        self.scale_domain, self.grad_scaling = scale_domain, grad_scaling
        self.state = None
        self.x_min_fp32 = self.x_max_fp32 = None
        self.register_buffer("_delta", None)
        self.register_buffer("_signed", None)
        self._signed_host = None

    @property
    def delta(self):
        if self._delta is not None:
            return self._delta
        raise QuantizerNotInitializedError()

    @property
    def is_initialized(self):
        return self._delta is not None

    @property
    def symmetric(self):
        return True

    @property
    def signed(self):
        if self._signed is None:
            raise QuantizerNotInitializedError()
        if self._signed_host is None:
            self._signed_host = bool(self._signed.item())     # once, at weight-cache time
        return self._signed_host

    @property
    def int_min(self):
        return -(2.0 ** (self.n_bits - 1)) if self.signed else 0

    @property
    def int_max(self):
        return 2.0 ** (self.n_bits - self.signed) - 1

    @property
    def scale(self):
        return torch.clamp(self.delta, min=self.eps)

    @property
    def zero_point(self):
        return 0.0

    def set_quant_range(self, x_min, x_max):
        self.x_min_fp32, self.x_max_fp32 = x_min, x_max
        if not torch.is_tensor(x_min):
            x_min, x_max = torch.tensor(x_min).float(), torch.tensor(x_max).float()
        x_min = torch.min(x_min, torch.zeros_like(x_min))                  # :199
        x_max = torch.max(x_max, torch.ones_like(x_max) * self.eps)        # :200
        self._signed = x_min.min() < 0                                     # :296
        self._signed_host = None
        x_absmax = torch.max(x_min.abs(), x_max)
        # :298-299.  Tensor / tensor: ATen-CUDA turns `tensor / python_scalar` into a multiply by the reciprocal (1 ulp off
        # the IEEE quotient the reference gets on the CPU and the oracle is pinned to)
        self._delta = torch.div(x_absmax, torch.full_like(x_absmax, self.int_max)).detach()

    def to_integer_forward(self, x_float):
        x_int = torch.round(x_float / self.scale) + self.zero_point
        return torch.clamp(x_int, self.int_min, self.int_max)

    def forward(self, x_float):
        return self.scale * (self.to_integer_forward(x_float) - self.zero_point)

    def reset(self):
        self._delta = None

    def fix_ranges(self):
        pass

    def make_range_trainable(self):
        raise NotImplementedError("learn_ranges is outside the eval hot path")

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        # ``_delta`` / ``_signed`` are None buffers until a range is set, so a fresh instance would reject the keys of a
        # checkpoint saved after one forward.  The reference sizes them with a dummy forward before loading
        # (base_quantized_model.py:34-82); here they are simply adopted.
        for name in ("_delta", "_signed"):
            key = prefix + name
            if key in state_dict:
                cur = self._buffers.get(name)
                if cur is None or cur.shape != state_dict[key].shape:
                    # sized like the checkpoint's; the base class then copies the value in (0-dim tensors mix with
                    # CUDA operands on any device)
                    dev = cur.device if cur is not None else state_dict[key].device
                    self._buffers[name] = torch.empty_like(state_dict[key], device=dev)
        self._signed_host = None
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class QMethods(ClassEnumOptions):
    symmetric_uniform = MethodMap(SymmetricUniformQuantizer)
    asymmetric_uniform = MethodMap(AsymmetricUniformQuantizer)


# ---------------------------------------------------------------------------------------------
# range estimators
# ---------------------------------------------------------------------------------------------


class RangeEstimatorBase(_StateHolder):
    def __init__(self, *args, per_channel=False, quantizer=None, **kwargs):
        super().__init__()
        if per_channel:
            raise NotImplementedError("per-channel range estimation is outside the accelerated hot path")
        self.per_channel = per_channel
        self.quantizer = None   # not registered as a submodule (the reference hides it from repr the same way)
        object.__setattr__(self, "quantizer", quantizer)

    @property
    def current_xmin(self):
        if not self._host_init:
            return None
        return self._qstate64[0] if getattr(self, "percentile", None) else self._qstate[L.QS_CUR_MIN]

    @property
    def current_xmax(self):
        if not self._host_init:
            return None
        return self._qstate64[1] if getattr(self, "percentile", None) else self._qstate[L.QS_CUR_MAX]

    def reset(self):
        self._host_init = False
        if self._qstate is not None:
            self._qstate[:2].zero_()
            self._qstate[L.QS_INITIALIZED] = 0.0
        if self._qstate64 is not None:
            self._qstate64.zero_()

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if self._host_init:
            destination[prefix + "current_xmin"] = self.current_xmin.clone()
            destination[prefix + "current_xmax"] = self.current_xmax.clone()
        # the reference registers the quantizer as a submodule of its estimator (range_estimators.py:18), so its
        # checkpoints carry the quantizer buffers a second time under "<estimator>.quantizer.*": emit them too
        q = self.quantizer
        if q is not None and q.is_initialized:
            q._save_to_state_dict(destination, prefix + "quantizer.", keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        for k in [k for k in state_dict if k.startswith(prefix + "quantizer.")]:
            state_dict.pop(k)              # duplicate of "<manager>.quantizer.*" (see _save_to_state_dict)
        ka, kb = prefix + "current_xmin", prefix + "current_xmax"
        if ka in state_dict and kb in state_dict:
            lo, hi = state_dict.pop(ka).reshape(()).float(), state_dict.pop(kb).reshape(()).float()
            dev = self._qstate.device if self._qstate is not None else (lo.device if lo.is_cuda else torch.device(
                "cuda", torch.cuda.current_device()))
            st = self._ensure_state(dev)
            st[L.QS_CUR_MIN], st[L.QS_CUR_MAX] = lo.to(dev), hi.to(dev)
            self._host_init = True
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def __repr__(self):
        lines = self.extra_repr().split("\n")
        extra_str = lines[0] if len(lines) == 1 else "\n  " + "\n  ".join(lines) + "\n"
        return self._get_name() + "(" + extra_str + ")"


class CurrentMinMaxEstimator(RangeEstimatorBase):
    """range_estimators.py:54-74, per-tensor.  With ``percentile`` the range is np.percentile(x, (p, 100-p)) --
    the reference passes the pair in THIS order (:65-67), i.e. x_min is the upper percentile; kept as is."""

    def __init__(self, *args, percentile=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.percentile = percentile

    def forward(self, x):
        st = self._ensure_state(x.device)
        if self.percentile:
            st64 = self._ensure_state64(x.device)
            st64[:2] = ops.percentile(x, self.percentile, 100 - self.percentile, group=_RANGE_SYNC_GROUP,
                                      sync=_RANGE_SYNC_ENABLED)
            st[:2] = st64[:2].float()
            self._host_init = True
            return st64[0], st64[1]
        mm = _sync_minmax(ops.minmax(x))
        st[:2] = mm
        self._host_init = True
        return st[L.QS_CUR_MIN], st[L.QS_CUR_MAX]


class RunningMinMaxEstimator(RangeEstimatorBase):
    """range_estimators.py:77-106: EMA of the per-batch global min/max, first batch taken as is."""

    def __init__(self, *args, momentum=0.9, percentile=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.momentum = momentum
        self.percentile = percentile

    def _ema_into_state(self, x, n_bits: int, eps: float):
        """min/max -> [all-reduce] -> EMA -> set_quant_range, all on the device (state shared with the quantizer).
        Percentile branch (range_estimators.py:89-94): exact np.percentile(x, (100-p, p)) by radix select (histograms
        all-reduced when range sync is on), then the EMA and set_quant_range in float64 like the reference."""
        st = self._ensure_state(x.device)
        if self.percentile:
            st64 = self._ensure_state64(x.device)
            mm64 = ops.percentile(x, 100 - self.percentile, self.percentile, group=_RANGE_SYNC_GROUP,
                                  sync=_RANGE_SYNC_ENABLED)
            ops.range_update_f64(st64, st, mm64, self.momentum, n_bits, eps)
            self._host_init = True
            q = self.quantizer
            if q is not None and isinstance(q, _StateHolder) and q._qstate is st:
                q._qstate64, q._range_is_f64 = st64, True
            return st
        mm = _sync_minmax(ops.minmax(x))
        ops.range_update(st, mm, self.momentum, n_bits, eps)
        self._host_init = True
        return st

    def forward(self, x):
        q = self.quantizer
        st = self._ema_into_state(x, q.n_bits if q is not None else 8, q.eps if q is not None else 1e-8)
        if self.percentile:
            return self._qstate64[0], self._qstate64[1]
        return st[L.QS_CUR_MIN], st[L.QS_CUR_MAX]


class RangeEstimators(ClassEnumOptions):
    current_minmax = MethodMap(CurrentMinMaxEstimator)
    running_minmax = MethodMap(RunningMinMaxEstimator)


# ---------------------------------------------------------------------------------------------
# manager and activation tap
# ---------------------------------------------------------------------------------------------


class QuantizationManager(nn.Module):
    """quantization/quantization_manager.py:11-108."""

    def __init__(self, qmethod=SymmetricUniformQuantizer, init=CurrentMinMaxEstimator, per_channel=False, x_min=None,
                 x_max=None, qparams=None, init_params=None):
        super().__init__()
        self.state = Qstates.estimate_ranges
        self.qmethod = qmethod
        self.init = init
        self.per_channel = per_channel
        self.qparams = qparams if qparams else {}
        self.init_params = init_params if init_params else {}
        self.range_estimator = None

        self.quantizer = self.qmethod(per_channel=self.per_channel, **self.qparams)
        self.quantizer.state = self.state
        self._shared = isinstance(self.quantizer, AsymmetricUniformQuantizer)

        if x_min is not None and x_max is not None:
            self.set_quant_range(x_min, x_max)
            self.fix_ranges()
        else:
            self.range_estimator = self.init(per_channel=self.per_channel, quantizer=self.quantizer, **self.init_params)
            if self._shared and isinstance(self.range_estimator, _StateHolder):
                # estimator and quantizer share one device state; the manager owns it
                self.range_estimator._owns_state = False
                self.quantizer._owns_state = False

    # shared-state plumbing -------------------------------------------------------------------------
    def _state_for(self, device) -> torch.Tensor:
        q = self.quantizer
        if q._qstate is None:
            q._qstate = ops.new_qstate(device)
        elif q._qstate.device != torch.device(device):
            q._qstate = q._qstate.to(device)
        if self.range_estimator is not None:
            self.range_estimator._qstate = q._qstate
        return q._qstate

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        if self._shared and self.range_estimator is not None and self.quantizer._qstate is not None:
            moved = fn(self.quantizer._qstate)
            if moved.dtype != torch.float32:
                moved = self.quantizer._qstate.to(moved.device)
            self.quantizer._qstate = moved
            self.range_estimator._qstate = moved
        return self

    @property
    def n_bits(self):
        return self.quantizer.n_bits

    def estimate_ranges(self):
        self.state = Qstates.estimate_ranges
        self.quantizer.state = self.state

    def fix_ranges(self):
        if self.quantizer.is_initialized:
            self.state = Qstates.fix_ranges
            self.quantizer.state = self.state
            self.quantizer.fix_ranges()
        else:
            raise QuantizerNotInitializedError()

    def learn_ranges(self):
        self.quantizer.make_range_trainable()

    def estimate_ranges_train(self):
        self.state = Qstates.estimate_ranges_train
        self.quantizer.state = self.state

    def reset_ranges(self):
        self.range_estimator.reset()
        self.quantizer.reset()
        self.estimate_ranges()

    @property
    def estimating(self) -> bool:
        return self.state == Qstates.estimate_ranges or (self.state == Qstates.estimate_ranges_train and self.training)

    def forward(self, x):
        if self.estimating:                                                  # quantization_manager.py:94-100
            if self._shared and isinstance(self.range_estimator, RunningMinMaxEstimator):
                self._state_for(x.device)
                self.range_estimator._ema_into_state(x, self.quantizer.n_bits, self.quantizer.eps)
                self.quantizer._host_init = True
            else:
                if self._shared:
                    self._state_for(x.device)
                cur_xmin, cur_xmax = self.range_estimator(x)
                self.set_quant_range(cur_xmin, cur_xmax)
        return self.quantizer(x)                                             # :102

    def set_quant_range(self, x_min, x_max):
        self.quantizer.set_quant_range(x_min, x_max)

    def extra_repr(self):
        return "state={}".format(self.state.name)


def estimate_and_quantize_group(managers, tensors):
    """``[m(x) for m, x in zip(managers, tensors)]`` for managers in ``estimate_ranges`` state whose inputs do NOT depend
    on each other's outputs (the q / k / v projection taps of one attention call): the batch (min, max) pairs of the whole
    group cross the ranks in ONE all-reduce instead of one per tap, then each tap runs its own
    EMA -> set_quant_range -> quantize exactly as QuantizationManager.forward does (quantization_manager.py:94-102).
    Returns None when the group is not eligible (the caller then runs the taps one by one)."""
    for m in managers:
        if not (isinstance(m, QuantizationManager) and m.estimating and m._shared
                and isinstance(m.range_estimator, RunningMinMaxEstimator) and not m.range_estimator.percentile):
            return None
    mms = _sync_minmax(torch.stack([ops.minmax(x) for x in tensors]))            # (n, 2), one collective
    outs = []
    for m, x, mm in zip(managers, tensors, mms):
        st = m._state_for(x.device)
        ops.range_update(st, mm, m.range_estimator.momentum, m.quantizer.n_bits, m.quantizer.eps)
        m.range_estimator._host_init = True
        m.quantizer._host_init = True
        outs.append(m.quantizer(x))
    return outs


def _set_layer_fix_ranges(layer):
    if isinstance(layer, QuantizationManager):
        if layer.quantizer.is_initialized:
            layer.fix_ranges()


def _set_layer_estimate_ranges(layer):
    if isinstance(layer, QuantizationManager):
        layer.estimate_ranges()


def _set_layer_estimate_ranges_train(layer):
    if isinstance(layer, QuantizationManager):
        if layer.quantizer.is_initialized:
            layer.estimate_ranges_train()


class QuantizedModule(nn.Module):
    """quantization/base_quantized_classes.py:38-179.  ``_quant_a`` / ``_quant_w`` stay 1-element Bool
    buffers (state_dict compatible) but are mirrored host-side so ``forward`` never syncs on them."""

    def __init__(self, *args, method=AsymmetricUniformQuantizer, act_method=None,
                 weight_range_method=CurrentMinMaxEstimator, act_range_method=RunningMinMaxEstimator, n_bits=8,
                 n_bits_act=None, per_channel_weights=False, percentile=None, weight_range_options=None,
                 act_range_options=None, scale_domain="linear", bayesian_bits_kwargs=None, prune_method=None,
                 prune_kwargs=None, **kwargs):
        kwargs.pop("act_quant_dict", None)
        kwargs.pop("quant_dict", None)
        kwargs.pop("quant_setup", None)
        super().__init__(*args, **kwargs)
        self.method = method
        self.act_method = act_method or method
        self.n_bits = n_bits
        self.n_bits_act = n_bits_act or n_bits
        self.per_channel_weights = per_channel_weights
        self.percentile = percentile
        self.weight_range_method = weight_range_method
        self.weight_range_options = weight_range_options if weight_range_options else {}
        self.act_range_method = act_range_method
        self.act_range_options = act_range_options if act_range_options else {}
        self.scale_domain = scale_domain
        self.bayesian_bits_kwargs = bayesian_bits_kwargs or {}
        self.prune_method = prune_method
        self.prune_kwargs = prune_kwargs
        self.cached_params = None
        self._caching = True
        self.quant_params = None
        self.register_buffer("_quant_w", torch.BoolTensor([False]))
        self.register_buffer("_quant_a", torch.BoolTensor([False]))
        self._qw_host = False
        self._qa_host = False
        self.act_qparams = dict(n_bits=self.n_bits_act, scale_domain=self.scale_domain, act_quant=True,
                                **self.bayesian_bits_kwargs)
        self.weight_qparams = dict(n_bits=self.n_bits, scale_domain=self.scale_domain, act_quant=False,
                                   **self.bayesian_bits_kwargs)

    @property
    def caching(self):
        return self._caching

    @caching.setter
    def caching(self, value: bool):
        self._caching = value
        if not value:
            self.cached_params = None

    def quantized_weights(self):
        self.cached_params = None
        self._quant_w = torch.BoolTensor([True])
        self._qw_host = True

    def full_precision_weights(self):
        self.cached_params = None
        self._quant_w = torch.BoolTensor([False])
        self._qw_host = False

    def quantized_acts(self):
        self._quant_a = torch.BoolTensor([True])
        self._qa_host = True

    def full_precision_acts(self):
        self._quant_a = torch.BoolTensor([False])
        self._qa_host = False

    def quantized(self):
        self.quantized_weights()
        self.quantized_acts()

    def full_precision(self):
        self.full_precision_weights()
        self.full_precision_acts()

    def get_quantizer_status(self):
        return dict(quant_a=self._qa_host, quant_w=self._qw_host)

    def set_quantizer_status(self, quantizer_status):
        (self.quantized_acts if quantizer_status["quant_a"] else self.full_precision_acts)()
        (self.quantized_weights if quantizer_status["quant_w"] else self.full_precision_weights)()

    def fix_ranges(self):
        self.apply(_set_layer_fix_ranges)

    def estimate_ranges(self):
        self.apply(_set_layer_estimate_ranges)

    def estimate_ranges_train(self):
        self.apply(_set_layer_estimate_ranges_train)

    def learn_ranges(self):
        raise NotImplementedError("learn_ranges is outside the eval hot path")

    def set_quant_state(self, weight_quant, act_quant):
        """QuantizedModel.set_quant_state (base_quantized_model.py:146-155) applied to the subtree."""
        def _fn(layer):
            if isinstance(layer, QuantizedModule):
                (layer.quantized_acts if act_quant else layer.full_precision_acts)()
                (layer.quantized_weights if weight_quant else layer.full_precision_weights)()
        self.apply(_fn)

    def train(self, mode=True):
        super().train(mode)
        if mode:
            self.cached_params = None
        return self

    def _apply(self, *args, **kwargs):
        self.cached_params = None
        return super()._apply(*args, **kwargs)

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        super()._load_from_state_dict(state_dict, prefix, *a, **k)
        self._qw_host = bool(self._quant_w.item())
        self._qa_host = bool(self._quant_a.item())

    def extra_repr(self):
        quant_state = "weight_quant={}, act_quant={}".format(self._qw_host, self._qa_host)
        parent_repr = super().extra_repr()
        return "{},\n{}".format(parent_repr, quant_state) if parent_repr else quant_state


class QuantizedActivation(QuantizedModule):
    """An activation tap (base_quantized_classes.py:182-199)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.activation_quantizer = QuantizationManager(qmethod=self.act_method, qparams=self.act_qparams,
                                                        init=self.act_range_method,
                                                        init_params=self.act_range_options)

    def quantize_activations(self, x):
        if self._qa_host:
            return self.activation_quantizer(x)
        return x

    def forward(self, x):
        return self.quantize_activations(x)

    # helpers for the fused attention kernel ------------------------------------------------------------
    @property
    def active(self) -> bool:
        return self._qa_host

    @property
    def fixed(self) -> bool:
        """Quantizing with a frozen range: the tap can be folded into the fused kernel."""
        return self._qa_host and not self.activation_quantizer.estimating

    def kernel_qparams(self) -> torch.Tensor:
        return self.activation_quantizer.quantizer.qparams


class FP32Acts(nn.Module):
    def forward(self, x):
        return x

    def reset_ranges(self):
        pass


class QuantLinear(QuantizedModule, nn.Linear):
    """nn.Linear with a cached fake-quantized weight and an output activation tap
    (autoquant_utils.py:18-20 on top of QuantizationHijacker, hijacker.py:27-134).  The GEMM stays a
    library (cuBLAS) call; the weight quantizer runs once per eval (cached), the output tap runs the CUDA
    fake-quant kernel."""

    def __init__(self, *args, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        if activation is not None and not isinstance(activation, (nn.ReLU, nn.ReLU6, nn.Hardtanh, nn.Sigmoid, nn.Tanh,
                                                                  nn.GELU, nn.PReLU)):
            raise NotImplementedError(f"activation {activation} is not in the hijacker's activations_set (hijacker.py:14-24)")
        # an activation that quantize_sequential fused into the layer runs BEFORE the tap (hijacker.py:115-117)
        self.activation_function = copy.deepcopy(activation) if activation is not None else None
        self.activation_quantizer = QuantizationManager(qmethod=self.act_method, init=self.act_range_method,
                                                        qparams=self.act_qparams, init_params=self.act_range_options)
        if self.weight_range_method is CurrentMinMaxEstimator:
            weight_init_params = dict(percentile=self.percentile)
        else:
            weight_init_params = self.weight_range_options
        self.weight_quantizer = QuantizationManager(qmethod=self.method, init=self.weight_range_method,
                                                    per_channel=self.per_channel_weights, qparams=self.weight_qparams,
                                                    init_params=weight_init_params)

    @classmethod
    def from_linear(cls, lin: nn.Linear, activation=None, **quant_params):
        q = cls(lin.in_features, lin.out_features, bias=lin.bias is not None, activation=activation, **quant_params)
        q.weight = lin.weight
        q.bias = lin.bias
        return q.to(lin.weight.device)

    def get_params(self):
        if not self.training and self.cached_params:
            return self.cached_params
        weight, bias = self.weight, self.bias
        if self._qw_host:
            weight = self.weight_quantizer(weight)
        if self._caching and not self.training and self.cached_params is None:
            self.cached_params = (weight.detach().clone(), bias.detach().clone() if bias is not None else None)
        return weight, bias

    def _fused_gemm_act(self, x, weight):
        """The activation name for ``ops.linear_fwd`` when this call can run on the library's own tcgen05 GEMM (16-bit CUDA
        eval call, no observers on the fused activation), else None."""
        act = self.activation_function
        if (self.training or not x.is_cuda or x.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != x.dtype
                or torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad) or weight.shape[0] % 8
                or x.shape[-1] % 8):
            return None
        if act is None:
            return "none"
        if act._forward_hooks or act._forward_pre_hooks:
            return None
        if isinstance(act, nn.ReLU):
            return "relu"
        if isinstance(act, nn.GELU) and act.approximate == "none":
            return "gelu"
        return None

    def linear_only(self, x):
        """The layer WITHOUT its output tap (cached fake-quantized weight, optional fused activation): for callers that
        fold the tap into the next kernel (``tap_fusable`` / ``tap_qparams``)."""
        weight, bias = self.get_params()
        act = self.activation_function
        if isinstance(act, nn.GELU) and self._fused_gemm_act(x, weight) == "gelu":
            return ops.linear_fwd(x, weight, bias, act="gelu")        # dense -> erf GELU in the GEMM epilogue
        if isinstance(act, nn.ReLU) and bias is not None and x.is_cuda and x.dtype == weight.dtype \
                and not torch.is_grad_enabled() and not act._forward_hooks and not act._forward_pre_hooks:
            # eval: bias + ReLU in the GEMM epilogue (one cuBLASLt call; max(., 0) of the same accumulator is exact)
            x2 = x.contiguous().view(-1, x.shape[-1])
            return torch._addmm_activation(bias, x2, weight.t()).view(*x.shape[:-1], weight.shape[0])
        res = F.linear(x.contiguous(), weight.contiguous(), bias=bias)
        if act is not None:
            res = act(res)
        return res

    @property
    def tap_fusable(self) -> bool:
        """The output tap is off, or quantises with a frozen asymmetric range (can be folded into a consumer kernel)."""
        if not self._qa_host:
            return True
        mgr = self.activation_quantizer
        return isinstance(mgr, QuantizationManager) and not mgr.estimating and \
            isinstance(mgr.quantizer, AsymmetricUniformQuantizer) and not self._forward_hooks

    def tap_qparams(self):
        return self.activation_quantizer.quantizer.qparams if self._qa_host else None

    def forward(self, x, offsets=None):
        if self._qa_host and self.tap_fusable:
            # frozen output tap: F.linear -> [activation] -> activation_quantizer (hijacker.py:78-134) as ONE GEMM launch with
            # the tap in its epilogue.  Measured against cuBLASLt + separate kernels (profiles/r02_gemm_*.jsonl): ahead for
            # wide outputs and whenever an activation is fused; narrow K-heavy layers keep the library GEMM.
            weight, bias = self.get_params()
            act = self._fused_gemm_act(x, weight)
            if act is not None and (act != "none" or weight.shape[0] >= 1536):
                return ops.linear_fwd(x, weight, bias, act=act, qparams=self.tap_qparams(),
                                      n_bits=self.activation_quantizer.n_bits)
        res = self.linear_only(x)
        if self._qa_host:
            res = self.activation_quantizer(res)
        return res
