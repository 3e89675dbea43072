"""The activation fake-quant sites around the attention core:

* ``QuantLayerNorm``  -- autoquant_utils.py:62-70 on top of QuantizationHijacker (hijacker.py:27-134): LayerNorm with a
  cached fake-quantized weight and an output activation tap.
* ``QuantizedBertSelfOutput`` / ``QuantizedBertOutput`` -- transformers_language/models/quantized_bert.py:540-569,
  :572-605: dense (QuantLinear) -> dropout -> residual add -> ``res_act_quantizer`` -> LayerNorm (QuantLayerNorm).

With fixed ranges the residual add and its tap are ONE kernel (``oeh_add_fake_quant``: 3 tensor passes instead of the
reference's ~14: add = 3, qdq = 6 elementwise ops with temporaries); while the tap is still estimating its range the
sum is materialised and goes through minmax -> EMA -> qdq (the range of the SAME tensor is needed before it is
quantised, quantization_manager.py:94-102).  Once BOTH taps of a residual block are frozen the whole tail
add -> res tap -> LayerNorm -> output tap is one kernel (``oeh_add_layer_norm``); a stand-alone QuantLayerNorm still
calls F.layer_norm.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .quantization import (AsymmetricUniformQuantizer, CurrentMinMaxEstimator, QuantizationManager, QuantizedActivation, QuantizedModule, QuantLinear)
from .quantized_attention import QuantizedModel, _linear_params


class QuantLayerNorm(QuantizedModule, nn.LayerNorm):
    def __init__(self, *args, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        if activation is not None:
            raise NotImplementedError("fused activation functions are not used on this path")
        self.activation_function = None
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
    def from_layer_norm(cls, ln: nn.LayerNorm, **quant_params):
        q = cls(ln.normalized_shape, eps=ln.eps, elementwise_affine=ln.elementwise_affine, **quant_params)
        q.weight, q.bias = ln.weight, ln.bias
        return q.to(ln.weight.device)

    def get_params(self):                                          # hijacker.py:84-101
        if not self.training and self.cached_params:
            return self.cached_params
        weight, bias = self.weight, self.bias
        if self._qw_host:
            weight = self.weight_quantizer(weight)
        if self._caching and not self.training and self.cached_params is None:
            self.cached_params = (weight.detach().clone(), bias.detach().clone() if bias is not None else None)
        return weight, bias

    def forward(self, x, offsets=None):
        weight, bias = self.get_params()
        out = self.activation_quantizer
        fixed = (not self._qa_host) or (not out.estimating and isinstance(out.quantizer, AsymmetricUniformQuantizer))
        if (fixed and not self.training and x.is_cuda and len(self.normalized_shape) == 1 and bias is not None
                and weight.dtype == x.dtype and not self._forward_hooks):
            try:                                    # LayerNorm + frozen output tap in one pass (oeh_add_layer_norm)
                return ops.add_layer_norm(x, None, weight, bias, self.eps,
                                          qparams_out=out.quantizer.qparams if self._qa_host else None,
                                          n_bits=out.n_bits if self._qa_host else 8)
            except NotImplementedError:             # row too wide / unaligned for the register-resident kernel
                pass
        res = F.layer_norm(input=x.contiguous(), normalized_shape=self.normalized_shape, weight=weight.contiguous(),
                           bias=bias.contiguous(), eps=self.eps)
        if self._qa_host:
            res = self.activation_quantizer(res)
        return res


def _residual_tap(tap: QuantizedActivation, hidden_states: torch.Tensor, input_tensor: torch.Tensor,
                  qparams_b: "torch.Tensor | None" = None) -> torch.Tensor:
    """``tap(hidden_states + input_tensor)``: fused when the tap quantises with a frozen range.  ``qparams_b``: frozen
    output tap of the layer that produced ``input_tensor``, folded into the same kernel."""
    if tap.fixed and hidden_states.is_cuda and hidden_states.shape == input_tensor.shape \
            and hidden_states.dtype == input_tensor.dtype:
        return ops.add_fake_quant(hidden_states, input_tensor, tap.kernel_qparams(), tap.activation_quantizer.n_bits,
                                  qparams_b=qparams_b).view(hidden_states.shape)
    if qparams_b is not None:
        raise RuntimeError("a folded producer tap needs a frozen residual tap")
    return tap(hidden_states + input_tensor)


class _QuantizedBertResidualBlock(QuantizedModel):
    def __init__(self, org_model, **quant_params):
        super().__init__()
        self.quant_dict = dict(quant_params.get("quant_dict") or {})
        if self.quant_dict.get("y") == "log" or self.quant_dict.get("layer_norm_res_self_output") \
                or self.quant_dict.get("layer_norm_res_output"):
            raise NotImplementedError("log-domain residual quantizers / tied LayerNorm input quantizers are not used by "
                                      "the reference's eval configs (quant_dict is empty there)")
        qp = _linear_params(quant_params)
        self.dense = QuantLinear.from_linear(org_model.dense, **qp)
        self.dropout = org_model.dropout
        self.res_act_quantizer = QuantizedActivation(**qp)
        self.LayerNorm = QuantLayerNorm.from_layer_norm(org_model.LayerNorm, **qp)

    def _fusable(self, hidden_states, input_tensor) -> bool:
        """Both taps quantise with frozen asymmetric ranges of the same width (or are off): add -> tap -> LayerNorm -> tap
        is ONE kernel (oeh_add_layer_norm)."""
        tap, ln = self.res_act_quantizer, self.LayerNorm
        if self.training or not hidden_states.is_cuda or hidden_states.shape != input_tensor.shape \
                or hidden_states.dtype != input_tensor.dtype or ln.weight.dtype != hidden_states.dtype or ln.bias is None:
            return False
        if tap.active and not tap.fixed:
            return False
        if ln._forward_hooks or ln._forward_pre_hooks or tap._forward_hooks or tap._forward_pre_hooks:
            return False                                            # somebody observes the intermediate tensors
        out = ln.activation_quantizer
        if ln._qa_host and (out.estimating or not isinstance(out.quantizer, AsymmetricUniformQuantizer)):
            return False
        if tap.active and ln._qa_host and tap.activation_quantizer.n_bits != out.n_bits:
            return False
        return True

    def forward(self, hidden_states, input_tensor):
        dense = self.dense
        fold_dense_tap = (not self.training) and dense.tap_fusable       # eval: dropout is the identity
        x = dense.linear_only(hidden_states) if fold_dense_tap else dense(hidden_states)
        x = self.dropout(x)
        if self._fusable(x, input_tensor):
            tap, ln = self.res_act_quantizer, self.LayerNorm
            taps = [m.n_bits for m, on in ((dense.activation_quantizer, fold_dense_tap and dense._qa_host),
                                           (tap.activation_quantizer, tap.active),
                                           (ln.activation_quantizer, ln._qa_host)) if on]
            if len(set(taps)) <= 1:                        # one kernel takes one bit width
                weight, bias = ln.get_params()
                # dense tap -> residual add -> res tap -> LayerNorm -> output tap: ONE kernel
                return ops.add_layer_norm(x, input_tensor, weight, bias, ln.eps,
                                          qparams_a=dense.tap_qparams() if fold_dense_tap else None,
                                          qparams_res=tap.kernel_qparams() if tap.active else None,
                                          qparams_out=ln.activation_quantizer.quantizer.qparams if ln._qa_host else None,
                                          n_bits=taps[0] if taps else 8)
        if fold_dense_tap and dense._qa_host:              # not fusable after all: apply the dense tap on its own
            x = dense.activation_quantizer(x)
        x = _residual_tap(self.res_act_quantizer, x, input_tensor)   # add + tap
        return self.LayerNorm(x)


class QuantizedBertSelfOutput(_QuantizedBertResidualBlock):
    """quantized_bert.py:540-569."""


class QuantizedBertOutput(_QuantizedBertResidualBlock):
    """quantized_bert.py:572-605."""
