"""outeffhop_b200 -- B200 (sm_100a) implementation of OutEffHop's attention + activation
fake-quant hot path behind the reference's own Python module interface.

Public surface (same names as the reference):
    SOFTMAX_MAPPING, AttentionGateType, logit,
    BertSelfAttentionWithExtras, OPTAttentionWithExtras,
    AsymmetricUniformQuantizer, RunningMinMaxEstimator, QuantizationManager, QuantizedActivation,
    QuantizedBertSelfAttentionWithExtras, QuantizedOPTAttentionWithExtras,
    QuantizedBertForMaskedLM, quantize_bert_for_masked_lm, pass_data_for_range_estimation, get_quant_config, val_qparams,
  SURVEY section 8f rows: percentile range estimation (RunningMinMaxEstimator(percentile=...)), QuantLayerNorm /
    QuantizedBertSelfOutput / QuantizedBertOutput, kurtosis / OutlierMeter, ViTSelfAttentionWithExtras,
    Association / Hopfield / HopfieldPooling
All compute goes through liboeh.so (include/oeh.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

import importlib

_SUBMODULES = ("_lib", "ops", "softmax", "quantization", "bert_attention", "opt_attention", "quantized_attention",
               "build", "eval_harness", "vit_attention", "hopfield", "quantized_layers", "outlier_metrics", "quantized_models",
               "hf_compat")
_EXPORTS = {
    "SOFTMAX_MAPPING": "softmax", "SoftmaxFn": "softmax", "softmax_1": "softmax", "clipped_softmax": "softmax",
    "clipped_softmax1": "softmax", "parse_attn_softmax": "softmax",
    "AttentionGateType": "bert_attention", "logit": "bert_attention",
    "BertSelfAttentionWithExtras": "bert_attention", "OPTAttentionWithExtras": "opt_attention",
    "AsymmetricUniformQuantizer": "quantization", "SymmetricUniformQuantizer": "quantization",
    "RunningMinMaxEstimator": "quantization", "CurrentMinMaxEstimator": "quantization",
    "QuantizationManager": "quantization", "QuantizedActivation": "quantization", "QuantizedModule": "quantization",
    "QuantLinear": "quantization", "QMethods": "quantization", "RangeEstimators": "quantization",
    "Qstates": "quantization", "QuantizerNotInitializedError": "quantization", "set_range_sync": "quantization",
    "QuantizedBertSelfAttentionWithExtras": "quantized_attention",
    "QuantizedOPTAttentionWithExtras": "quantized_attention", "QuantizedModel": "quantized_attention",
    "ViTSelfAttentionWithExtras": "vit_attention",
    "Association": "hopfield", "Hopfield": "hopfield", "HopfieldPooling": "hopfield",
    "QuantLayerNorm": "quantized_layers", "QuantizedBertSelfOutput": "quantized_layers",
    "QuantizedBertOutput": "quantized_layers",
    "QuantizedBertForMaskedLM": "quantized_models", "QuantizedBertModel": "quantized_models",
    "QuantizedBertLayer": "quantized_models", "QuantizedBertEmbeddings": "quantized_models",
    "QuantEmbedding": "quantized_models", "quantize_bert_for_masked_lm": "quantized_models",
    "QuantizedOPTForCausalLM": "quantized_models", "QuantizedOPTDecoder": "quantized_models",
    "QuantizedOPTDecoderLayer": "quantized_models", "quantize_opt_for_causal_lm": "quantized_models",
    "pass_data_for_range_estimation": "quantized_models", "get_quant_config": "quantized_models",
    "val_qparams": "quantized_models", "DotDict": "quantized_models",
    "kurtosis": "outlier_metrics", "OutlierMeter": "outlier_metrics", "attach_bert_outlier_hooks": "outlier_metrics",
}


def __getattr__(name):
    # lazy: `import outeffhop_b200` must stay cheap (no torch import) until a symbol is used
    if name in _SUBMODULES:
        return importlib.import_module(f"{__name__}.{name}")
    if name in _EXPORTS:
        return getattr(importlib.import_module(f"{__name__}.{_EXPORTS[name]}"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + list(_SUBMODULES) + list(_EXPORTS))
