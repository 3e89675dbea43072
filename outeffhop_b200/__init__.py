"""outeffhop_b200 -- B200 (sm_100a) implementation of OutEffHop's attention + activation
fake-quant hot path behind the reference's own Python module interface.

Public surface (same names as the reference):
    SOFTMAX_MAPPING, AttentionGateType, logit,
    BertSelfAttentionWithExtras, OPTAttentionWithExtras,
    AsymmetricUniformQuantizer, RunningMinMaxEstimator, QuantizationManager, QuantizedActivation,
    QuantizedBertSelfAttentionWithExtras, QuantizedOPTAttentionWithExtras
All compute goes through liboeh.so (include/oeh.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def __getattr__(name):
    # lazy: importing the package must not import torch-heavy modules before they are needed
    import importlib
    for mod in ("softmax", "bert_attention", "opt_attention", "quantization", "quantized_attention", "ops"):
        try:
            m = importlib.import_module(f"{__name__}.{mod}")
        except ModuleNotFoundError:
            continue
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError(name)
