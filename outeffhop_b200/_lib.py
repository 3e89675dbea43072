"""ctypes binding of liboeh.so (include/oeh.h).  There is NO fallback: if the CUDA library is
missing or cannot be loaded, every product entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OEH_LIB_PATH: developer override used by kernel experiments (tools/); the product always loads lib/liboeh.so
LIB_PATH = os.environ.get("OEH_LIB_PATH") or os.path.join(_HERE, "lib", "liboeh.so")

OEH_OK, OEH_ERR_BAD_ARG, OEH_ERR_UNSUPPORTED, OEH_ERR_CUDA, OEH_ERR_WORKSPACE, OEH_ERR_NOT_INITIALIZED = range(6)
OEH_F32, OEH_F16, OEH_BF16 = 0, 1, 2
KIND_VANILLA, KIND_SOFTMAX1, KIND_CLIPPED, KIND_CLIPPED1 = 0, 1, 2, 3
QS_CUR_MIN, QS_CUR_MAX, QS_DELTA, QS_ZERO_FLOAT, QS_SCALE, QS_ZERO_POINT, QS_INITIALIZED, QS_RESERVED = range(8)
QS_FLOATS = 8


class OehAttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_strides", C.c_int64 * 3), ("k_strides", C.c_int64 * 3), ("v_strides", C.c_int64 * 3),
        ("out_strides", C.c_int64 * 3),
        ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32), ("d", C.c_int32),
        ("dtype", C.c_int32), ("scale", C.c_float), ("kind", C.c_int32),
        ("gamma", C.c_double), ("eta", C.c_double),
        ("causal", C.c_int32),
        ("mask", C.c_void_p), ("mask_dtype", C.c_int32), ("mask_strides", C.c_int64 * 3),
        ("clamp_min", C.c_int32),
        ("gate_logits", C.c_void_p), ("gate_scale", C.c_float),
        ("qdq_scores", C.c_void_p), ("qdq_probs", C.c_void_p), ("qdq_ctx", C.c_void_p),
        ("n_bits", C.c_int32), ("ctx_qdq_before_gate", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("stream", C.c_void_p),
    ]


class OehLinearArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("y", C.c_void_p),
        ("ldx", C.c_int64), ("ldw", C.c_int64), ("ldy", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("dtype", C.c_int32), ("act", C.c_int32),
        ("n_seg", C.c_int32), ("n_bits", C.c_int32), ("tile_n", C.c_int32),
        ("qparams", C.c_void_p * 4),
        ("stream", C.c_void_p),
    ]


class OehLinearLnArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p), ("y", C.c_void_p), ("sum_out", C.c_void_p),
        ("ldx", C.c_int64), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("dtype", C.c_int32), ("n_bits", C.c_int32),
        ("eps", C.c_double),
        ("q_a", C.c_void_p), ("q_res", C.c_void_p), ("q_out", C.c_void_p),
        ("stream", C.c_void_p),
    ]


ACT_NONE, ACT_RELU, ACT_GELU_ERF = 0, 1, 2


class OehUnsupported(NotImplementedError):
    """The C layer declined this shape/dtype/branch (OEH_ERR_UNSUPPORTED)."""


class QuantizerNotInitializedError(Exception):
    """Raised when a quantizer has not initialized (quantization/quantizers/quantizer_utils.py:31-37)."""

    def __init__(self):
        super().__init__("Quantizer has  not been initialized yet")


# every symbol include/oeh.h declares: (name, restype, argtypes)
_I64P = C.POINTER(C.c_int64)
SYMBOLS = {
    "oeh_version": (C.c_int, []),
    "oeh_last_error": (C.c_char_p, []),
    "oeh_device_supported": (C.c_int, []),
    "oeh_attn_workspace_bytes": (C.c_size_t, [C.POINTER(OehAttnArgs)]),
    "oeh_attn_fwd": (C.c_int, [C.POINTER(OehAttnArgs)]),
    "oeh_softmax_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_float, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_int32, _I64P,
                                   C.c_int32, C.c_int32, C.c_void_p]),
    "oeh_gate_logits": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "oeh_fake_quant_asym": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p]),
    "oeh_minmax_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "oeh_minmax": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "oeh_range_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_double, C.c_void_p]),
    "oeh_set_quant_range_host": (C.c_int, [C.c_float, C.c_float, C.c_int32, C.c_double, C.POINTER(C.c_float)]),
    "oeh_estimate_and_quantize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_double,
                                            C.c_int32, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p]),
    "oeh_percentile_workspace_bytes": (C.c_size_t, []),
    "oeh_percentile_passes": (C.c_int, [C.c_int32]),
    "oeh_percentile_plan": (C.c_int, [C.c_int64, C.c_double, C.c_double, _I64P, C.POINTER(C.c_double)]),
    "oeh_percentile_begin": (C.c_int, [C.c_void_p, C.c_size_t, _I64P, C.c_void_p]),
    "oeh_percentile_hist": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "oeh_percentile_scan": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "oeh_percentile_finish": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p]),
    "oeh_percentile": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "oeh_range_update_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_double,
                                       C.c_void_p]),
    "oeh_add_fake_quant": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_void_p]),
    "oeh_outlier_stats_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "oeh_outlier_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "oeh_ce_loss_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "oeh_ce_loss": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "oeh_ce_loss_shifted": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "oeh_embed_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "oeh_add_layer_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_void_p]),
    "oeh_add_layer_norm_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                          C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "oeh_tap_add_fake_quant": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_int32, C.c_void_p]),
    "oeh_fake_quant_asym_segments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64,
                                               C.c_int64, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "oeh_gelu_erf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "oeh_linear_fwd": (C.c_int, [C.POINTER(OehLinearArgs)]),
    "oeh_linear_add_layer_norm": (C.c_int, [C.POINTER(OehLinearLnArgs)]),
    "oeh_linear_add_layer_norm_supported": (C.c_int, [C.c_int32]),
    "oeh_embed_layer_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_double,
                                       C.c_int32, C.c_void_p]),
}
# test-only entry point (not part of the public header)
_PRIVATE = {
    "oeh_attn_fwd_debug": (C.c_int, [C.POINTER(OehAttnArgs), C.c_void_p, C.c_void_p]),
    "oeh_set_override": (C.c_int, [C.c_char_p, C.c_longlong]),
    "oeh_sm_clock_probe": (C.c_int, [C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library.  Raises (loudly) when it is absent: build it with
    `python -m outeffhop_b200.build` or `__graft_entry__.build()`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(python -m outeffhop_b200.build). There is no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in {**SYMBOLS, **_PRIVATE}.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.oeh_version() != 100:
            raise RuntimeError(f"liboeh.so version {handle.oeh_version()} does not match the Python shim (100)")
        _lib = handle
    return _lib


def set_override(name: str, value) -> None:
    """Dispatch override of the attention entry point (tests / tuning): ``OEH_ATTN_KERNEL`` ("tile" / "ws"),
    ``OEH_WS_PERSISTENT``, ``OEH_WS_TILES``, ``OEH_CAUSAL_CHUNK``, ``OEH_ONE_REGS``, ``OEH_MID_KERNEL``, ``OEH_ONE_SPLIT``; ``None`` clears it.  Process wide,
    thread-safe; the library reads the environment variables of the same names once, at its first call."""
    if value is None:
        v = -1
    elif name == "OEH_ATTN_KERNEL":
        v = {"tile": 0, "ws": 1}[value]
    else:
        v = int(value)
    check(lib().oeh_set_override(name.encode(), v))


def sm_clock_probe(out, stream: int) -> None:
    """Measurement helper: queue a ~20 us single-thread kernel on ``stream`` that writes the SM clock (MHz, clock64 against
    %globaltimer) to the one-float device tensor ``out`` -- the clock the kernel queued just before it ran at."""
    check(lib().oeh_sm_clock_probe(out.data_ptr(), stream))


def last_error() -> str:
    return lib().oeh_last_error().decode("utf-8", "replace")


def check(rc: int):
    """Map a C status to the reference's Python error conventions."""
    if rc == OEH_OK:
        return
    msg = last_error()
    if rc == OEH_ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == OEH_ERR_UNSUPPORTED:
        raise OehUnsupported(msg)
    if rc == OEH_ERR_NOT_INITIALIZED:
        raise QuantizerNotInitializedError()
    raise RuntimeError(f"liboeh status {rc}: {msg}")
