"""ctypes front for oracle/oeh_oracle.c (TEST INFRASTRUCTURE ONLY, see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboeh_oracle.so")
_lib = None

_f = C.POINTER(C.c_float)
_i64 = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oeh_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        base = ["make", "-s", "-B", "-C", _HERE]
        for extra in ([], ["OEH_CC=gcc"], ["OEH_CC=cc"], ["OEH_CC=gcc", "OMP="]):   # last resort: single thread
            if subprocess.call(base + extra, stderr=subprocess.DEVNULL) == 0:
                break
        else:
            raise RuntimeError("could not compile oracle/oeh_oracle.c")
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oeh_ref_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_f)


def _strides(a):
    """element strides (b,h,t) of a (B,H,T,d) array whose last dim is contiguous."""
    assert a.strides[-1] == a.itemsize
    return (C.c_int64 * 3)(*[s // a.itemsize for s in a.strides[:3]])


def num_threads() -> int:
    return lib().oeh_ref_num_threads()


def set_threads(n: int) -> None:
    lib().oeh_ref_set_threads(C.c_int(int(n)))


def softmax_rows(x: np.ndarray, kind: int, gamma: float = 0.0, eta: float = 1.0) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().oeh_ref_softmax_rows(_p(x), _p(y), C.c_int64(x.size // x.shape[-1]), C.c_int(x.shape[-1]), C.c_int(kind),
                               C.c_double(gamma), C.c_double(eta))
    return y


def attention(q, k, v, *, scale, kind, gamma=0.0, eta=1.0, causal=False, mask=None, clamp_min=False,
              gate=None, gate_scale=1.0, out_bthd: bool = False) -> np.ndarray:
    """q,k,v (B,H,T,d) fp32 views with contiguous last dim.  Returns (B,H,Tq,d), or (B,Tq,H*d)
    when out_bthd (the permute of bert_attention.py:335-337 folded into the store)."""
    B, H, Tq, d = q.shape
    Tk = k.shape[2]
    if out_bthd:
        base = np.empty((B, Tq, H, d), np.float32)
        out = base.transpose(0, 2, 1, 3)
    else:
        base = out = np.empty((B, H, Tq, d), np.float32)
    ms = None
    if mask is not None:
        mask = np.broadcast_to(np.asarray(mask, np.float32), (B, H, Tq, Tk))
        assert mask.strides[-1] in (4, 0) and (Tk == 1 or mask.strides[-1] == 4)
        ms = (C.c_int64 * 3)(*[s // 4 for s in mask.strides[:3]])
    if gate is not None:
        gate = np.ascontiguousarray(gate, np.float32)
    lib().oeh_ref_attention(_p(q), _p(k), _p(v), _p(out), B, H, Tq, Tk, d, _strides(q), _strides(k), _strides(v),
                            _strides(out), C.c_float(scale), kind, C.c_double(gamma), C.c_double(eta), int(causal),
                            _p(mask) if mask is not None else None, ms, int(clamp_min),
                            _p(gate) if gate is not None else None, C.c_float(gate_scale))
    return base.reshape(B, Tq, H * d) if out_bthd else base


def set_quant_range(x_min, x_max, n_bits=8, eps=1e-8):
    out = np.empty(4, np.float32)
    lib().oeh_ref_set_quant_range(C.c_float(x_min), C.c_float(x_max), n_bits, C.c_float(eps), _p(out))
    return out   # delta, zero_float, scale, zero_point


def fake_quant(x: np.ndarray, scale, zp, n_bits=8, want_codes=False):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    codes = np.empty_like(x) if want_codes else None
    lib().oeh_ref_fake_quant(_p(x), _p(y), _p(codes) if want_codes else None, C.c_int64(x.size), C.c_float(scale),
                             C.c_float(zp), n_bits)
    return (y, codes) if want_codes else y


def minmax(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(2, np.float32)
    lib().oeh_ref_minmax(_p(x), C.c_int64(x.size), _p(out))
    return out


def running_minmax_update(cur2: np.ndarray, new2: np.ndarray, momentum: float, initialized: bool):
    cur2 = np.array(cur2, np.float32)
    new2 = np.ascontiguousarray(new2, np.float32)
    lib().oeh_ref_running_minmax_update(_p(cur2), _p(new2), C.c_double(momentum), int(initialized))
    return cur2


def gate_logits(x: np.ndarray, w: np.ndarray, b: np.ndarray, H: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    B, T, HD = x.shape
    d = HD // H
    out = np.empty((B, H, T), np.float32)
    lib().oeh_ref_gate_logits(_p(x), _p(np.ascontiguousarray(w, np.float32)), _p(np.ascontiguousarray(b, np.float32)),
                              _p(out), B, T, H, d)
    return out
