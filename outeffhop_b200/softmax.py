"""``SOFTMAX_MAPPING`` -- the reference's CLI-string -> softmax-function table
(transformers_language/models/softmax.py:22-64), re-hosted on the CUDA row-softmax kernel.

Every value is a :class:`SoftmaxFn`: a callable ``fn(data, dim=-1, **kw)`` exactly like the reference's
functions / partials, that additionally exposes ``(kind, gamma, eta)`` so the attention modules can
hand the whole ``QK^T -> fn -> PV`` chain to the fused tcgen05 kernel instead of calling it.

The 40 literal keys are reproduced, including the two table quirks (``"clipped(-.005:1.005)"`` has
gamma=-0.003, softmax.py:57; ``"clippedsoftmax1(-.025:1)"`` has eta=1.1, softmax.py:61).
``parse_attn_softmax`` additionally accepts any ``clipped(a:b)`` / ``clippedsoftmax1(a:b)`` string
(gamma=a, eta=b); for table keys it returns the table's values.
"""
from __future__ import annotations

import re
from typing import Optional

import torch

from . import _lib as L
from . import ops

KIND_VANILLA, KIND_SOFTMAX1, KIND_CLIPPED, KIND_CLIPPED1 = (L.KIND_VANILLA, L.KIND_SOFTMAX1, L.KIND_CLIPPED,
                                                             L.KIND_CLIPPED1)


class SoftmaxFn:
    """softmax / softmax_1 / clipped softmax / clipped softmax_1 as one parametrised callable."""

    __slots__ = ("kind", "gamma", "eta", "__name__")

    def __init__(self, kind: int, gamma: float = 0.0, eta: float = 1.0, name: Optional[str] = None):
        self.kind, self.gamma, self.eta = int(kind), float(gamma), float(eta)
        self.__name__ = name or ("softmax", "softmax_1", "clipped_softmax", "clipped_softmax1")[self.kind]

    @property
    def keywords(self):   # functools.partial look-alike (the reference's clipped entries are partials)
        return {"gamma": self.gamma, "eta": self.eta} if self.kind >= KIND_CLIPPED else {}

    def __call__(self, data: torch.Tensor, dim: int = -1, dtype: Optional[torch.dtype] = None, **kw) -> torch.Tensor:
        """``dtype`` casts the input first, like F.softmax(..., dtype=).  The reference's softmax_1 has no
        ``dtype`` kwarg (vutils/softmax_1.py:24) so OPT fp16 + softmax1 raises TypeError there
        (opt_attention.py:228); here the kwarg is honoured for all kinds."""
        if kw:
            raise TypeError(f"{self.__name__}() got unexpected keyword arguments {sorted(kw)}")
        x = data if dtype is None else data.to(dtype)
        nd = x.dim()
        d = dim % nd
        if d != nd - 1:
            x = x.transpose(d, nd - 1)
        y = ops.softmax_rows(x, self.kind, self.gamma, self.eta)
        y = y.view(x.shape)
        if d != nd - 1:
            y = y.transpose(d, nd - 1)
        return y

    def __repr__(self):
        return f"SoftmaxFn({self.__name__}, gamma={self.gamma}, eta={self.eta})"

    def __eq__(self, other):
        return isinstance(other, SoftmaxFn) and (self.kind, self.gamma, self.eta) == (other.kind, other.gamma, other.eta)

    def __hash__(self):
        return hash((self.kind, self.gamma, self.eta))


def softmax_1(input: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """exp(x_i) / (1 + sum_j exp(x_j))  (vutils/softmax_1.py:24-28)."""
    return SoftmaxFn(KIND_SOFTMAX1)(input, dim=dim)


def clipped_softmax(data, dim=1, eta=1.1, gamma=-0.1, **kw):
    """transformers_language/models/softmax.py:10-13 (same defaults, including dim=1)."""
    return SoftmaxFn(KIND_CLIPPED, gamma, eta)(data, dim=dim, **kw)


def clipped_softmax1(data, dim=1, eta=1.1, gamma=-0.1, **kw):
    """transformers_language/models/softmax.py:16-19."""
    return SoftmaxFn(KIND_CLIPPED1, gamma, eta)(data, dim=dim, **kw)


def entmax15(X: torch.Tensor, dim: int = -1, k=None) -> torch.Tensor:
    """1.5-entmax (vutils/entmax.py:160-235).  Outside the accelerated hot path: the
    key stays selectable, evaluated with plain sort-based tensor ops on the device."""
    x = X / 2
    x = x - x.max(dim=dim, keepdim=True).values
    xs, _ = torch.sort(x, dim=dim, descending=True)
    n = x.shape[dim]
    shape = [1] * x.dim()
    shape[dim] = n
    rho = torch.arange(1, n + 1, device=x.device, dtype=x.dtype).view(shape)
    mean = xs.cumsum(dim) / rho
    mean_sq = (xs ** 2).cumsum(dim) / rho
    ss = rho * (mean_sq - mean ** 2)
    delta = (1 - ss) / rho
    tau = mean - torch.sqrt(torch.clamp(delta, min=0))
    support = (tau <= xs).sum(dim=dim, keepdim=True)
    tau_star = tau.gather(dim, support - 1)
    return torch.clamp(x - tau_star, min=0) ** 2


def _build_mapping():
    m = {
        "vanilla": SoftmaxFn(KIND_VANILLA, name="softmax"),                                      # softmax.py:23
        "softmax1": SoftmaxFn(KIND_SOFTMAX1),                                                    # :24
        "entmax": entmax15,                                                                      # :25
    }
    for eta in ("1.0003", "1.001", "1.002", "1.003", "1.004", "1.01", "1.02", "1.03", "1.1"):   # :27-35
        m[f"clipped(0:{eta})"] = SoftmaxFn(KIND_CLIPPED, 0, float(eta))
    for g in (".1", ".00001", ".00003", ".0001", ".0003", ".0005", ".001", ".002", ".0025", ".003", ".004", ".005",
              ".01", ".015", ".02", ".025", ".03", ".04"):                                       # :36-53
        m[f"clipped(-{g}:1)"] = SoftmaxFn(KIND_CLIPPED, -float("0" + g), 1.0)
    m["clipped(-.001:1.001)"] = SoftmaxFn(KIND_CLIPPED, -0.001, 1.001)                           # :54
    m["clipped(-.002:1.002)"] = SoftmaxFn(KIND_CLIPPED, -0.002, 1.002)                           # :55
    m["clipped(-.003:1.003)"] = SoftmaxFn(KIND_CLIPPED, -0.003, 1.003)                           # :56
    m["clipped(-.005:1.005)"] = SoftmaxFn(KIND_CLIPPED, -0.003, 1.005)                           # :57 (table quirk)
    m["clipped(-.01:1.01)"] = SoftmaxFn(KIND_CLIPPED, -0.01, 1.01)                               # :58
    m["clipped(-.03:1.03)"] = SoftmaxFn(KIND_CLIPPED, -0.03, 1.03)                               # :59
    m["clipped(-.1:1.1)"] = SoftmaxFn(KIND_CLIPPED, -0.1, 1.1)                                   # :60
    m["clippedsoftmax1(-.025:1)"] = SoftmaxFn(KIND_CLIPPED1, -0.025, 1.1)                        # :61 (table quirk)
    m["clippedsoftmax1(-.00001:1)"] = SoftmaxFn(KIND_CLIPPED1, -0.00001, 1.0)                    # :62
    m["clippedsoftmax1(-.0001:1)"] = SoftmaxFn(KIND_CLIPPED1, -0.0001, 1.0)                      # :63
    return m


SOFTMAX_MAPPING = _build_mapping()

_PAT = re.compile(r"^(clipped|clippedsoftmax1)\(\s*([-+.\deE]+)\s*:\s*([-+.\deE]+)\s*\)$")


def parse_attn_softmax(key: str):
    """``--attn_softmax`` string -> callable.  Table keys return the table entry; any other
    ``clipped(a:b)`` / ``clippedsoftmax1(a:b)`` is parsed as gamma=a, eta=b."""
    if key in SOFTMAX_MAPPING:
        return SOFTMAX_MAPPING[key]
    mt = _PAT.match(key)
    if not mt:
        raise KeyError(f"unknown --attn_softmax value {key!r}")
    kind = KIND_CLIPPED if mt.group(1) == "clipped" else KIND_CLIPPED1
    return SoftmaxFn(kind, float(mt.group(2)), float(mt.group(3)))


def as_softmax_fn(fn) -> Optional[SoftmaxFn]:
    """Recognise a softmax callable the fused kernel can evaluate; None -> call it on materialised scores."""
    if isinstance(fn, SoftmaxFn):
        return fn
    if fn is torch.nn.functional.softmax or fn is torch.softmax:
        return SOFTMAX_MAPPING["vanilla"]
    if fn is softmax_1:
        return SOFTMAX_MAPPING["softmax1"]
    func = getattr(fn, "func", None)     # functools.partial(clipped_softmax, gamma=..., eta=...)
    if func in (clipped_softmax, clipped_softmax1):
        kw = dict(eta=1.1, gamma=-0.1)
        kw.update(getattr(fn, "keywords", {}) or {})
        return SoftmaxFn(KIND_CLIPPED if func is clipped_softmax else KIND_CLIPPED1, kw["gamma"], kw["eta"])
    return None
