"""Outlier metrics of the FP validate loops on the device.

The reference hooks the FFN-output dense and LayerNorm of every layer, and for each batch calls ``.item()`` on every
sample's inf-norm and kurtosis (validate_mlm_config.py:499-540; validate_clm.py:565-586;
``kurtosis`` at transformers_language/utils.py:9-20) -- 32 x 36 host syncs per BERT batch.  Here one
``oeh_outlier_stats`` launch pair per tensor writes the per-sample values and folds them into a device-resident
AverageMeter; nothing is read back until ``summary()``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

from . import ops


def kurtosis(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """x - (B, d) -> (B,) kurtosis (transformers_language/utils.py:9-20)."""
    return ops.outlier_stats(x.reshape(x.size(0), -1), eps)[1].to(x.dtype if x.dtype != torch.float64 else torch.float32)


class OutlierMeter:
    """AverageMeters ``act_inf_norms[name]`` / ``act_kurtoses[name]`` of the validate loops, kept on the device as
    float64 [sum inf-norm, sum kurtosis, samples] per tensor name."""

    def __init__(self, kurtosis_batches: int = 256):
        self.meters: Dict[str, torch.Tensor] = OrderedDict()
        self.inf_only: Dict[str, torch.Tensor] = OrderedDict()
        self.kurtosis_batches = kurtosis_batches      # the reference stops updating kurtosis after batch 256 (:521)

    def update(self, name: str, x: torch.Tensor, batch_idx: int = 0, with_kurtosis: bool = True):
        store = self.meters if with_kurtosis and batch_idx <= self.kurtosis_batches else self.inf_only
        m = store.get(name)
        if m is None:
            m = store[name] = torch.zeros(3, dtype=torch.float64, device=x.device)
        ops.outlier_stats(x.reshape(x.size(0), -1), meter=m)

    def summary(self) -> Dict[str, float]:
        """One read-back: {name: avg inf-norm}, plus the reference's aggregate metrics (:553-575)."""
        inf: Dict[str, float] = OrderedDict()
        kurt: Dict[str, float] = OrderedDict()
        names = list(OrderedDict.fromkeys(list(self.meters) + list(self.inf_only)))
        for name in names:
            tot = torch.zeros(3, dtype=torch.float64)
            k = None
            if name in self.meters:
                t = self.meters[name].cpu()
                tot += t
                k = float(t[1] / t[2])
            if name in self.inf_only:
                tot += self.inf_only[name].cpu()
            inf[name] = float(tot[0] / tot[2])
            if k is not None:
                kurt[name] = k
        out = OrderedDict(inf)
        dense = [v for k2, v in inf.items() if "dense" in k2]
        ln_out = [v for k2, v in inf.items() if k2.endswith("LayerNorm")]
        ln_in = [v for k2, v in inf.items() if "input" in k2]
        if dense:
            out["max_ffn_out_inf_norm"] = max(dense)
        if ln_out:
            out["max_LN_out_inf_norm"] = max(ln_out)
        if ln_in:
            out["max_LN_inp_inf_norm"] = max(ln_in)
        if kurt:
            out["avg_kurtosis"] = sum(kurt.values()) / len(kurt)
            out["max_kurtosis"] = max(kurt.values())
        return out


def attach_bert_outlier_hooks(model, meter: OutlierMeter):
    """Forward hooks on ``bert.encoder.layer.{j}.output.dense`` and ``.output.LayerNorm`` (+ the LayerNorm input),
    the tensors of validate_mlm_config.py:501-540.  Returns the hook handles."""
    handles = []
    state = {"batch_idx": 0}

    def mk(name, with_input):
        def hook(_mod, inp, out):
            meter.update(name, out, state["batch_idx"])
            if with_input:
                meter.update(name + ".input", inp[0], state["batch_idx"], with_kurtosis=False)
        return hook

    for j, layer in enumerate(model.bert.encoder.layer):
        handles.append(layer.output.dense.register_forward_hook(mk(f"bert.encoder.layer.{j}.output.dense", False)))
        handles.append(layer.output.LayerNorm.register_forward_hook(mk(f"bert.encoder.layer.{j}.output.LayerNorm", True)))
    meter._state = state
    return handles
