#!/usr/bin/env python
"""One warm eager fixed-range INT8 BERT-base eval step between cudaProfilerStart/Stop (for an ncu launch list):
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file x.csv python tools/prof_int8_step.py [--opt]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import eval_harness as EH, quantized_models as QM  # noqa: E402

opt = "--opt" in sys.argv
if opt:
    model = EH.build_opt_125m(seed=6789, attn_softmax="softmax1")
    q = QM.quantize_opt_for_causal_lm(model)
    host = [EH.synthetic_clm_batch(4, 512, seed=i) for i in range(5)]
    step = EH.make_opt_int8_eval_step(q, host[0], host[:4], use_graph=False)
else:
    model = EH.build_bert_base(seed=4000)
    q = QM.quantize_bert_for_masked_lm(model)
    host = [EH.synthetic_mlm_batch(32, 128, seed=i) for i in range(5)]
    step = EH.make_bert_int8_eval_step(q, host[0], host[:4], use_graph=False)
step.load(host[4])
for _ in range(3):
    step.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(step.loss))
