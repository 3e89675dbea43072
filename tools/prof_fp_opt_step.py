#!/usr/bin/env python
"""One warm eager step of BASELINE config 3 (OPT-125m fp16 causal, per-token gate) between cudaProfilerStart/Stop."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import eval_harness as EH
model = EH.build_opt_125m(seed=5678, attn_softmax="vanilla", attn_gate_type="conditional_per_token", attn_gate_init=0.25)
host = [EH.synthetic_clm_batch(4, 512, seed=i) for i in range(2)]
step = EH.make_opt_eval_step(model, host[0], use_graph=False)
step.load(host[1])
for _ in range(3):
    step.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(step.loss))
