#!/usr/bin/env python
"""Model-level throughput of the BASELINE configs other than the headline (run on the GPU box):
  config 3  OPT-125m fp16 causal, vanilla softmax + conditional_per_token gate (init 0.25), block 512, batch 4
  config 4  OPT-125m INT8 (--quantize, fp32_head, running_minmax, asymmetric 8-bit activations) with softmax1, block 512,
            batch 4: fixed-range eval step (fp16 activations; the bit-exact range contract is tested in fp32)
  (extra)   BERT-base INT8, clippedsoftmax1(-.025:1), 32 x 128
Each step = forward + loss, CUDA-graph replay, inputs copied device-to-device per step; CUDA-event timed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import eval_harness as EH, quantized_models as QM  # noqa: E402


def timed(step, pool, steps=50, warmup=5):
    def body(n):
        for i in range(n):
            for k, v in pool[i % len(pool)].items():
                step.static_inputs[k].copy_(v, non_blocking=True)
            step.run()
    body(warmup)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    body(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    out = []
    # config 3
    model = EH.build_opt_125m(seed=5678, attn_softmax="vanilla", attn_gate_type="conditional_per_token", attn_gate_init=0.25)
    host = [EH.synthetic_clm_batch(4, 512, seed=5678 + i) for i in range(8)]
    pool = [{k: v.cuda() for k, v in b.items()} for b in host]
    step = EH.make_opt_eval_step(model, host[0])
    ms = timed(step, pool)
    out.append(dict(config="3: OPT-125m fp16 causal gated (vanilla, conditional_per_token 0.25), 4x512", ms_per_step=ms,
                    tokens_per_s=4 * 512 / ms * 1e3, loss=float(step.loss)))
    # config 4
    model = EH.build_opt_125m(seed=6789, attn_softmax="softmax1")
    q = QM.quantize_opt_for_causal_lm(model, quant_setup="fp32_head")
    host = [EH.synthetic_clm_batch(4, 512, seed=6789 + i) for i in range(8)]
    pool = [{k: v.cuda() for k, v in b.items()} for b in host]
    step = EH.make_opt_int8_eval_step(q, host[0], host[:4])
    ms = timed(step, pool)
    out.append(dict(config="4: OPT-125m INT8 (fp32_head, running_minmax, W8A8) softmax1, 4x512, fixed ranges after 4 estimation batches",
                    ms_per_step=ms, tokens_per_s=4 * 512 / ms * 1e3, loss=float(step.loss)))
    # BERT INT8
    model = EH.build_bert_base(seed=4000)
    q = QM.quantize_bert_for_masked_lm(model)
    host = [EH.synthetic_mlm_batch(32, 128, seed=4000 + i) for i in range(8)]
    pool = [{k: v.cuda() for k, v in b.items()} for b in host]
    step = EH.make_bert_int8_eval_step(q, host[0], host[:4])
    ms = timed(step, pool)
    out.append(dict(config="BERT-base INT8 (W8A8, FP32 head) clippedsoftmax1(-.025:1), 32x128, fixed ranges after 4 estimation batches",
                    ms_per_step=ms, tokens_per_s=32 * 128 / ms * 1e3, loss=float(step.loss)))
    # config 1 shape on the GPU: BERT-base fp32 (the reference's eval precision), softmax1, 32 x 128 -- exact unfused
    # attention (library GEMMs + oeh_softmax_rows) vs the fused kernel with fp16 operands (fused_fp32)
    for fused in (False, True):
        from outeffhop_b200 import hf_compat
        from transformers import BertConfig, BertForMaskedLM
        cfg = BertConfig()
        cfg._attn_implementation = "eager"
        torch.manual_seed(3000)
        model = BertForMaskedLM(cfg).to("cuda").eval()
        model = hf_compat.fuse_bert_gelu(hf_compat.fuse_bert_residual_blocks(
            hf_compat.swap_bert_attention(model, attn_softmax="softmax1", max_seq_length=128, fused_fp32=fused)))
        host = [EH.synthetic_mlm_batch(32, 128, seed=3000 + i) for i in range(8)]
        pool = [{k: v.cuda() for k, v in b.items()} for b in host]
        step = EH.make_bert_eval_step(model, host[0])
        ms = timed(step, pool)
        out.append(dict(config=f"1 (GPU): BERT-base fp32 softmax1 32x128, attention {'fused, fp16 operands' if fused else 'unfused, exact fp32'}",
                        ms_per_step=ms, tokens_per_s=32 * 128 / ms * 1e3, loss=float(step.loss)))
        del model, step
    for r in out:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
