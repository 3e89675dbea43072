#!/usr/bin/env python
"""Batch-sharded range estimation over NCCL (BASELINE config 4): run under torchrun with N >= 2 GPUs.

Every rank builds the same small INT8 BERT, estimates activation ranges on ITS shard of the estimation batches with the
MIN/MAX all-reduce enabled (quantization.set_range_sync), and the script checks that
  (1) every quantizer state is bit-identical on all ranks, and
  (2) it equals (<= 1e-3 relative: GEMM kernels may differ with the batch size) a single-process estimation at the
      global batch (the shards of a step concatenated), computed on rank 0 with the sync off.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_range_sync_nccl.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import eval_harness as EH, quantization as Q, quantized_models as QM  # noqa: E402


def states(qmodel):
    out = {}
    for n, m in qmodel.named_modules():
        if isinstance(m, Q.QuantizationManager) and m.quantizer.is_initialized and isinstance(m.quantizer, Q.AsymmetricUniformQuantizer):
            out[n] = torch.stack((m.quantizer._delta.float().reshape(()), m.quantizer._zero_float.float().reshape(())))
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    steps, b = 3, 4
    batches = [[EH.synthetic_mlm_batch(b, 128, seed=100 * s + r, pin=False) for r in range(world)] for s in range(steps)]

    model = EH.build_bert_base(seed=4000, dtype=torch.float32, num_hidden_layers=2, attn_softmax="softmax1")
    q = QM.quantize_bert_for_masked_lm(model)
    Q.set_range_sync(True)
    QM.pass_data_for_range_estimation([batches[s][rank] for s in range(steps)], q, act_quant=True, max_num_batches=steps)
    Q.set_range_sync(False)
    q.fix_ranges()
    mine = states(q)
    names = sorted(mine)
    flat = torch.stack([mine[n] for n in names]).contiguous()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], g) for g in gathered)

    result = {"world": world, "quantizers": len(names), "identical_across_ranks": bool(identical)}
    if rank == 0:
        model1 = EH.build_bert_base(seed=4000, dtype=torch.float32, num_hidden_layers=2, attn_softmax="softmax1")
        q1 = QM.quantize_bert_for_masked_lm(model1)
        glob = [{k: torch.cat([batches[s][r][k] for r in range(world)]) for k in ("input_ids", "labels")} for s in range(steps)]
        QM.pass_data_for_range_estimation(glob, q1, act_quant=True, max_num_batches=steps)
        q1.fix_ranges()
        ref = states(q1)
        rel = max(float((mine[n][0] - ref[n][0]).abs() / ref[n][0].abs()) for n in names)
        exact = sum(bool(torch.equal(mine[n], ref[n])) for n in names)
        result.update({"max_rel_delta_vs_single_process_global_batch": rel, "bit_identical_vs_single_process": exact,
                       "pass": bool(identical and rel <= 1e-3)})
        print(json.dumps(result), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not result["pass"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
