#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_range_sync_nccl.py > gpurun_out/q_range_sync.json 2> gpurun_out/q_range_sync.err; echo "rc=$?"; cat gpurun_out/q_range_sync.json; tail -5 gpurun_out/q_range_sync.err
