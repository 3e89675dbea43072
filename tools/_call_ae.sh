#!/bin/bash
# GPU call ae (2 GPUs): N=1 and N=2 (torchrun) bench lines on the same box, reference arm under torchrun, NCCL range-sync check
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-sweep --no-cpu > gpurun_out/ae_bench_n1.json 2> gpurun_out/ae_bench_n1.err; tail -c 600 gpurun_out/ae_bench_n1.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-sweep --no-cpu > gpurun_out/ae_bench_n2.json 2> gpurun_out/ae_bench_n2.err; grep '^{' gpurun_out/ae_bench_n2.json | cut -c1-900; tail -3 gpurun_out/ae_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/ae_bench_ref_n2.json 2> gpurun_out/ae_bench_ref_n2.err; grep '^{' gpurun_out/ae_bench_ref_n2.json | cut -c1-500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_range_sync_nccl.py > gpurun_out/ae_range_sync.json 2> gpurun_out/ae_range_sync.err; tail -c 600 gpurun_out/ae_range_sync.json
