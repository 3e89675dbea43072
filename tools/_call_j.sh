#!/bin/bash
# GPU call j (2 GPUs): torchrun bench at N=2, plus N=1 bench / reference arm / full tests on the final library
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/j_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/j_bench_n2.json 2> gpurun_out/j_bench_n2.err; echo "n2 rc=$?"; tail -c 400 gpurun_out/j_bench_n2.json; tail -3 gpurun_out/j_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/j_bench_n1.json 2> gpurun_out/j_bench_n1.err; echo "n1 rc=$?"; tail -c 2200 gpurun_out/j_bench_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/j_bench_ref_n2.json 2> gpurun_out/j_bench_ref_n2.err; echo "ref rc=$?"; cat gpurun_out/j_bench_ref_n2.json | tail -c 600
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log; tail -4 gpurun_out/j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; tail -2 gpurun_out/j_smoke.log
