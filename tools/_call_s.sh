#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quant.py tests/test_add_layer_norm.py tests/test_gpu_next_rows.py tests/test_quantized_model.py tests/test_gpu_modules.py -x -q > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log; tail -25 gpurun_out/s_pytest.log | cut -c1-250
timeout 300 python tools/bench_kernels.py --seqs 128 --kinds 1 --no-causal > gpurun_out/s_kernels.txt 2>&1; grep -v '"T"' gpurun_out/s_kernels.txt | cut -c1-200
timeout 300 python tools/bench_models.py > gpurun_out/s_bench_models.jsonl 2>gpurun_out/s_bench_models.err; cat gpurun_out/s_bench_models.jsonl
