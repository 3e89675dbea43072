#!/bin/bash
# GPU call ag: packed 16-bit fake-quant chain with the reciprocal quotient and magic-number rounding: bit-exactness tests
# (all 65536 patterns vs ATen), quant kernel GB/s, model-level INT8 rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quant.py tests/test_add_layer_norm.py tests/test_gpu_next_rows.py tests/test_quantized_model.py tests/test_gpu_modules.py -x -q -m gpu > gpurun_out/ag_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ag_pytest.log; tail -25 gpurun_out/ag_pytest.log | cut -c1-250
timeout 300 python tools/bench_kernels.py --seqs 128 --kinds 1 --no-causal > gpurun_out/ag_kernels.txt 2>&1; grep -v '"T"' gpurun_out/ag_kernels.txt | cut -c1-200
timeout 300 python tools/bench_models.py > gpurun_out/ag_models.jsonl 2>gpurun_out/ag_models.err; cut -c1-50,100-230 gpurun_out/ag_models.jsonl
