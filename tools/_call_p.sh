#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_quantized_model.py tests/test_gpu_next_rows.py tests/test_add_layer_norm.py tests/test_gpu_modules.py tests/test_gpu_models.py -x -q 2>&1 | tail -3
timeout 600 python tools/bench_models.py > gpurun_out/p_bench_models.jsonl 2> gpurun_out/p_bench_models.err; cat gpurun_out/p_bench_models.jsonl; tail -5 gpurun_out/p_bench_models.err
