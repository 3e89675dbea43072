#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_add_layer_norm.py tests/test_gpu_models.py tests/test_quantized_model.py -x -q 2>&1 | tail -6
timeout 300 python tools/bench_models.py > gpurun_out/t_bench_models.jsonl 2>gpurun_out/t_bench_models.err; cat gpurun_out/t_bench_models.jsonl; tail -3 gpurun_out/t_bench_models.err
