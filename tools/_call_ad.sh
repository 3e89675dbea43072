#!/bin/bash
# GPU call ad: QuantLinear ReLU epilogue -- quantized-model parity tests and the model-level rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_quantized_model.py tests/test_gpu_models.py tests/test_gpu_quant.py -x -q -m gpu > gpurun_out/ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ad_pytest.log | cut -c1-300
timeout 300 python tools/bench_models.py > gpurun_out/ad_models.jsonl 2> gpurun_out/ad_models.err; cut -c1-220 gpurun_out/ad_models.jsonl; tail -2 gpurun_out/ad_models.err
