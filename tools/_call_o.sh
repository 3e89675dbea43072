#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_quantized_model.py -x -q > gpurun_out/o_pytest_qmodel.log 2>&1; echo "rc=$?" >> gpurun_out/o_pytest_qmodel.log
tail -40 gpurun_out/o_pytest_qmodel.log
