#!/bin/bash
# GPU call g: ws3 (P through smem) parity + speed; full test suite and bench on the default library
mkdir -p gpurun_out
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws3.so timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_modules.py -x -q > gpurun_out/g_pytest_ws3.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest_ws3.log
tail -4 gpurun_out/g_pytest_ws3.log
timeout 900 bash tools/run_variants.sh 2048,8192 1,3 base ws3 ws3poly0 ws3poly4 ws3nomufu ws2two > gpurun_out/g_variants.txt 2>&1
cat gpurun_out/g_variants.txt
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws3trace.so timeout 120 python tools/trace_ws.py --kind 1 > gpurun_out/g_trace_ws3_sm1.txt 2>&1; cat gpurun_out/g_trace_ws3_sm1.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -5 gpurun_out/g_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -c 1500 gpurun_out/g_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-sweep --no-cpu --profile-one-step > gpurun_out/g_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/g_launches.csv > gpurun_out/g_launches.summary.txt 2>&1; head -16 gpurun_out/g_launches.summary.txt
