#!/bin/bash
# GPU call i: ws4 (early S release, single S buffer per tile) parity + speed
mkdir -p gpurun_out
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws4.so timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_modules.py -x -q > gpurun_out/i_pytest_ws4.log 2>&1; echo "rc=$?" >> gpurun_out/i_pytest_ws4.log
tail -4 gpurun_out/i_pytest_ws4.log
timeout 900 bash tools/run_variants.sh 2048,8192 1,3 base ws4 ws4poly2 ws4poly4 ws4poly5 ws4nomufu > gpurun_out/i_variants.txt 2>&1
cat gpurun_out/i_variants.txt
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws4trace.so timeout 120 python tools/trace_ws.py --kind 1 > gpurun_out/i_trace_ws4_sm1.txt 2>&1; cat gpurun_out/i_trace_ws4_sm1.txt
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws4.so timeout 300 python tools/bench_kernels.py --seqs 256,512,1024,2048,4096,8192 --kinds 1,3 --no-quant > gpurun_out/i_sweep_ws4.txt 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/i_sweep_ws4.txt'):
    try: r=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(f"T={r['T']:5d} kind={r['kind']} causal={int(r['causal'])} {r['ms']:.4f} ms {r['tflops']:.1f} TF/s")
PY
