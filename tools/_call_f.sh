#!/bin/bash
# GPU call f: ws2 (parity streams) parity + speed, no-MUFU timing experiments, MMA microbenchmark
mkdir -p gpurun_out
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws2.so timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_modules.py -x -q > gpurun_out/f_pytest_ws2.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest_ws2.log
tail -4 gpurun_out/f_pytest_ws2.log
timeout 900 bash tools/run_variants.sh 2048,8192 1,3 base ws2 ws2poly2 ws2poly4 nomufu ws2nomufu > gpurun_out/f_variants.txt 2>&1
cat gpurun_out/f_variants.txt
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws2trace.so timeout 120 python tools/trace_ws.py --kind 1 > gpurun_out/f_trace_ws2_sm1.txt 2>&1; cat gpurun_out/f_trace_ws2_sm1.txt
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_ws2.so timeout 300 python tools/bench_kernels.py --seqs 256,512,1024,2048,4096,8192 --kinds 0,1,3 --no-quant > gpurun_out/f_sweep_ws2.txt 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/f_sweep_ws2.txt'):
    try: r=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(f"T={r['T']:5d} kind={r['kind']} causal={int(r['causal'])} {r['ms']:.4f} ms {r['tflops']:.1f} TF/s")
PY
cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../outeffhop_b200/csrc -o mma_bench mma_bench.cu 2>/dev/null; timeout 60 ./mma_bench > ../../gpurun_out/f_mma_bench.txt 2>&1; cd ../..
cat gpurun_out/f_mma_bench.txt
