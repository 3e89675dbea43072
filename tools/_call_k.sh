#!/bin/bash
# GPU call k: one Q tile per CTA (two CTAs per SM) vs two tiles per CTA
mkdir -p gpurun_out
OEH_WS_TILES=1 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_modules.py -x -q > gpurun_out/k_pytest_tiles1.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest_tiles1.log
tail -4 gpurun_out/k_pytest_tiles1.log
for t in 2 1; do echo "== OEH_WS_TILES=$t"; OEH_WS_TILES=$t timeout 300 python tools/bench_kernels.py --seqs 256,512,1024,2048,4096,8192 --kinds 1,3 --no-quant 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(f\"  T={r['T']:5d} kind={r['kind']} causal={int(r['causal'])}  {r['ms']:.4f} ms  {r['tflops']:.1f} TF/s (best {r['tflops_best']:.1f})\")
"; done > gpurun_out/k_tiles.txt 2>&1
cat gpurun_out/k_tiles.txt
python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu > gpurun_out/k_bench_short.json 2>gpurun_out/k_bench_short.err; python -c "
import json; d=json.loads(open('gpurun_out/k_bench_short.json').read().strip().splitlines()[-1]); print(d['value'], d['clocks'])"
timeout 200 python tools/bench_kernels.py --seqs 128 --kinds 3 --no-causal > gpurun_out/k_kernels_small.txt 2>&1; tail -12 gpurun_out/k_kernels_small.txt
