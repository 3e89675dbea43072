#!/bin/bash
mkdir -p gpurun_out
OEH_WS_TILES=1 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_modules.py -x -q > gpurun_out/l_pytest_tiles1.log 2>&1; echo "rc=$?" >> gpurun_out/l_pytest_tiles1.log
tail -4 gpurun_out/l_pytest_tiles1.log
for t in 2 1; do echo "== OEH_WS_TILES=$t"; OEH_WS_TILES=$t timeout 300 python tools/bench_kernels.py --seqs 256,512,1024,2048,4096,8192 --kinds 1,3 --no-quant 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print(f\"  T={r['T']:5d} kind={r['kind']} causal={int(r['causal'])}  {r['ms']:.4f} ms  {r['tflops']:.1f} TF/s (best {r['tflops_best']:.1f})\")
"; done > gpurun_out/l_tiles.txt 2>&1
cat gpurun_out/l_tiles.txt
