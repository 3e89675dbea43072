#!/bin/bash
mkdir -p gpurun_out
for t in 2 1; do
OEH_WS_PERSISTENT=1 OEH_WS_TILES=$t timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_modules.py -x -q > gpurun_out/u_pytest_persist_t$t.log 2>&1; echo "tiles=$t rc=$?"; tail -3 gpurun_out/u_pytest_persist_t$t.log | cut -c1-200
done
for cfg in "0 2" "1 2" "0 1" "1 1"; do set -- $cfg; echo "== persistent=$1 tiles=$2"; OEH_WS_PERSISTENT=$1 OEH_WS_TILES=$2 timeout 300 python tools/bench_kernels.py --seqs 512,1024,2048,4096,8192 --kinds 1,3 --no-quant 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print(f\"  T={r['T']:5d} kind={r['kind']} causal={int(r['causal'])}  {r['ms']:.4f} ms  {r['tflops']:.1f} TF/s (best {r['tflops_best']:.1f})\")
"; done > gpurun_out/u_persist.txt 2>&1
cat gpurun_out/u_persist.txt
