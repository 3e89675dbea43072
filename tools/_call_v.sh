#!/bin/bash
# GPU call v/w: causal work order (heaviest-first over chunks of heads) -- causal sweep old vs new order, non-persistent vs
# persistent; erf GELU kernel tests + timing; bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gelu.py -x -q > gpurun_out/w_pytest_gelu.log 2>&1; echo "gelu rc=$?"; tail -15 gpurun_out/w_pytest_gelu.log | cut -c1-300
timeout 200 python tools/bench_gelu.py > gpurun_out/w_gelu.jsonl 2> gpurun_out/w_gelu.err; cat gpurun_out/w_gelu.jsonl; tail -3 gpurun_out/w_gelu.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-sweep --no-cpu > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; tail -c 1500 gpurun_out/w_bench.json | cut -c1-1500
for cfg in "0 0 1" "0 0 2" "x 0 1" "x 0 2" "x 1 2" "x 1 1" "16 0 1" "16 1 2"; do set -- $cfg
  echo "== chunk=$1 persistent=$2 tiles=$3"
  if [ "$1" = x ]; then unset OEH_CAUSAL_CHUNK; else export OEH_CAUSAL_CHUNK=$1; fi
  OEH_WS_PERSISTENT=$2 OEH_WS_TILES=$3 timeout 300 python tools/bench_kernels.py --seqs 512,1024,2048,4096,8192 --kinds 0,1,3 --no-quant 2>&1 | python tools/fmt_sweep.py --causal-only
done > gpurun_out/w_causal.txt 2>&1
unset OEH_CAUSAL_CHUNK
cat gpurun_out/w_causal.txt
