#!/bin/bash
# GPU call ab: wave shaping of the two elementwise-class kernels: GELU grid = resident blocks (8 x 256 threads at 32 registers vs
# 6 at <= 42), add+LayerNorm with 8 / 4 / 2 rows (warps) per block; kernel timings and the model step
mkdir -p gpurun_out
for v in base gb6 lnw4 lnw2; do
  if [ "$v" = base ]; then unset OEH_LIB_PATH; else export OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so; fi
  echo "== $v"
  case $v in base|gb6) timeout 200 python tools/bench_gelu.py 2>&1 | grep float16 | cut -c1-200;; esac
  case $v in base|lnw4|lnw2) timeout 200 python tools/bench_kernels.py --seqs 128 --kinds 1 --no-causal 2>&1 | grep -i "layer_norm" | cut -c1-300;; esac
  for i in 1 2; do
  timeout 300 python bench.py --steps 100 --warmup 10 --no-sweep --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  ms_per_step', round(d['ms_per_step'], 4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
"
  done
done > gpurun_out/ab_waves.txt 2>&1
cat gpurun_out/ab_waves.txt
