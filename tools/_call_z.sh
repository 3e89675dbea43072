#!/bin/bash
# GPU call z: add+LayerNorm cache policy inside the model step (streaming vs plain accesses), A/B/A/B on the bench step
mkdir -p gpurun_out
for i in 1 2; do for v in base lnplain; do
  if [ "$v" = base ]; then unset OEH_LIB_PATH; else export OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so; fi
  echo "== $v run $i"
  timeout 300 python bench.py --steps 100 --warmup 10 --no-sweep --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  ms_per_step', round(d['ms_per_step'], 4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline_us', round(d['roofline']['launch_us'], 2))
"
  timeout 300 python tools/bench_models.py 2>/dev/null | cut -c1-60,100-200
done; done > gpurun_out/z_ln_policy.txt 2>&1
cat gpurun_out/z_ln_policy.txt
