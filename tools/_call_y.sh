#!/bin/bash
# GPU call y: one-tile kernel (C-BERT shape) -- S row in registers (3 CTAs / SM) vs chunked TMEM passes (4 CTAs / SM),
# polynomial exp2 share 0 / 3 / 6 of 16
mkdir -p gpurun_out
OEH_ONE_REGS=1 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_modules.py tests/test_gpu_models.py -x -q > gpurun_out/y_pytest_regs.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/y_pytest_regs.log | cut -c1-300
for v in base poly0 poly6; do
  if [ "$v" = base ]; then unset OEH_LIB_PATH; else export OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so; fi
  for r in 0 1; do
    echo "== $v regs=$r"; OEH_ONE_REGS=$r timeout 200 python tools/bench_bert_attn.py --kernels default --kinds 1,3 2>&1 | cut -c1-260
  done
done > gpurun_out/y_one_variants.txt 2>&1
unset OEH_LIB_PATH
cat gpurun_out/y_one_variants.txt
