#!/bin/bash
# GPU call x: full GPU suite at the new dispatch rule + GELU kernel; GELU launch-shape variants; default-dispatch causal sweep;
# model-level rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log; tail -4 gpurun_out/x_pytest.log | cut -c1-300
for v in base gu1b8 gu4 gu4b4 gu8b2; do
  if [ "$v" = base ]; then unset OEH_LIB_PATH; else export OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so; fi
  echo "== $v"; timeout 200 python tools/bench_gelu.py 2>&1 | grep float16 | cut -c1-260
done > gpurun_out/x_gelu_variants.txt 2>&1
unset OEH_LIB_PATH
cat gpurun_out/x_gelu_variants.txt
echo "== default dispatch" > gpurun_out/x_causal_default.txt
timeout 300 python tools/bench_kernels.py --seqs 512,1024,2048,4096,8192 --kinds 0,1,3 --no-quant 2>&1 | python tools/fmt_sweep.py >> gpurun_out/x_causal_default.txt 2>&1
cat gpurun_out/x_causal_default.txt
timeout 300 python tools/bench_models.py > gpurun_out/x_models.jsonl 2> gpurun_out/x_models.err; cut -c1-200 gpurun_out/x_models.jsonl
