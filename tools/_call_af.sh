#!/bin/bash
# GPU call af: fused fake-quant taps of the attention kernels: rint by two FADDs after the clamp vs FRND (A/B), INT8 parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_quantized_model.py tests/test_gpu_modules.py tests/test_gpu_models.py -x -q -m gpu -k "int8 or quant or tap or Quant" > gpurun_out/af_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/af_pytest.log | cut -c1-300
for i in 1 2; do for v in frnd base; do
  if [ "$v" = base ]; then unset OEH_LIB_PATH; else export OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so; fi
  echo "== $v run $i"; timeout 300 python tools/bench_models.py 2>/dev/null | cut -c1-50,100-230
done; done > gpurun_out/af_qdq_rint.txt 2>&1
cat gpurun_out/af_qdq_rint.txt
