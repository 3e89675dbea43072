#!/bin/bash
# GPU call h: consolidated ws kernel (uniform warp index), full tests, sweeps, ncu --set full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
timeout 600 bash tools/run_variants.sh 2048,8192 1,3 base poly0 poly3 poly4 > gpurun_out/h_variants.txt 2>&1
cat gpurun_out/h_variants.txt
OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_trace.so timeout 120 python tools/trace_ws.py --kind 1 > gpurun_out/h_trace_ws_sm1.txt 2>&1; cat gpurun_out/h_trace_ws_sm1.txt | head -20
timeout 300 python tools/bench_kernels.py --out gpurun_out/h_kernel_sweep.json > gpurun_out/h_kernel_sweep.log 2>&1; tail -14 gpurun_out/h_kernel_sweep.log
timeout 300 python tools/bench_kernels.py --seqs 2048,8192 --kinds 3 --no-quant --no-causal --gamma 0.0 > gpurun_out/h_clip_dense.txt 2>&1; cat gpurun_out/h_clip_dense.txt
# ncu --set full: one-tile kernel at the C-BERT shape, ws kernel one-pass and two-pass at T=4096
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_one -c 1 -s 2 -o gpurun_out/h_one_clip1_T128 -f python tools/prof_attn.py --T 128 --B 32 --kind 3 --bert-layout > gpurun_out/h_ncu_one.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_ws -c 1 -s 2 -o gpurun_out/h_ws_sm1_T4096 -f python tools/prof_attn.py --T 4096 --B 4 --kind 1 > gpurun_out/h_ncu_ws1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_ws -c 1 -s 2 -o gpurun_out/h_ws_clip1_T4096 -f python tools/prof_attn.py --T 4096 --B 4 --kind 3 > gpurun_out/h_ncu_ws3.log 2>&1
for r in h_one_clip1_T128 h_ws_sm1_T4096 h_ws_clip1_T4096; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.txt 2>&1
  ncu -i gpurun_out/$r.ncu-rep --page source --csv > gpurun_out/$r.source.csv 2>/dev/null
  python tools/ncu_hot.py gpurun_out/$r.source.csv 25 >> gpurun_out/$r.txt 2>&1
done
ls -la gpurun_out/*.ncu-rep
head -40 gpurun_out/h_one_clip1_T128.txt
