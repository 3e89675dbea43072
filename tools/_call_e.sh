#!/bin/bash
# GPU call e: tests, bench, trace, poly variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/e_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -3 gpurun_out/e_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; tail -c 600 gpurun_out/e_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-sweep --no-cpu --profile-one-step > gpurun_out/e_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/e_launches.csv > gpurun_out/e_launches.summary.txt 2>&1; head -25 gpurun_out/e_launches.summary.txt
for v in trace; do OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so timeout 120 python tools/trace_ws.py --kind 1 > gpurun_out/e_trace_sm1.txt 2>&1; OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so timeout 120 python tools/trace_ws.py --kind 3 > gpurun_out/e_trace_clip.txt 2>&1; done
cat gpurun_out/e_trace_sm1.txt
timeout 900 bash tools/run_variants.sh 2048,8192 1,3 base poly2 poly4 poly6 poly8 noskip stag400 stag800 > gpurun_out/e_variants.txt 2>&1
cat gpurun_out/e_variants.txt
timeout 200 python tools/bench_kernels.py --seqs 2048,8192 --kinds 3 --no-quant --no-causal --qscale 4 > gpurun_out/e_peaky.txt 2>&1; cat gpurun_out/e_peaky.txt
cd tools/micro && for b in mufu_bench mufu16_bench; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $b $b.cu 2>/dev/null; timeout 60 ./$b > ../../gpurun_out/e_$b.txt 2>&1; done; cd ../..
cat gpurun_out/e_mufu16_bench.txt | tail -20
