#!/bin/bash
# GPU call n: verification of the final round-1 library: tests, smoke, bench, launch list, kernel sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log; tail -4 gpurun_out/n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; tail -1 gpurun_out/n_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -c 2500 gpurun_out/n_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-sweep --no-cpu --profile-one-step > gpurun_out/n_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/n_launches.csv > gpurun_out/n_launches.summary.txt 2>&1; head -14 gpurun_out/n_launches.summary.txt
timeout 400 python tools/bench_kernels.py --out gpurun_out/n_kernel_sweep.json > gpurun_out/n_kernel_sweep.log 2>&1; grep -c . gpurun_out/n_kernel_sweep.log
