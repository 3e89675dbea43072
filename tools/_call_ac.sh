#!/bin/bash
# GPU call ac: verification at HEAD on a fresh box: GPU tests, smoke, bench (own arm + reference arm), kernel sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ac_pytest.log; tail -4 gpurun_out/ac_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ac_smoke.log 2>&1; tail -1 gpurun_out/ac_smoke.log
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/ac_bench_reference.json 2> gpurun_out/ac_bench_reference.err; tail -c 700 gpurun_out/ac_bench_reference.json
timeout 600 python bench.py > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err; tail -c 3200 gpurun_out/ac_bench.json
timeout 400 python tools/bench_kernels.py --out gpurun_out/ac_kernel_sweep.json > gpurun_out/ac_kernel_sweep.log 2>&1; python tools/fmt_sweep.py < gpurun_out/ac_kernel_sweep.log | grep -E "T= *(2048|4096|8192)"
