set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_d.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; cat gpurun_out/bench_d.json; tail -5 gpurun_out/bench_d.err
timeout 600 python tools/bench_kernels.py > gpurun_out/kernel_sweep_d.json 2> gpurun_out/kernel_sweep_d.err; echo "sweep rc=$?"; tail -5 gpurun_out/kernel_sweep_d.err
