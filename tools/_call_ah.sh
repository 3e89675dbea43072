#!/bin/bash
# GPU call ah (8 GPUs): N=1 / 2 / 4 / 8 bench lines back to back on one box (weak scaling, no data-path collective)
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-sweep --no-cpu > gpurun_out/ah_bench_n1.json 2> gpurun_out/ah_bench_n1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 50 --warmup 5 --no-sweep --no-cpu > gpurun_out/ah_bench_n$n.json 2> gpurun_out/ah_bench_n$n.err
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        for l in open(f"gpurun_out/ah_bench_n{n}.json"):
            if l.startswith("{"):
                d = json.loads(l); print(n, round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/ah_bench_n8.err
