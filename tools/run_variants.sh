#!/bin/bash
# Run the attention kernel sweep for several liboeh_<NAME>.so variants (built by tools/build_variant.sh):
#   tools/run_variants.sh "<seqs>" "<kinds>" NAME...     ("base" = the product library)
SEQS=$1; KINDS=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset OEH_LIB_PATH; else export OEH_LIB_PATH=$PWD/outeffhop_b200/lib/liboeh_$v.so; fi
  echo "== $v"
  python tools/bench_kernels.py --seqs $SEQS --kinds $KINDS --no-quant --no-causal 2>&1 | \
    python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(f\"  T={r['T']:5d} kind={r['kind']} causal={int(r['causal'])}  {r['ms']:.4f} ms  {r['tflops']:.1f} TF/s (best {r['tflops_best']:.1f})\")
"
done
