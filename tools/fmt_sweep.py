#!/usr/bin/env python
"""stdin: JSON lines of tools/bench_kernels.py -> one short line per attention row (--causal-only filters)."""
import json
import sys

causal_only = "--causal-only" in sys.argv
for l in sys.stdin:
    try:
        r = json.loads(l)
    except Exception:
        print(l.rstrip()[:200])
        continue
    if "T" not in r or (causal_only and not r.get("causal")):
        continue
    print(f"  T={r['T']:5d} kind={r['kind']} causal={int(r['causal'])}  {r['ms']:.4f} ms  {r['tflops']:.1f} TF/s (best {r['tflops_best']:.1f})")
