#!/bin/bash
# Build an experimental copy of the library with extra -D switches (kernel experiments only; the product is lib/liboeh.so):
#   tools/build_variant.sh NAME -DOEH_POLY_PAIRS=4 ...   -> outeffhop_b200/lib/liboeh_NAME.so   (see tools/run_variants.sh)
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=/tmp/oeh_variant_$NAME
mkdir -p $OBJ $ROOT/outeffhop_b200/lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-O2 -I $ROOT/include"
pids=()
for s in api quant softmax_rows gate attn_fwd_sm100 percentile outlier_stats ce_loss layernorm layernorm_f32 layernorm_f16 layernorm_bf16 gelu gemm_sm100 gemm_ln_sm100; do
  [ -f $ROOT/outeffhop_b200/csrc/$s.cu ] || continue
  nvcc -c $ROOT/outeffhop_b200/csrc/$s.cu -o $OBJ/$s.o $FLAGS "$@" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $ROOT/outeffhop_b200/lib/liboeh_$NAME.so $OBJ/*.o -gencode arch=compute_100a,code=sm_100a
echo built $ROOT/outeffhop_b200/lib/liboeh_$NAME.so
