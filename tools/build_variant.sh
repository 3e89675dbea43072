#!/bin/bash
# Build an experimental variant of liboeh.so with extra -D flags for the attention kernels:
#   tools/build_variant.sh NAME -DOEH_EXP=1      -> outeffhop_b200/lib/liboeh_NAME.so   (use with OEH_LIB_PATH)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
C=outeffhop_b200/csrc
nvcc -c $C/attn_fwd_sm100.cu -o $C/_obj/attn_fwd_sm100_$NAME.o -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  --expt-relaxed-constexpr -Xcompiler -fPIC,-O2 -I include "$@"
nvcc -shared -o outeffhop_b200/lib/liboeh_$NAME.so $C/_obj/api.o $C/_obj/quant.o $C/_obj/softmax_rows.o $C/_obj/gate.o \
  $C/_obj/attn_fwd_sm100_$NAME.o -gencode arch=compute_100a,code=sm_100a
echo outeffhop_b200/lib/liboeh_$NAME.so
