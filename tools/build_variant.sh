#!/bin/bash
# Build an experimental variant of the library: tools/build_variant.sh NAME -DRC_PROD_WARPS=8 ...
# -> tools/librc_NAME.so (use with RC_B200_LIB=tools/librc_NAME.so).  Only rc_gine_tiled.cu is recompiled.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=raincast_gnn_b200/csrc
python -m raincast_gnn_b200.csrc.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-O2,-ffp-contract=off \
  -I include -I $C "$@" -c $C/rc_gine_tiled.cu -o /tmp/rc_gine_tiled_$name.o
objs=$(ls $C/*.o | grep -v rc_gine_tiled.o)
nvcc -shared -o tools/librc_$name.so $objs /tmp/rc_gine_tiled_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo tools/librc_$name.so
