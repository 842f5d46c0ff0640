#!/bin/bash
# Build libgsn with extra -D flags into variants/<name>.so (tuning experiments; not part of the product build).
#   tools/build_variant.sh name -DGSN_WIDE_WARPS=4 -DGSN_WIDE_CTAS=2 -DGSN_WIDE_STAGES=4
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d); mkdir -p "$root/variants"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC $*"
cd "$root/gaussn_b200/csrc"
nvcc $F -c -o $T/api.o gsn_api.cu &
nvcc $F -DGSN_SLICE_CFG=CfgWide -DGSN_SLICE_NAME=launch_logl_wide_a -DGSN_SLICE_LO=0 -DGSN_SLICE_HI=4 -c -o $T/wa.o gsn_kernels.cu &
nvcc $F -DGSN_SLICE_CFG=CfgWide -DGSN_SLICE_NAME=launch_logl_wide_b -DGSN_SLICE_LO=5 -DGSN_SLICE_HI=10 -c -o $T/wb.o gsn_kernels.cu &
nvcc $F -DGSN_SLICE_CFG=CfgNarrow -DGSN_SLICE_NAME=launch_logl_narrow_a -DGSN_SLICE_LO=0 -DGSN_SLICE_HI=4 -c -o $T/na.o gsn_kernels.cu &
nvcc $F -DGSN_SLICE_CFG=CfgNarrow -DGSN_SLICE_NAME=launch_logl_narrow_b -DGSN_SLICE_LO=5 -DGSN_SLICE_HI=10 -c -o $T/nb.o gsn_kernels.cu &
nvcc $F -DGSN_SLICE_CFG=CfgTiny -DGSN_SLICE_NAME=launch_logl_tiny_a -DGSN_SLICE_LO=0 -DGSN_SLICE_HI=4 -c -o $T/ta.o gsn_kernels.cu &
nvcc $F -DGSN_SLICE_CFG=CfgTiny -DGSN_SLICE_NAME=launch_logl_tiny_b -DGSN_SLICE_LO=5 -DGSN_SLICE_HI=10 -c -o $T/tb.o gsn_kernels.cu &
nvcc $F -DGSN_SLICE_NAME=launch_logl_cluster_a -DGSN_SLICE_LO=0 -DGSN_SLICE_HI=4 -c -o $T/ca.o gsn_cluster.cu &
nvcc $F -DGSN_SLICE_NAME=launch_logl_cluster_b -DGSN_SLICE_LO=5 -DGSN_SLICE_HI=10 -c -o $T/cb.o gsn_cluster.cu &
nvcc $F -DGSN_SLICE_NAME=launch_logl_onchip_a -DGSN_SLICE_LO=0 -DGSN_SLICE_HI=4 -c -o $T/oa.o gsn_onchip.cu &
nvcc $F -DGSN_SLICE_NAME=launch_logl_onchip_b -DGSN_SLICE_LO=5 -DGSN_SLICE_HI=10 -c -o $T/ob.o gsn_onchip.cu &
nvcc $F -DGSN_SLICE_NAME=launch_predict_a -DGSN_SLICE_LO=0 -DGSN_SLICE_HI=4 -c -o $T/pa.o gsn_predict.cu &
nvcc $F -DGSN_SLICE_NAME=launch_predict_b -DGSN_SLICE_LO=5 -DGSN_SLICE_HI=10 -c -o $T/pb.o gsn_predict.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/variants/$name.so" $T/*.o
rm -rf $T; echo "$root/variants/$name.so"
