#!/bin/bash
# tools/build_variant.sh NAME "EXTRA NVCC FLAGS"  -> mulit_view_object_detection_b200/libmvfusion_NAME.so (A/B measurement builds;
# tools/k1_variants.sh swaps them in on the GPU box).  Only unproject_tc.cu is recompiled, the other objects come from `make`.
set -e
cd "$(dirname "$0")/../mulit_view_object_detection_b200/csrc"
NAME=$1; shift
mkdir -p /tmp/mvf_var_$NAME
for f in unproject_tc api; do
SRC=$f.cu; if [ "$f" = unproject_tc ] && [ -n "$K1T_SRC" ]; then SRC=$K1T_SRC; fi
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off -I../../include -I. "$@" -c $SRC -o /tmp/mvf_var_$NAME/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libmvfusion_$NAME.so unproject.o /tmp/mvf_var_$NAME/unproject_tc.o project.o fuse.o convlstm_tc.o roi_align.o detection.o /tmp/mvf_var_$NAME/api.o
echo built libmvfusion_$NAME.so
