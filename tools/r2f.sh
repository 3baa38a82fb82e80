#!/bin/bash
mkdir -p gpurun_out/r2f
P=mulit_view_object_detection_b200
timeout 200 python tools/k1t_debug.py timing 2>&1 | tee gpurun_out/r2f/timing_prod.log
cp $P/libmvfusion.so /tmp/lib_prod.so
cp $P/libmvfusion_prof.so $P/libmvfusion.so
for d in 0 24; do echo "== prof MVF_K1T_DBG=$d"; MVF_K1T_DBG=$d timeout 100 python tools/k1t_debug.py prof 2>&1; done | tee gpurun_out/r2f/prof.log
cp /tmp/lib_prod.so $P/libmvfusion.so
