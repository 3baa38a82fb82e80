#!/bin/bash
# GPU-box script: timing of the DEBUG_ENV build under the K1T ablation switches (see K1T_DBG in unproject_tc.cu)
P=mulit_view_object_detection_b200
cp $P/libmvfusion.so /tmp/lib_prod.so; cp $P/libmvfusion_dbg.so $P/libmvfusion.so
for d in ${@:-0 1 2 4 6 64 128 24 56}; do echo -n "MVF_K1T_DBG=$d  "; MVF_K1T_DBG=$d timeout 100 python tools/k1t_debug.py timing 2>&1 | grep "tensor_cores=True"; done
cp /tmp/lib_prod.so $P/libmvfusion.so
