#!/bin/bash
# GPU-box script: one ncu --set full capture of the K1T kernel (4th launch) with source correlation.
OUT=gpurun_out/${1:-r2k}; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject_tc_kernel --launch-skip 3 --launch-count 1 -o $OUT/k1t -f python tools/k1t_debug.py timing > $OUT/ncu.log 2>&1
tail -5 $OUT/ncu.log
