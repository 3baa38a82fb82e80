#!/bin/bash
# GPU-box script: K1T parity (pytest), production timing, per-role cycle counters of the profiling build.
OUT=gpurun_out/${1:-r2g}; mkdir -p $OUT
P=mulit_view_object_detection_b200
timeout 300 python -m pytest tests/test_gpu_unproject_tc.py -x -q -m gpu --timeout 60 --timeout-method thread > $OUT/pytest_k1t.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_k1t.log
timeout 200 python tools/k1t_debug.py timing 2>&1 | tee $OUT/timing_prod.log
if [ -f $P/libmvfusion_prof.so ]; then
cp $P/libmvfusion.so /tmp/lib_prod.so
cp $P/libmvfusion_prof.so $P/libmvfusion.so
timeout 100 python tools/k1t_debug.py prof 2>&1 | tee $OUT/prof.log
cp /tmp/lib_prod.so $P/libmvfusion.so
fi
