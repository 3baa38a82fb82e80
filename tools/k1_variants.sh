#!/bin/bash
# GPU-box script: parity (K1T pytest) and timing of every mulit_view_object_detection_b200/libmvfusion_*.so variant
# (tools/build_variant.sh) against the production build.
P=mulit_view_object_detection_b200
cp $P/libmvfusion.so /tmp/lib_prod.so
echo "== production"; timeout 100 python tools/k1t_debug.py timing 2>&1 | grep "tensor_cores=True"
for f in $P/libmvfusion_*.so; do
  case $f in *_prof.so|*_dbg.so) continue;; esac
  echo "== $f"; cp $f $P/libmvfusion.so
  timeout 300 python -m pytest tests/test_gpu_unproject_tc.py -x -q -m gpu --timeout 60 --timeout-method thread 2>&1 | tail -1
  timeout 100 python tools/k1t_debug.py timing 2>&1 | grep "tensor_cores=True"
done
cp /tmp/lib_prod.so $P/libmvfusion.so
