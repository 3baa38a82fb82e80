#!/bin/bash
# GPU-box script: K1T parity + timing of the production build, then the MMA-count ablation on the debug build.
mkdir -p gpurun_out/r2e
P=mulit_view_object_detection_b200
timeout 600 python -m pytest tests/test_gpu_unproject_tc.py -x -q -m gpu > gpurun_out/r2e/pytest_k1t.log 2>&1; echo "pytest rc=$?"
timeout 200 python tools/k1t_debug.py timing 2>&1 | tee gpurun_out/r2e/timing_prod.log
cp $P/libmvfusion.so /tmp/lib_prod.so
cp $P/libmvfusion_dbg.so $P/libmvfusion.so
for d in 0 8 24; do echo "== dbg MVF_K1T_DBG=$d"; MVF_K1T_DBG=$d timeout 100 python tools/k1t_debug.py timing 2>&1 | grep "tensor_cores=True"; done | tee gpurun_out/r2e/ablate.log
cp $P/libmvfusion_prof.so $P/libmvfusion.so
timeout 100 python tools/k1t_debug.py prof 2>&1 | tee gpurun_out/r2e/prof.log
cp /tmp/lib_prod.so $P/libmvfusion.so
