#!/bin/bash
# GPU-box script (round 2 evidence): GPU test suite, bench lines of both arms, ncu launch list of the bench command, one ncu --set full
# capture of the step's kernels (4 scenes), head kernels, K1T role counters.  Each ncu pass runs only after its command exited 0 without ncu.
OUT=gpurun_out/r2ev2; mkdir -p $OUT
P=mulit_view_object_detection_b200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 180 --timeout-method thread > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_line.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > $OUT/bench_reference_line.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --steps 2 --warmup 1 > $OUT/bench_small.json 2>/dev/null && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 > $OUT/ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 300 python bench.py --scenes 4 --steps 2 --warmup 1 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1t_presplit|unproject_tc_kernel|project_rays' --launch-skip 6 --launch-count 3 -o $OUT/k1t_step -f python bench.py --scenes 4 --steps 2 --warmup 1 > $OUT/ncu_step.log 2>&1
echo "ncu step rc=$?"; tail -2 $OUT/ncu_step.log
timeout 300 python tools/bench_heads.py > $OUT/heads_c2.json 2> $OUT/heads.err; echo "heads rc=$?"
timeout 120 python tools/bench_nms.py > $OUT/nms.json 2>&1
if [ -f $P/libmvfusion_prof.so ]; then cp $P/libmvfusion.so /tmp/lib_prod.so; cp $P/libmvfusion_prof.so $P/libmvfusion.so; timeout 100 python tools/k1t_debug.py prof > $OUT/k1t_roles.txt 2>&1; cp /tmp/lib_prod.so $P/libmvfusion.so; fi
ls -la $OUT
