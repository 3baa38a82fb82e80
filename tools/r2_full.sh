#!/bin/bash
# GPU-box script: whole GPU test suite (per-test timeout), bench line, reference-arm line.
OUT=gpurun_out/${1:-r2full}; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method thread > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
