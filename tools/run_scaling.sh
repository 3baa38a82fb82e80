#!/bin/bash
# One 8-GPU box: host-link probe at 2/4/8 ranks (with and without NUMA binding) and bench.py at 4 and 8 ranks.
# usage (GPU box): bash tools/run_scaling.sh <outdir>
OUT=${1:-gpurun_out/r2d}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)" >> $OUT/topo.txt
python tools/pcie_probe.py > $OUT/probe_n1.json 2>/dev/null
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port 2951$N tools/pcie_probe.py > $OUT/probe_n$N.json 2>/dev/null
  $TR --nproc-per-node $N --master-port 2952$N tools/pcie_probe.py --numa > $OUT/probe_n${N}_numa.json 2>/dev/null
done
for N in 8 4; do
  timeout 400 $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29549 bench.py --gpus 8 --steps 5 --warmup 3 --no-numa-bind --no-cooperative > $OUT/bench_n8_nonuma.json 2> $OUT/bench_n8_nonuma.err
cat $OUT/probe_n*.json
