#!/bin/bash
# ncu --set full of one launch (over a read chunk of configs[2] at quarter size) of the kernels named on the command line;
# summaries land in gpurun_out/r2/prof/<kernel>.txt   usage: tools/prof_stage.sh k_select k_gap k_emit
O=gpurun_out/r2/prof; mkdir -p $O
for K in "$@"; do
  timeout 600 ncu --clock-control none --set full --import-source on -k $K -s 7 -c 1 -f -o $O/${K}_c2 python bench.py --scale 0.25 --steps 1 --warmup 3 --no-cpu --no-parity > /dev/null 2>&1
  python tools/ncu_summary.py $O/${K}_c2.ncu-rep > $O/${K}.txt 2>&1
  tail -n +1 $O/${K}.txt | head -40
done
