#!/bin/bash
# usage: prof.sh name "k,w"
mkdir -p gpurun_out/r2
SWEEP_KW="$2" timeout 600 ncu --set full --clock-control none --import-source on -k k_tile -s 1 -c 1 -o gpurun_out/r2/$1 python tests/scale/sketch_sweep.py --bases 100e6 > /dev/null 2>&1
