#!/bin/bash
# Round-2 profiling pass on one B200 (run under gpurun; everything lands in gpurun_out/r2/prof/):
#   launch list of the bench steps, ncu --set full of the dominant sketch kernel, of the lookup on a DRAM-resident
#   (human-size) index and of the single-pass tile kernel, T_e2e-file of the drop-in CLI, compute-sanitizer on a test subset.
O=gpurun_out/r2/prof; mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list: configs[2] at quarter size (1 Gbp of reads = two resident chunks), 2 steps
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $O/launches_c2_quarter.csv python bench.py --scale 0.25 --steps 2 --warmup 3 --no-cpu --no-parity > /dev/null 2>&1
python tools/launches.py $O/launches_c2_quarter.csv > $O/launches_c2_quarter.txt 2>&1
# 2. dominant kernel, full capture of one read-chunk launch
timeout 600 $NCU --set full --import-source on -k k_dense -s 6 -c 1 -o $O/k_dense_c2 python bench.py --scale 0.25 --steps 1 --warmup 3 --no-cpu --no-parity > /dev/null 2>&1
# 3. lookup on the DRAM-resident index of configs[3] (3.1 Gbp target, 1 GB table), 1 Gbp of reads
timeout 900 $NCU --set full -k k_lookup -s 2 -c 1 -o $O/k_lookup_c3 python bench.py --config c3 --reads-per-gpu 1e9 --steps 1 --warmup 3 --no-cpu --no-parity > $O/bench_c3_1gbp.json 2> $O/bench_c3_1gbp.err
# 4. the single-pass tile kernel on the same read chunk as 2.
NTL_TILE=1 timeout 600 $NCU --set full --import-source on -k k_tile -s 6 -c 1 -o $O/k_tile_c2 python bench.py --scale 0.25 --steps 1 --warmup 3 --no-cpu --no-parity > /dev/null 2>&1
NTL_TILE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_c2_tile.json 2> /dev/null
# 5. T_e2e-file
timeout 300 python tools/cli_e2e.py --config c1 > $O/cli_e2e_c1.json 2>&1
timeout 300 python tools/cli_e2e.py --config c1 --gz > $O/cli_e2e_c1_gz.json 2>&1
timeout 600 python tools/cli_e2e.py --config c2 --scale 0.25 > $O/cli_e2e_c2_quarter.json 2>&1
# 6. sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tile.py tests/test_gpu_synth.py -x -q -k "scaffold_gaps or low_complexity or deferred or generator or timed_entry" > $O/sanitizer_memcheck.txt 2>&1
tail -3 $O/sanitizer_memcheck.txt
ls -la $O
