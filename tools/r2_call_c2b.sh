#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"^(k_shade|k_resolve)$" -s 4 -c 2 -o $OUT/prof_c2_r2f python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r2f_ncu_c2.log 2>&1; ls -la $OUT/prof_c2_r2f.ncu-rep; grep "==PROF== Profiling" $OUT/r2f_ncu_c2.log | cut -c1-120
