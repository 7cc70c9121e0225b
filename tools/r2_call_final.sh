#!/bin/bash
# final one-GPU evidence at HEAD: bench line (+ reference arm), ncu launch list of the same command, full capture of the frame's kernels
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader; nproc
timeout 200 python bench.py --steps 20 --warmup 3 > $OUT/bench_r2f.json 2> $OUT/bench_r2f.err; python tools/bench_brief.py c3 < $OUT/bench_r2f.json || tail -3 $OUT/bench_r2f.err; t
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_r2f.json 2> $OUT/bench_reference_r2f.err; cut -c1-200 $OUT/bench_reference_r2f.json; t
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $OUT/launches_r2f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r2f_ncu_launches.log 2>&1; tail -1 $OUT/launches_r2f.csv | cut -c1-200; t
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_geom_raster|k_resolve_shade|k_medium" -s 6 -c 6 -o $OUT/prof_r2f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r2f_ncu_full.log 2>&1; ls -la $OUT/prof_r2f.ncu-rep; t
