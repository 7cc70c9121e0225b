#!/bin/bash
# the other workloads at HEAD and one full ncu capture of the AO shading kernel on C2 (a-9 had no capture at HEAD)
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
b() { local tag=$1; shift; env "$@" > /dev/null 2>&1; }
run() { local tag=$1; shift; timeout 120 env "$@" > $OUT/bench_${tag}_r2f.json 2> $OUT/bench_${tag}_r2f.err; python tools/bench_brief.py $tag < $OUT/bench_${tag}_r2f.json || tail -3 $OUT/bench_${tag}_r2f.err; }
run c2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload C2
run c1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload C1
run close python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload C3-close
run exact PRC_FMA=exact python bench.py --steps 20 --warmup 3 --no-cpu-baseline; t
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_shade<|k_resolve<" -s 4 -c 2 -o $OUT/prof_c2_r2f python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r2f_ncu_c2.log 2>&1; ls -la $OUT/prof_c2_r2f.ncu-rep; tail -2 $OUT/r2f_ncu_c2.log | cut -c1-200; t
