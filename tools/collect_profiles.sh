#!/bin/bash
# Run on the GPU box (under gpurun): bench line + ncu launch list + one full capture of the dominant kernel.
# usage: tools/collect_profiles.sh <tag>
set -x
TAG=${1:-r1}
OUT=gpurun_out
python bench.py --steps 20 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_${TAG}.json 2> $OUT/bench_reference_${TAG}.err
PRC_FMA=exact python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_exact_${TAG}.json 2>/dev/null
PRC_FMA=fast python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_fast_${TAG}.json 2>/dev/null
# launch list of the same command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 30 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
# full capture of the geometry/raster kernels (shadow sweep + camera) and the shading kernels of one frame
ncu --set full --clock-control none --import-source on -k regex:"k_geom_raster|k_resolve|k_shade" -s 5 -c 5 -o $OUT/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
# C5 multi-view throughput on one GPU (the N>1 points come from a --gpus N box, tools/multiview_bench.py under torchrun)
python tools/multiview_bench.py --views 256 --repeat 2 > $OUT/multiview_1.json 2> $OUT/multiview_1.err
python tools/bench_brief.py final < $OUT/bench_${TAG}.json
