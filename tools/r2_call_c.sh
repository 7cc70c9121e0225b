#!/bin/bash
# group/batch tests + regression + A/B of the compact shading build against the previous one (variants/lib_base.so)
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_group.py -m gpu -q -x 2>&1 | tail -15 > $OUT/c_group.log; tail -4 $OUT/c_group.log
timeout 400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_group.py 2>&1 | tail -8 > $OUT/c_pytest.log; tail -3 $OUT/c_pytest.log
ab() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/c_$tag.json 2> $OUT/c_$tag.err
  python tools/bench_brief.py $tag < $OUT/c_$tag.json || tail -3 $OUT/c_$tag.err
}
for rep in 1 2; do
ab base$rep PRC_LIB=$PWD/polyred_b200/csrc/variants/lib_base.so
ab new$rep A=1
done
ab base_exact PRC_LIB=$PWD/polyred_b200/csrc/variants/lib_base.so PRC_FMA=exact
ab new_exact PRC_FMA=exact
env PRC_LIB=$PWD/polyred_b200/csrc/variants/lib_base.so timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload C3-close > $OUT/c_base_close.json 2>/dev/null; python tools/bench_brief.py base_close < $OUT/c_base_close.json
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload C3-close > $OUT/c_new_close.json 2>/dev/null; python tools/bench_brief.py new_close < $OUT/c_new_close.json
# C5 on one GPU: per-view calls (round 1) vs prc_render_batch
timeout 200 python tools/multiview_bench.py --views 64 --repeat 2 --per-view-calls > $OUT/c_mv_perview.json 2> $OUT/c_mv_perview.err; cat $OUT/c_mv_perview.json | cut -c1-400
timeout 200 python tools/multiview_bench.py --views 64 --repeat 2 > $OUT/c_mv_batch.json 2> $OUT/c_mv_batch.err; cat $OUT/c_mv_batch.json | cut -c1-400
tail -3 $OUT/c_mv_batch.err
