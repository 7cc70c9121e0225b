#!/bin/bash
# key clear under the shadow sweep (second stream): the whole -m gpu suite with it, then A/B against PRC_NO_EARLY_CLEAR=1
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 > $OUT/l_pytest_all.log; tail -2 $OUT/l_pytest_all.log
ab() { local tag=$1; shift; env "$@" timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $OUT/l_$tag.json 2> $OUT/l_$tag.err; python tools/bench_brief.py $tag < $OUT/l_$tag.json || tail -3 $OUT/l_$tag.err; }
ab early_clear A=1
ab no_early_clear PRC_NO_EARLY_CLEAR=1
