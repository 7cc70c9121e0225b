#!/bin/bash
# zero-copy readback (the shading kernel stores the frame into the page-locked host image) and fewer row bands, one process; then the
# parity / edge / group tests with the switch on
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
timeout 240 python tools/e2e_variants.py "PRC_ZERO_COPY_OUT=1" "PRC_SHADE_BANDS=4" "PRC_SHADE_BANDS=6" "PRC_ZERO_COPY_OUT=1 PRC_TWO_STREAMS=1" "PRC_SHADE_BANDS=1" \
    > $OUT/h_e2e_variants.jsonl 2> $OUT/h_e2e_variants.err; cut -c1-200 $OUT/h_e2e_variants.jsonl; tail -2 $OUT/h_e2e_variants.err; t
timeout 240 python tools/e2e_variants.py --workload C3-close "PRC_ZERO_COPY_OUT=1" > $OUT/h_e2e_variants_close.jsonl 2> $OUT/h_e2e_variants_close.err; cut -c1-200 $OUT/h_e2e_variants_close.jsonl; t
PRC_ZERO_COPY_OUT=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_group.py tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -8 > $OUT/h_zero_copy_tests.log; tail -2 $OUT/h_zero_copy_tests.log; t
PRC_ZERO_COPY_OUT=1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/h_zero_copy.json 2> $OUT/h_zero_copy.err; python tools/bench_brief.py zero_copy < $OUT/h_zero_copy.json || tail -3 $OUT/h_zero_copy.err; t
