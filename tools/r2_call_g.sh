#!/bin/bash
# One 1-GPU call after the last peer-path commit: (1) the group / peer tests with full output (the 8-GPU call saw one group test fail
# without keeping its message), with the peer switches one at a time if they fail; (2) the rest of -m gpu; (3) e2e / two-stream /
# shade-occupancy tuning variants, each with the frame CRC or the parity tests in front of the timing.
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
timeout 300 python -m pytest tests/test_gpu_group.py tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -80 > $OUT/g_group_peer.log; tail -4 $OUT/g_group_peer.log; t
if grep -q "failed\|error" $OUT/g_group_peer.log; then
  for v in PRC_PEER_NO_EARLY_PUSH=1 PRC_PEER_NO_FUSED_SIGNALS=1 PRC_PEER_ONE_STREAM=1 PRC_GROUP_BALANCE=0; do
    env $v timeout 200 python -m pytest tests/test_gpu_group.py tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -40 > $OUT/g_group_peer_$v.log; echo "[$v] $(tail -1 $OUT/g_group_peer_$v.log)"
  done; t
fi
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_group.py --deselect tests/test_gpu_peer.py 2>&1 | tail -25 > $OUT/g_pytest_rest.log; tail -3 $OUT/g_pytest_rest.log; t
timeout 240 python tools/e2e_variants.py "PRC_SHADE_BANDS=16" "PRC_SHADE_BANDS=32" "PRC_STAGE_UNIFORMS=1" "PRC_STAGE_UNIFORMS=1 PRC_SHADE_BANDS=16" "PRC_TWO_STREAMS=1" \
    "PRC_TWO_STREAMS=1 PRC_STAGE_UNIFORMS=1 PRC_SHADE_BANDS=16" > $OUT/g_e2e_variants.jsonl 2> $OUT/g_e2e_variants.err; cat $OUT/g_e2e_variants.jsonl | cut -c1-330; tail -2 $OUT/g_e2e_variants.err; t
PRC_TWO_STREAMS=1 PRC_STAGE_UNIFORMS=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -8 > $OUT/g_two_streams_tests.log; tail -2 $OUT/g_two_streams_tests.log; t
ab() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/g_$tag.json 2> $OUT/g_$tag.err
  python tools/bench_brief.py $tag < $OUT/g_$tag.json || tail -3 $OUT/g_$tag.err
}
ab default A=1
for v in fused10 fused12; do
  lib=$PWD/polyred_b200/csrc/variants/lib_$v.so
  [ -f $lib ] || continue
  PRC_LIB=$lib timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
  ab $v PRC_LIB=$lib
done
ab default_exact PRC_FMA=exact
ab two_streams PRC_TWO_STREAMS=1; t
