#!/bin/bash
# 1 GPU: the group / peer tests after PRC_FRAME_IMAGE_AT_SYNC (same-device ranks), the whole suite in one process, smoke
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
timeout 500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12 > $OUT/k_pytest_all.log; tail -3 $OUT/k_pytest_all.log; t
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300 > $OUT/k_smoke.log; cat $OUT/k_smoke.log; t
timeout 120 python tools/group_bench.py --devices 0,0 --steps 20 > $OUT/k_group00.json 2> $OUT/k_group00.err; python -c "
import json; d=json.load(open('$OUT/k_group00.json')); print('[group 0,0] dev', round(d['ms_per_frame_device_resident_wall'],4), 'e2e', round(d['e2e_ms_per_frame'],4), 'match', d['matches_1gpu'], d['strip_rows'])" || tail -3 $OUT/k_group00.err; t
PRC_FORCE_CHUNK_CULL=1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/k_force_cull.json 2> $OUT/k_force_cull.err; python tools/bench_brief.py force_chunk_cull < $OUT/k_force_cull.json || tail -3 $OUT/k_force_cull.err; t
