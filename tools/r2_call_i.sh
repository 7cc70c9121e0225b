#!/bin/bash
# the lazy-loading stall of same-device ranks: reproduce it without the preload, show it gone with it, then the whole -m gpu suite in ONE
# process exactly as the driver runs it, then the configuration that failed in call H three times over
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
PRC_NO_PRELOAD=1 timeout 120 python -m pytest tests/test_gpu_group.py -m gpu -q -x -k "submits_late" 2>&1 | tail -6 > $OUT/i_late_nopreload.log; tail -3 $OUT/i_late_nopreload.log; t
PRC_DEBUG_PRELOAD=1 timeout 120 python -m pytest tests/test_gpu_group.py -m gpu -q -x -s -k "submits_late" 2>&1 | tail -6 > $OUT/i_late_preload.log; tail -4 $OUT/i_late_preload.log; t
timeout 500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12 > $OUT/i_pytest_all.log; tail -3 $OUT/i_pytest_all.log; t
for k in 1 2 3; do
  PRC_ZERO_COPY_OUT=1 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -1
done; t
