#!/bin/bash
# First GPU call of round 2 (one B200, ~6-8 min): everything that was written after round 1's GPU budget was spent and has
# not run on hardware yet, each step under its own timeout, logs under gpurun_out/r2_*.log.
#   (here)  make -C polyred_b200/csrc variants
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
step() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > gpurun_out/r2_$name.log 2>&1
  echo "[$name] exit $? in $(( $(date +%s) - t0 )) s: $(tail -1 gpurun_out/r2_$name.log | cut -c1-200)"
}
# 1. the peer-memory protocol incl. the two changes made after its hardware run (per-frame consumer set, strip readback)
step peer_selfcheck 60 python tools/peer_selfcheck.py 3
# 2. the regular GPU suite (must stay green) + the opt-in tests
step pytest_gpu 400 python -m pytest tests -m gpu -x -q
PRC_TEST_EDGE=1 PRC_TEST_PEER_READBACK=1 step pytest_optin 300 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_peer.py -m gpu -q
grep -E "passed|failed|FAILED" gpurun_out/r2_pytest_optin.log | tail -15
# 3. smoke + the bench line at HEAD (device / e2e), then the tuning builds and the band sweep
step smoke 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step ab_variants 420 bash tools/ab_variants.sh
cat gpurun_out/r2_ab_variants.log | tail -20
