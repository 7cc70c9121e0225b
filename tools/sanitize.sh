#!/bin/bash
# compute-sanitizer over the small parity scenes ("compute-sanitizer --tool racecheck/memcheck on the raster
# kernels"). Run on the GPU box:   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# Writes gpurun_out/sanitize_{memcheck,racecheck,initcheck}.log; a tool's exit status is non-zero if it reported an error.
# (The atomics that resolve depth are fire-and-forget RED.MAX on global memory: racecheck only looks at shared memory — the
# CTA candidate queue, the chunk's shared vertices, the tile raster's staged records.)
set -u
mkdir -p gpurun_out
TESTS="tests/test_gpu_parity.py::test_c3_city_eight_lights_four_casting tests/test_gpu_parity.py::test_c2_shadow_ao_gamma_clipped_ground"
rc=0
for tool in memcheck racecheck initcheck; do
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 \
    python -m pytest $TESTS -m gpu -x -q > gpurun_out/sanitize_$tool.log 2>&1
  s=$?
  echo "$tool: exit $s  $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log) summary line(s): $(grep 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  [ $s -ne 0 ] && rc=$s
done
exit $rc
