#!/bin/bash
# Round-2 one-GPU evidence pass (gpurun --timeout 1200 -- 'bash tools/r2_gpu1.sh [tag]'): GPU tests, the bench line and its
# reference arm, the other modes / workloads, the ncu launch list and one full capture of the frame's kernels, then
# compute-sanitizer on the small parity scenes. Every step under its own timeout; logs under gpurun_out/<tag>_*.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
step() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > $OUT/${TAG}_$name.log 2>&1
  echo "[$name] exit $? in $(( $(date +%s) - t0 )) s (t+$(( $(date +%s) - T0 ))): $(tail -1 $OUT/${TAG}_$name.log | cut -c1-160)"
}
bench() {  # name, timeout, env..., -- args
  local name=$1 t=$2; shift 2
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local t0=$(date +%s)
  env "${envs[@]}" timeout "$t" python bench.py "$@" > $OUT/bench_${name}_${TAG}.json 2> $OUT/bench_${name}_${TAG}.err
  echo "[bench $name] exit $? in $(( $(date +%s) - t0 )) s (t+$(( $(date +%s) - T0 )))"
  python tools/bench_brief.py "$name" < $OUT/bench_${name}_${TAG}.json 2>/dev/null || tail -3 $OUT/bench_${name}_${TAG}.err
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
nproc
step pytest_gpu 700 python -m pytest tests -m gpu -q -x --durations=8
grep -E "passed|failed|skipped" $OUT/${TAG}_pytest_gpu.log | tail -3
step smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
bench c3 300 A=1 -- --steps 20 --warmup 3
bench reference 300 A=1 -- --impl reference --steps 2 --warmup 1
bench exact 200 PRC_FMA=exact -- --steps 20 --warmup 3 --no-cpu-baseline
bench close 200 A=1 -- --steps 20 --warmup 3 --no-cpu-baseline --workload C3-close
bench c1 120 A=1 -- --steps 30 --warmup 5 --no-cpu-baseline --workload C1
bench c2 120 A=1 -- --steps 30 --warmup 5 --no-cpu-baseline --workload C2
# launch list of the bench command (cold-cache, serialised: compare SHARES)
step ncu_launches 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
# full capture of the frame's kernels (geometry/raster sweeps, shading)
step ncu_full 400 ncu --set full --clock-control none --import-source on -k regex:"k_geom_raster|k_resolve|k_shade|k_medium|k_tile" -s 8 -c 8 -o $OUT/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline
ls -la $OUT/prof_${TAG}.ncu-rep
[ -n "$SKIP_SANITIZE" ] || step sanitize 700 bash tools/sanitize.sh
cat $OUT/${TAG}_sanitize.log 2>/dev/null | tail -5
echo "total $(( $(date +%s) - T0 )) s"
