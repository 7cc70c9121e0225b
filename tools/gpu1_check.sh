#!/bin/bash
# One-GPU regression + timing pass (gpurun --timeout 900 -- 'bash tools/gpu1_check.sh [tag]'): GPU tests, then the device / e2e bench
# lines of C3 (default and with full-frame chunk culling forced), C1 and C2. Logs under gpurun_out/<tag>_*.
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
b() {  # name, env..., -- args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - "$name" gpurun_out/${TAG}_$name.json <<'PY' || tail -5 gpurun_out/${TAG}_$name.err
import json, sys
d = json.load(open(sys.argv[2]))
k = {a: round(b, 4) for a, b in d["kernel_ms_per_step"].items()}
print(f"[{sys.argv[1]}] dev {d['ms_per_step']:.4f} ms  e2e {d['e2e']['ms_per_step']:.3f}  launches/frame {d['gpu_launches'] / d['steps']:.0f}  frac {d['frame_roofline']['frac']:.3f}")
print("    ", k)
PY
}
b c3 A=1 --
b c3_cull PRC_FORCE_CHUNK_CULL=1 --
b c3_exact PRC_FMA=exact --
b c1 A=1 -- --workload C1
b c2 A=1 -- --workload C2
