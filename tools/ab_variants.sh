#!/bin/bash
# A/B of the tuning builds against the default library, on the GPU box (build them HERE first: make -C polyred_b200/csrc variants).
# Each variant must first pass the bit-exact parity tests, then is timed with the bench's device leg.
#   gpurun --timeout 600 -- 'bash tools/ab_variants.sh'
mkdir -p gpurun_out
for lib in default polyred_b200/csrc/variants/*.so; do
  name=$(basename $lib .so)
  [ "$lib" = default ] && unset PRC_LIB || export PRC_LIB=$PWD/$lib
  if [ "$lib" != default ]; then
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/ab_${name}_tests.log 2>&1 || { echo "$name: PARITY TESTS FAILED"; tail -5 gpurun_out/ab_${name}_tests.log; continue; }
  fi
  for rep in 1 2; do
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${name}_$rep.json 2> gpurun_out/ab_${name}_$rep.err
    python - "$name" gpurun_out/ab_${name}_$rep.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
k = d["kernel_ms_per_step"]
print(f"{sys.argv[1]:>12}: {d['ms_per_step']:.4f} ms/frame  e2e {d['e2e']['ms_per_step']:.3f}  shadow sweep {k['geom_raster_shadow']:.4f}  camera {k['geom_raster_camera']:.4f}  shade {k['shade']:.4f}")
PY
  done
done
# e2e only: row bands of the shading pass under the device->host copy (default 8)
unset PRC_LIB
for nb in 8 12 16 24; do
  PRC_SHADE_BANDS=$nb python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_bands_$nb.json 2> gpurun_out/ab_bands_$nb.err
  python - "$nb" gpurun_out/ab_bands_$nb.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print(f"shade bands {sys.argv[1]:>3}: e2e {d['e2e']['ms_per_step']:.4f} ms/frame  (device {d['ms_per_step']:.4f})")
PY
done
