#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 4 --steps 40 --warmup 5 --no-cpu-baseline > $OUT/n4_bench.json 2> $OUT/n4_bench.err
python tools/bench_brief.py n4 < $OUT/n4_bench.json || tail -5 $OUT/n4_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/n4_bench.json"))
print("matches_1gpu", d.get("matches_1gpu"), "strips", d["run"].get("strip_rows"), "waits", d.get("peer_wait_ms_per_step_rank0"))
PY
