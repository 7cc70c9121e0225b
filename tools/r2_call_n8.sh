#!/bin/bash
# The one 8-GPU call of round 2: bit-identity at 8 ranks, the peer bench line (early / late key push), the one-process group API and
# the C5 view batch over 8 devices. Logs under gpurun_out/n8_*.
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
N=8
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 tests/multigpu_check.py > $OUT/multigpu_check_$N.log 2>&1
grep -c "= 0$" $OUT/multigpu_check_$N.log; grep "multigpu_check" $OUT/multigpu_check_$N.log | grep -v "= 0$" | head -5; t
run() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --mgpu peer > $OUT/n8_$tag.json 2> $OUT/n8_$tag.err
  python - "$tag" $OUT/n8_$tag.json <<'PY' || tail -5 $OUT/n8_$tag.err
import json, sys
d = json.load(open(sys.argv[2]))
k = {a: round(b, 4) for a, b in d["kernel_ms_per_step"].items() if b}
print(f"[{sys.argv[1]}] N={d['n_gpus']} dev {d['ms_per_step']:.4f} ms  wall {d['wall_ms_per_step']:.4f}  e2e {d['e2e']['ms_per_step']:.3f}  launches/frame {d['gpu_launches'] / d['steps']:.0f} matches_1gpu {d.get('matches_1gpu')}")
print("   kernels(rank0):", k, "sum", round(sum(k.values()), 4), "waits:", d.get("peer_wait_ms_per_step_rank0"))
print("   strips", d["run"].get("strip_rows"), "last balance:", (d.get("balance_rounds") or [None])[-1])
PY
  t
}
run default A=1
run noearly PRC_PEER_NO_EARLY_PUSH=1
run traced PRC_PEER_TRACE=1
timeout 120 python tools/group_bench.py --devices 0,1,2,3,4,5,6,7 --steps 40 > $OUT/n8_group.json 2> $OUT/n8_group.err; cut -c1-150 $OUT/n8_group.json; python -c "
import json; d=json.load(open('$OUT/n8_group.json')); print('[group8] dev', round(d['ms_per_frame_device_resident_wall'],4), 'e2e', round(d['e2e_ms_per_frame'],4), 'match', d['matches_1gpu'], d['strip_rows'])" || tail -3 $OUT/n8_group.err; t
timeout 150 python tools/multiview_bench.py --views 256 --repeat 2 --devices 0,1,2,3,4,5,6,7 > $OUT/n8_multiview_group.json 2> $OUT/n8_multiview_group.err; cut -c1-330 $OUT/n8_multiview_group.json; tail -1 $OUT/n8_multiview_group.err | cut -c1-200; t
timeout 100 python -m pytest tests/test_gpu_group.py -m gpu -q -x -k "equals_the_one_context" 2>&1 | tail -2; t
