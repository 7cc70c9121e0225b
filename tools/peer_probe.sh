#!/bin/bash
# Multi-GPU probe of the peer-memory path (run under gpurun --gpus N): bit-identity check, then bench.py --mgpu peer in a few
# variants (default library, traced, tuning builds given as extra arguments: name=path pairs).
# usage: tools/peer_probe.sh N [steps] [name=ENV=VAL,ENV=VAL ...]   e.g. persist=PRC_LIB=polyred_b200/csrc/variants/lib_persist.so
N=$1; STEPS=${2:-40}; shift 2
OUT=gpurun_out
mkdir -p $OUT
run() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline --mgpu peer > $OUT/peer${N}_$tag.json 2> $OUT/peer${N}_$tag.err
  python - "$tag" $OUT/peer${N}_$tag.json <<'PY' || tail -5 $OUT/peer${N}_$tag.err
import json, sys
d = json.load(open(sys.argv[2]))
k = {a: round(b, 4) for a, b in d["kernel_ms_per_step"].items()}
print(f"[{sys.argv[1]}] N={d['n_gpus']} dev {d['ms_per_step']:.4f} ms  wall {d['wall_ms_per_step']:.4f}  e2e {d['e2e']['ms_per_step']:.3f}  launches/frame {d['gpu_launches'] / d['steps']:.0f}")
print("   kernels(rank0):", k, "sum", round(sum(k.values()), 4))
print("   waits:", d.get("peer_wait_ms_per_step_rank0"), "strips", d["run"].get("strip_rows"), "matches_1gpu", d.get("matches_1gpu"))
b = d.get("balance_rounds") or []
if b: print("   last balance:", b[-1])
PY
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 tests/multigpu_check.py > $OUT/multigpu_check_$N.log 2>&1
grep "multigpu_check\|Error\|error" $OUT/multigpu_check_$N.log | head -20
run default A=1
run traced PRC_PEER_TRACE=1
for kv in "$@"; do   # name=ENV1=V1,ENV2=V2 (a value starting with lib: is a library path relative to the repo)
  name=${kv%%=*}; envs=${kv#*=}
  run "$name" $(echo "$envs" | tr ',' ' ' | sed "s#PRC_LIB=#PRC_LIB=$PWD/#")
done
