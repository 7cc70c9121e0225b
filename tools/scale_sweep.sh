#!/bin/bash
# Run on a multi-GPU box under gpurun --gpus N: bit-identity check at N (NCCL exchange AND peer-memory frames), then
# bench.py at the given GPU counts in both multi-GPU modes.
# usage: tools/scale_sweep.sh "8 4 2" [steps] [modes: "nccl peer"]
OUT=gpurun_out
STEPS=${2:-20}
MODES=${3:-"nccl peer"}
N0=$(echo $1 | cut -d' ' -f1)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N0 --master-addr 127.0.0.1 --master-port 29540 tests/multigpu_check.py > $OUT/multigpu_check_$N0.log 2>&1
grep "multigpu_check\|Error\|error" $OUT/multigpu_check_$N0.log | head -20
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline > $OUT/scale_$n.json 2> $OUT/scale_$n.err
    python tools/bench_brief.py N=$n < $OUT/scale_$n.json
    continue
  fi
  for mode in $MODES; do
    tag=$n; [ "$mode" = "peer" ] && tag=${n}_peer
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) bench.py --gpus $n --steps $STEPS --warmup 5 --no-cpu-baseline --mgpu $mode > $OUT/scale_$tag.json 2> $OUT/scale_$tag.err
    python tools/bench_brief.py "N=$n/$mode" < $OUT/scale_$tag.json || tail -5 $OUT/scale_$tag.err
    if [ "$mode" = "peer" ]; then
      # a second, traced run (not a bench value): where rank 0 idles for its peers, and the balance rounds
      PRC_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29570 + n)) bench.py --gpus $n --steps $STEPS --warmup 5 --no-cpu-baseline --mgpu peer > $OUT/scale_${tag}_traced.json 2> $OUT/scale_${tag}_traced.err
      python -c "import json,sys; d=json.load(open('$OUT/scale_${tag}_traced.json')); print('  traced:', round(d['ms_per_step'],4), 'ms; waits', d.get('peer_wait_ms_per_step_rank0'), '; strips', d['config'].get('strip_rows'), '; last balance round', (d.get('balance_rounds') or [None])[-1])" || true
    fi
  done
done
