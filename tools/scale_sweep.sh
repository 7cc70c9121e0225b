#!/bin/bash
# Run on a multi-GPU box under gpurun --gpus N: bit-identity check at N, then bench.py at the given GPU counts.
# usage: tools/scale_sweep.sh "8 4 2" [steps]
OUT=gpurun_out
STEPS=${2:-20}
N0=$(echo $1 | cut -d' ' -f1)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N0 --master-addr 127.0.0.1 --master-port 29540 tests/multigpu_check.py 2>&1 | grep multigpu_check
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) bench.py --gpus $n --steps $STEPS --warmup 5 --no-cpu-baseline > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  fi
  python tools/bench_brief.py N=$n < $OUT/scale_$n.json
done
