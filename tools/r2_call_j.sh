#!/bin/bash
# 2 GPUs: the peer path after the key-plane change, across processes (multigpu_check incl. the moving-camera leg, through pytest), the
# one-process group on two devices, and the N=2 bench line
OUT=gpurun_out; mkdir -p $OUT; T0=$(date +%s)
t() { echo "  (t+$(( $(date +%s) - T0 )) s)"; }
PRC_MULTIGPU_LOG=$OUT/multigpu_check_2.log timeout 400 python -m pytest tests/test_gpu_group.py tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -30 > $OUT/j_group_peer_2gpu.log; tail -3 $OUT/j_group_peer_2gpu.log; t
grep "multigpu_check" $OUT/multigpu_check_2.log | cut -c1-220; t
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline > $OUT/j_bench2.json 2> $OUT/j_bench2.err
python tools/bench_brief.py n2 < $OUT/j_bench2.json || tail -5 $OUT/j_bench2.err; t
timeout 120 python tools/group_bench.py --devices 0,1 --steps 40 > $OUT/j_group2.json 2> $OUT/j_group2.err; cut -c1-100 $OUT/j_group2.json; python -c "
import json; d=json.load(open('$OUT/j_group2.json')); print('[group2] dev', round(d['ms_per_frame_device_resident_wall'],4), 'e2e', round(d['e2e_ms_per_frame'],4), 'match', d['matches_1gpu'], d['strip_rows'])" || tail -3 $OUT/j_group2.err; t
