#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
g() { local tag=$1; shift; env "$@" timeout 150 python tools/group_bench.py --devices 0,1 --steps 40 > $OUT/f_group_$tag.json 2> $OUT/f_group_$tag.err; echo "[group $tag] $(cut -c1-420 $OUT/f_group_$tag.json)"; tail -1 $OUT/f_group_$tag.err | cut -c1-300; }
g default A=1
g nofuse PRC_PEER_NO_FUSED_SIGNALS=1
bash tools/peer_probe.sh 2 30 onestream=PRC_PEER_ONE_STREAM=1 nofuse=PRC_PEER_NO_FUSED_SIGNALS=1
