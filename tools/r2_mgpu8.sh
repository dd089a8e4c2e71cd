#!/bin/bash
# Round 2, eight B200s (charged 8x: only what needs the box): sharded parity at 4 and 8 ranks,
# C4 slab-sharded at 4 and 8, C5 all-pairs all-gather sharded (X1) at 4 and 8.
# The 1- and 2-GPU points of the same lines come from the one- and two-GPU calls.
set -u
O=gpurun_out
mkdir -p $O
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for n in 8 4; do
  run $n 2952$n tests/mgpu_check.py > $O/r2_mgpu_check_g$n.log 2>&1
  grep -E "mgpu|MGPU_OK|Error|error|assert" $O/r2_mgpu_check_g$n.log | tail -10
done
for n in 8 4; do run $n 2953$n bench.py --gpus $n > $O/r2_scale8_c4_g$n.json 2>> $O/r2_mgpu8.err; done
for n in 8 4; do run $n 2954$n bench.py --gpus $n --workload c5 --method allpairs --steps 3 --warmup 1 --no-alt > $O/r2_x1_c5_allpairs_g$n.json 2>> $O/r2_mgpu8.err; done
cat $O/r2_scale8_c4_g*.json $O/r2_x1_c5_allpairs_g[48].json | python tools/bench_brief.py | cut -c1-300
tail -5 $O/r2_mgpu8.err
