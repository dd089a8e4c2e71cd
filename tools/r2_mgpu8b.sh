#!/bin/bash
# Round 2, eight B200s, second call (charged 8x: the two items the first call left open): the FAST
# sharded parity test at 8 ranks on its corrected scenario, and the 4-GPU point of X1 (C5 all-pairs,
# all-gather sharded).
set -u
O=gpurun_out
mkdir -p $O
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
MGPU_ONLY=fast run 8 29628 tests/mgpu_check.py > $O/r2_mgpu_check_g8_fast.log 2>&1
grep -E "mgpu|MGPU_OK|Error|error|assert" $O/r2_mgpu_check_g8_fast.log | tail -6
run 4 29634 bench.py --gpus 4 --workload c5 --method allpairs --steps 3 --warmup 1 --no-alt > $O/r2_x1_c5_allpairs_g4.json 2>> $O/r2_mgpu8b.err
grep "^{" $O/r2_x1_c5_allpairs_g4.json | python tools/bench_brief.py | cut -c1-300
tail -3 $O/r2_mgpu8b.err
