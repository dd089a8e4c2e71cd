#!/bin/bash
# Round 2, two B200s, second call: the sharded parity suite on the kernels as they now stand, and the
# 2-GPU points of both sharded workloads with them.
set -u
O=gpurun_out
mkdir -p $O
RUN="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$RUN --master-port 29711 tests/mgpu_check.py > $O/r2_mgpu_check_g2_final.log 2>&1
grep -E "mgpu|MGPU_OK|Error|error|assert" $O/r2_mgpu_check_g2_final.log | tail -12
$RUN --master-port 29713 bench.py --gpus 2 > $O/r2_scale_c4_g2_final.json 2>> $O/r2_mgpu2b.err
$RUN --master-port 29714 bench.py --gpus 2 --workload c5 --method allpairs --steps 3 --warmup 1 --no-alt > $O/r2_x1_c5_allpairs_g2_final.json 2>> $O/r2_mgpu2b.err
cat $O/r2_scale_c4_g2_final.json $O/r2_x1_c5_allpairs_g2_final.json | grep "^{" | python tools/bench_brief.py | cut -c1-330
tail -3 $O/r2_mgpu2b.err
