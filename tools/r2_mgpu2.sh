#!/bin/bash
# Round 2, two B200s: the sharded parity suite (tests/mgpu_check.py: all-pairs all-gather bit-exact,
# slabs with candidate lists, lazy halo over peer stores and over NCCL, FAST numerics, write_local)
# and bench lines of both sharded workloads.
set -u
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
FP_NL_TRACE=1 $RUN --master-port 29511 tests/mgpu_check.py > $O/r2_mgpu_check_g2.log 2>&1
grep -E "mgpu|MGPU_OK|Error|error|assert" $O/r2_mgpu_check_g2.log | tail -12
MGPU_ONLY=lazy FP_SHARD_PEER=0 $RUN --master-port 29512 tests/mgpu_check.py > $O/r2_mgpu_check_g2_nccl_halo.log 2>&1
grep -E "mgpu|MGPU_OK|Error|error|assert" $O/r2_mgpu_check_g2_nccl_halo.log | tail -4
python bench.py --no-cpu-baseline > $O/r2_scale_c4_g1.json 2>> $O/r2_mgpu2.err
$RUN --master-port 29513 bench.py --gpus 2 > $O/r2_scale_c4_g2.json 2>> $O/r2_mgpu2.err
$RUN --master-port 29514 bench.py --gpus 2 --workload c5 --method allpairs --steps 3 --warmup 1 > $O/r2_x1_c5_allpairs_g2.json 2>> $O/r2_mgpu2.err
cat $O/r2_scale_c4_g1.json $O/r2_scale_c4_g2.json $O/r2_x1_c5_allpairs_g2.json | python tools/bench_brief.py | cut -c1-330
tail -5 $O/r2_mgpu2.err
