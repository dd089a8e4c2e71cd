#!/bin/bash
# Round 2, eight B200s: sharded parity at 4 and 8 ranks, C4 slab-sharded at 1/2/4/8, C5 all-pairs
# all-gather sharded at 1/2/4/8 (X1).
set -u
O=gpurun_out
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for n in 4 8; do
  run $n 2952$n tests/mgpu_check.py > $O/r2_mgpu_check_g$n.log 2>&1
  grep -E "mgpu|MGPU_OK|Error|error|assert" $O/r2_mgpu_check_g$n.log | tail -10
done
python bench.py --no-cpu-baseline > $O/r2_scale8_c4_g1.json 2>> $O/r2_mgpu8.err
for n in 2 4 8; do run $n 2953$n bench.py --gpus $n > $O/r2_scale8_c4_g$n.json 2>> $O/r2_mgpu8.err; done
python bench.py --workload c5 --method allpairs --steps 3 --warmup 1 --no-cpu-baseline > $O/r2_x1_c5_allpairs_g1.json 2>> $O/r2_mgpu8.err
for n in 2 4 8; do run $n 2954$n bench.py --gpus $n --workload c5 --method allpairs --steps 3 --warmup 1 > $O/r2_x1_c5_allpairs_g$n.json 2>> $O/r2_mgpu8.err; done
cat $O/r2_scale8_c4_g*.json $O/r2_x1_c5_allpairs_g*.json | python tools/bench_brief.py | cut -c1-300
tail -5 $O/r2_mgpu8.err
