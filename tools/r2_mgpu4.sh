#!/bin/bash
# Round 2, four B200s: the 4-GPU point of C4 on the final kernels.
set -u
O=gpurun_out
mkdir -p $O
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29844 bench.py --gpus 4 --no-alt > $O/r2_scale_c4_g4_final.json 2>> $O/r2_mgpu4.err
grep "^{" $O/r2_scale_c4_g4_final.json | python tools/bench_brief.py | cut -c1-330
tail -3 $O/r2_mgpu4.err
