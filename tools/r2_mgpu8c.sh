#!/bin/bash
# Round 2, eight B200s, third call: the 8-GPU point of C4 on the final kernels (nothing else: 8x charge).
set -u
O=gpurun_out
mkdir -p $O
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29848 bench.py --gpus 8 --no-alt > $O/r2_scale_c4_g8_final.json 2>> $O/r2_mgpu8c.err
grep "^{" $O/r2_scale_c4_g8_final.json | python tools/bench_brief.py | cut -c1-330
tail -3 $O/r2_mgpu8c.err
