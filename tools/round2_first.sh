#!/bin/bash
# First GPU call of the next round (one B200, ~12 min): put the candidate-list walk
# (FP_WALK_VARIANT=41) through everything the production walk has been through,
# and profile it.  Outputs in gpurun_out/r2_*.
#   gpurun --timeout 900 -- 'bash tools/round2_first.sh'
# Then, on two GPUs:  gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_first.sh sharded'
set -u
O=gpurun_out
mkdir -p $O
if [ "${1:-}" = "sharded" ]; then
  FP_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -x -q -k sharded > $O/r2_nl_sharded_tests.log 2>&1
  tail -3 $O/r2_nl_sharded_tests.log
  for v in 31 42; do
    FP_WALK_VARIANT=$v FP_NL_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 > $O/r2_bench_c4_g2_v$v.json 2> $O/r2_bench_c4_g2_v$v.err
  done
  exit 0
fi
# 1. the opt-in suite: grid + rebin tests on the variant, state hashes, transitions, overflow
FP_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -x -q > $O/r2_nl_tests.log 2>&1
tail -3 $O/r2_nl_tests.log
# 2. bench lines, production vs lists
for w in c4 c3 c5; do
  for v in 31 41 43 44 45 46 47; do  # (each line ~10-25 s)
    FP_WALK_VARIANT=$v python bench.py --workload $w --no-cpu-baseline > $O/r2_bench_${w}_v$v.json 2>> $O/r2.err
  done
done
for w in c3 c5; do   # the grid centred on the flock (no sliver rows; C5 has 139 sliver-row CTAs when anchored)
  for v in 31 41; do
    FP_GRID_CENTER=1 FP_WALK_VARIANT=$v python bench.py --workload $w --no-cpu-baseline --no-e2e > $O/r2_bench_${w}_centred_v$v.json 2>> $O/r2.err
  done
done
# 3. launch list and full captures of the two new kernels (source page: tools/ncu_source.py)
FP_WALK_VARIANT=41 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/r2_c4_nl_launches.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r2.err
FP_WALK_VARIANT=41 ncu --set full --clock-control none --import-source on -k regex:nl_walk -s 5 -c 1 -f \
    -o $O/r2_prof_nl_walk_c4 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r2.err
FP_WALK_VARIANT=41 ncu --set full --clock-control none --import-source on -k regex:nl_build -s 0 -c 1 -f \
    -o $O/r2_prof_nl_build_c4 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r2.err
tail -5 $O/r2.err
ls -la $O | tail -20
