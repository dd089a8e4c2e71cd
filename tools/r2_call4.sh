#!/bin/bash
# Round 2, fourth GPU call (one B200): suite on the visible-first fast walk, the lane-parallel small
# kernel, the dense-flock bypass; bench lines (fast numerics is the bench default now); ncu.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q -s > $O/r2d_tests.log 2>&1
grep -E "passed|failed|State<Point>|C2 fast|Error|error" $O/r2d_tests.log | tail -15
for w in c4 c3 c5 c2 c1; do
  python bench.py --workload $w --no-cpu-baseline > $O/r2d_bench_$w.json 2>> $O/r2d.err
done
cat $O/r2d_bench_*.json | python tools/bench_brief.py | cut -c1-330
ncu --set full --clock-control none --import-source on -k regex:nl_fast -s 5 -c 1 -f \
    -o $O/r2d_prof_nl_fast_c4 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2d.err
ncu --set full --clock-control none --import-source on -k regex:nl_walk_kernel -s 5 -c 1 -f \
    -o $O/r2d_prof_nl_walk_c4 python bench.py --numerics exact --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2d.err
tail -5 $O/r2d.err
