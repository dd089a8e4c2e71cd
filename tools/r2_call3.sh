#!/bin/bash
# Round 2, third GPU call (one B200): whole -m gpu suite (State<T>, SPH, two-phase fast walk,
# branch-free fast all-pairs), bench lines with segment resets, ncu of the two fast kernels.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q -s > $O/r2c_tests.log 2>&1
grep -E "passed|failed|State<Point>|C2 fast|Error|error" $O/r2c_tests.log | tail -15
for spec in "c4 exact" "c3 exact" "c5 exact" "c2 fast" "c1 exact"; do
  set -- $spec
  python bench.py --workload $1 --numerics $2 --no-cpu-baseline > $O/r2c_bench_$1_$2.json 2>> $O/r2c.err
done
cat $O/r2c_bench_*.json | python tools/bench_brief.py | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:nl_fast -s 5 -c 1 -f \
    -o $O/r2c_prof_nl_fast_c4 python bench.py --numerics fast --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2c.err
ncu --set full --clock-control none --import-source on -k regex:allpairs_fast -s 2 -c 1 -f \
    -o $O/r2c_prof_apf_c2 python bench.py --workload c2 --numerics fast --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2c.err
tail -5 $O/r2c.err
