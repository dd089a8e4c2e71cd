#!/bin/bash
# Round 2, GPU call 8 (one B200): the whole GPU suite on the kernels as they now stand (fast walk with
# explicit shared addressing and batch-wide open flag, no L2 warming; exact list walk with the packed
# pre-gate; demo-sized flocks through mapped host memory), bench lines, ncu of both list walks.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q -s > $O/r2g_tests.log 2>&1
grep -E "passed|failed|Error|error" $O/r2g_tests.log | tail -8
for w in c4 c1 c2; do
  python bench.py --workload $w --no-cpu-baseline > $O/r2g_bench_$w.json 2>> $O/r2g.err
done
cat $O/r2g_bench_*.json | python tools/bench_brief.py | cut -c1-330
ncu --set full --clock-control none --import-source on -k regex:nl_fast -s 5 -c 1 -f \
    -o $O/r2g_prof_nl_fast_c4 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2g.err
ncu --set full --clock-control none --import-source on -k regex:nl_walk -s 5 -c 1 -f \
    -o $O/r2g_prof_nl_walk_c4 python bench.py --numerics exact --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2g.err
tail -5 $O/r2g.err
