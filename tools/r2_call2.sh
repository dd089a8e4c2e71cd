#!/bin/bash
# Round 2, second GPU call (one B200): the whole -m gpu suite on the new defaults (candidate lists,
# centred grid), FAST numerics tests, parity at C2/C4/C5, bench lines, ncu of the fast list walk.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2b_tests.log 2>&1
tail -15 $O/r2b_tests.log
for spec in "c4 exact" "c4 fast" "c3 fast" "c5 fast" "c2 fast" "c1 exact"; do
  set -- $spec
  python bench.py --workload $1 --numerics $2 --no-cpu-baseline > $O/r2b_bench_$1_$2.json 2>> $O/r2b.err
  tail -c 600 $O/r2b_bench_$1_$2.json | head -c 300; echo
done
ncu --set full --clock-control none --import-source on -k regex:nl_fast -s 5 -c 1 -f \
    -o $O/r2b_prof_nl_fast_c4 python bench.py --numerics fast --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2b.err
ncu --set full --clock-control none --import-source on -k regex:allpairs_fast -s 2 -c 1 -f \
    -o $O/r2b_prof_apf_c2 python bench.py --workload c2 --numerics fast --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2b.err
tail -5 $O/r2b.err
ls -la $O | tail -12
