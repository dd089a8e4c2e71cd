#!/bin/bash
# Round 2, GPU call 9 (one B200): the whole GPU suite, the launch list of the default bench command,
# bench lines of the remaining workloads.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -s > $O/r2h_tests.log 2>&1
grep -E "passed|failed|Error|error" $O/r2h_tests.log | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2h_c4_launches.csv \
    python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2h.err
for w in c3 c5; do
  python bench.py --workload $w --no-cpu-baseline > $O/r2h_bench_$w.json 2>> $O/r2h.err
done
python bench.py > $O/r2h_bench_default.json 2>> $O/r2h.err
python bench.py --impl reference > $O/r2h_bench_reference.json 2>> $O/r2h.err
cat $O/r2h_bench_*.json | python tools/bench_brief.py | cut -c1-330
tail -5 $O/r2h.err
