#!/bin/bash
# Round 2, GPU call 6 (one B200): FP32 instruction rates, the suite on the new list layout
# (8-byte entry words, padding entries), bench lines, ncu of the fast walk.
set -u
O=gpurun_out
mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32_rates tools/micro/fp32_rates.cu && /tmp/fp32_rates > $O/r2_fp32_rates.txt 2>&1
cat $O/r2_fp32_rates.txt
python -m pytest tests -m gpu -x -q -s > $O/r2e_tests.log 2>&1
grep -E "passed|failed|Error|error" $O/r2e_tests.log | tail -8
for w in c4 c3 c5 c2; do
  python bench.py --workload $w --no-cpu-baseline > $O/r2e_bench_$w.json 2>> $O/r2e.err
done
cat $O/r2e_bench_*.json | python tools/bench_brief.py | cut -c1-330
ncu --set full --clock-control none --import-source on -k regex:nl_fast -s 5 -c 1 -f \
    -o $O/r2e_prof_nl_fast_c4 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2e.err
tail -5 $O/r2e.err
