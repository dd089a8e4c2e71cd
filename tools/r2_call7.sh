#!/bin/bash
# Round 2, GPU call 7 (one B200): the fast walk with packed FP32 (FADD2 / FMUL2 / FFMA2), the
# positions-only tile form for CTAs whose intervals overflow the full tile, list words four
# batches ahead and L2 warming for the CTA that comes next.  Parity suites that exercise it,
# bench lines with and without the L2 warming, ncu of the new kernel.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_fast.py tests/test_gpu_lists.py tests/test_gpu_scale.py tests/test_gpu_grid.py tests/test_gpu_rebin.py -m gpu -x -q -s > $O/r2f_tests.log 2>&1
grep -E "passed|failed|Error|error" $O/r2f_tests.log | tail -8
python bench.py --workload c4 --no-cpu-baseline > $O/r2f_bench_c4.json 2>> $O/r2f.err
FP_NL_TRACE=1 FP_NL_PREFETCH=0 python bench.py --workload c4 --no-cpu-baseline --no-alt --no-e2e > $O/r2f_bench_c4_nopf.json 2>> $O/r2f.err
for w in c3 c5; do
  python bench.py --workload $w --no-cpu-baseline > $O/r2f_bench_$w.json 2>> $O/r2f.err
done
cat $O/r2f_bench_*.json | python tools/bench_brief.py | cut -c1-330
ncu --set full --clock-control none --import-source on -k regex:nl_fast -s 5 -c 1 -f \
    -o $O/r2f_prof_nl_fast_c4 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > /dev/null 2>> $O/r2f.err
tail -5 $O/r2f.err
