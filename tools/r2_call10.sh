#!/bin/bash
# Round 2, GPU call 10 (one B200): the all-pairs / single-CTA suite with the mapped-memory test, bench
# lines whose parity block changed (C1: lead rows pushed before the tap; C2: the reference's order
# sensitivity beside the bar), the 1-GPU point of X1, the default line with roofline.traffic.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_allpairs.py tests/test_cpp_host.py -m gpu -q -s > $O/r2i_tests.log 2>&1
grep -E "passed|failed|Error|error" $O/r2i_tests.log | tail -8
python bench.py --workload c1 > $O/r2i_bench_c1.json 2>> $O/r2i.err
python bench.py --workload c2 --no-cpu-baseline > $O/r2i_bench_c2.json 2>> $O/r2i.err
python bench.py --workload c5 --method allpairs --steps 3 --warmup 1 --no-alt --no-cpu-baseline > $O/r2i_x1_c5_allpairs_g1.json 2>> $O/r2i.err
python bench.py > $O/r2i_bench_default.json 2>> $O/r2i.err
cat $O/r2i_bench_*.json $O/r2i_x1_*.json | python tools/bench_brief.py | cut -c1-330
python -c "
import json
for f in ['c1','c2']:
    d=json.loads([l for l in open('gpurun_out/r2i_bench_%s.json'%f) if l.startswith('{')][-1]); print(f, d['parity'])
"
tail -5 $O/r2i.err
