#!/bin/bash
# Round 2, GPU call 11 (one B200): smoke(), the all-pairs / single-CTA suite, and the C1 bench line
# after fp_flock_set_leads stopped forcing pending rows onto the device.
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2j_smoke.log 2>&1; tail -2 $O/r2j_smoke.log
python -m pytest tests/test_gpu_allpairs.py tests/test_cpp_host.py tests/test_gpu_state.py -m gpu -q -s > $O/r2j_tests.log 2>&1
grep -E "passed|failed|Error|error" $O/r2j_tests.log | tail -8
python bench.py --workload c1 > $O/r2j_bench_c1.json 2>> $O/r2j.err
python bench.py --workload c1 > $O/r2j_bench_c1_again.json 2>> $O/r2j.err
cat $O/r2j_bench_c1*.json | python tools/bench_brief.py | cut -c1-400
python -c "
import json
for f in ['c1','c1_again']:
    d=json.loads([l for l in open('gpurun_out/r2j_bench_%s.json'%f) if l.startswith('{')][-1]); print(f, d['parity']['max_rel_accel'], d['e2e']['value'], d['e2e']['frame']['value'], d['cpu_baseline']['value'])
"
tail -3 $O/r2j.err
