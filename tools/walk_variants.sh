#!/bin/bash
# Times the grid walk kernel variants (FP_WALK_VARIANT) on workload $1 (default c3).
W=${1:-c3}; shift
for v in "$@"; do
  FP_WALK_VARIANT=$v python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); t=d['timing']; print('variant $v', '$W', 'ms/step %.4f' % d['ms_per_step'], 'walk %.4f' % t['influence_ms_per_step'], 'sort %.4f' % t['sort_phase_ms_per_step'])"
done
