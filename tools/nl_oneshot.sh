#!/bin/bash
# GPU: first contact of the standing-candidate-list walk (FP_WALK_VARIANT=41) with hardware.
# Correctness first (state hashes against the production walk, variant 31), then timing.
# Every run is bounded; output goes to gpurun_out/nl_oneshot.log.
mkdir -p gpurun_out
L=gpurun_out/nl_oneshot.log
: > $L
run() {  # variant, args...
  local v=$1; shift
  FP_WALK_VARIANT=$v FP_NL_TRACE=1 timeout 45 python tools/nl_state_hash.py "$@" >> $L 2>&1 || echo "variant $v args $* -> exit $?" >> $L
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L 2>&1
if [ "$1" = "quick" ]; then   # re-check after a fix: the two cases that exercise CTAs without lists
  run 31 1048576 816 600 11
  run 41 1048576 816 600 11
  run 41 1048576 816 600 11
  run 31 60000 315 80 5 6000
  run 41 60000 315 80 5 6000
  FP_WALK_VARIANT=41 timeout 55 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/nl_bench_c4.json 2>> $L
  cat $L; exit 0
fi
run 31 200000 470 120 7
run 41 200000 470 120 7
run 31 1048576 816 600 11
run 41 1048576 816 600 11
run 31 60000 315 80 5 6000
run 41 60000 315 80 5 6000
run 31 16777216 2048 100 3
run 41 16777216 2048 100 3
cat $L
