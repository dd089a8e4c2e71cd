#!/bin/bash
# Round artefacts on one B200: bench lines for every SURVEY 8d workload, the reference arm,
# the ncu launch list and one full capture of the dominant kernel.  Outputs in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/fin_bench_c4.json 2> $O/fin.err
python bench.py --workload c3 > $O/fin_bench_c3.json 2>> $O/fin.err
python bench.py --workload c5 > $O/fin_bench_c5.json 2>> $O/fin.err
python bench.py --workload c2 > $O/fin_bench_c2.json 2>> $O/fin.err
python bench.py --workload c1 > $O/fin_bench_c1.json 2>> $O/fin.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/fin_bench_reference_c4.json 2>> $O/fin.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/fin_c4_launches.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/fin.err
ncu --set full --clock-control none --import-source on -k regex:grid_walk3 -s 5 -c 1 -f -o $O/fin_prof_walk_c4 \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/fin.err
ncu --set full --clock-control none --import-source on -k regex:allpairs -s 4 -c 1 -f -o $O/fin_prof_allpairs_c2 \
    python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/fin.err
tail -5 $O/fin.err
ls -la $O | tail -12
