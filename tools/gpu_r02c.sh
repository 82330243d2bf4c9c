#!/bin/bash
# round 2, GPU call C (1 GPU): k_move+leader fusion, foe terms in k_notify, graph ring -- full GPU suite, bench, ncu launch list + full capture
set -u
O=gpurun_out/r02c; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu.log
( timeout 400 python bench.py --steps 200 --no-sweep 2>&1 | tail -3 ) > $O/bench_n1.log
( timeout 400 python tools/lc_gpu_check.py 4 4 300 2>&1 | tail -5 ) > $O/lc_4x4.log
# launch list (device time per launch, cold cache, serialised) and one full capture of each step kernel
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python bench.py --steps 3 --profile-steps 4 --no-cpu-baseline --no-parity > $O/ncu_launches.out 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_(ingest|notify|control|move|leader)" -c 10 -o $O/prof_step python bench.py --steps 3 --profile-steps 2 --no-cpu-baseline --no-parity > $O/ncu_full.out 2>&1 )
ls -la $O
for f in $O/pytest_gpu.log $O/bench_n1.log $O/lc_4x4.log; do echo "=== $f"; tail -c 2500 $f; done
head -c 1500 $O/launches.csv
