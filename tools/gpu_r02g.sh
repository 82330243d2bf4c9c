#!/bin/bash
# round 2, GPU call G (1 GPU): final single-GPU state -- full GPU suite, bench (thread sweep, parity), lane-change line,
# launch list + full ncu capture of the step kernels, bandwidth-regime point
set -u
O=gpurun_out/r02g; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 ) > $O/pytest_gpu.log
( timeout 400 python bench.py --steps 200 2>&1 | tail -3 ) > $O/bench_n1.log
( timeout 300 python bench.py --steps 100 --lane-change --no-cpu-baseline 2>&1 | tail -3 ) > $O/bench_n1_lc.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python bench.py --steps 3 --profile-steps 4 --no-cpu-baseline --no-parity > $O/ncu_launches.out 2>&1 )
( timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_(ingest|notify|control|move|leader)" -c 10 -o $O/prof_step python bench.py --steps 3 --profile-steps 2 --no-cpu-baseline --no-parity > $O/ncu_full.out 2>&1 )
( timeout 400 python bench.py --rows 100 --cols 100 --steps 30 --no-cpu-baseline --no-parity 2>&1 | tail -3 ) > $O/bench_100x100.log
for f in $O/pytest_gpu.log $O/bench_n1.log $O/bench_n1_lc.log $O/bench_100x100.log; do echo "=== $f"; tail -c 1500 $f; done
