#!/bin/bash
# round 2, GPU call A: first execution of the lane-change kernels + dead-end stop, fuzz, sanitizer, bench, shard loop-back vs reference
set -u
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/pytest_gpu.log
( timeout 400 python tools/lc_gpu_check.py 4 4 600 2>&1 | tail -30 ) > $O/lc_4x4.log
( timeout 400 python tools/lc_gpu_check.py 6 6 500 2>&1 | tail -30 ) > $O/lc_6x6.log
( CITYFLOW_B200_LC_SERIAL=1 timeout 400 python tools/lc_gpu_check.py 4 4 300 2>&1 | tail -30 ) > $O/lc_4x4_serial.log
( timeout 600 python tools/gpu_fuzz_check.py 1 13 2>&1 | tail -40 ) > $O/fuzz.log
( timeout 300 python bench.py --steps 200 --no-cpu-baseline 2>&1 | tail -3 ) > $O/bench_n1.log
( timeout 900 python tools/shard_loopback_ref_check.py 30 60 2 1225 25 2>&1 | tail -40 ) > $O/shard_loopback_30x60.log
( timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_run.py 60 2>&1 | tail -40 ) > $O/racecheck.log
tail -5 $O/*.log
