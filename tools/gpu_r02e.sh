#!/bin/bash
# round 2, GPU call E (2 GPUs): send kernels with one system fence + PDL on the seam kernels; full single-GPU suite on the
# macro-free sources; lane-change bench line; bandwidth-regime point (100x100, 1.5e6 vehicles) with an ncu capture
set -u
O=gpurun_out/r02e; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "30x60" 2>&1 | tail -12 ) > $O/pytest_multi.log
( timeout 300 $TR --master-port 29561 bench.py --gpus 2 --steps 200 2>&1 | tail -4 ) > $O/bench_n2_weak.log
( timeout 300 $TR --master-port 29562 bench.py --gpus 2 --steps 200 --multi strong --no-parity 2>&1 | tail -4 ) > $O/bench_n2_strong.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > $O/pytest_gpu.log
( timeout 400 python bench.py --steps 100 --lane-change --no-cpu-baseline 2>&1 | tail -3 ) > $O/bench_n1_lc.log
( timeout 600 python bench.py --rows 100 --cols 100 --steps 50 --no-cpu-baseline --no-parity 2>&1 | tail -3 ) > $O/bench_100x100.log
( timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_(notify|control|move|leader)" -c 4 -o $O/prof_100x100 python bench.py --rows 100 --cols 100 --steps 3 --profile-steps 1 --no-cpu-baseline --no-parity > $O/ncu_100x100.out 2>&1 )
for f in $O/*.log; do echo "=== $f"; tail -c 1000 $f; done
