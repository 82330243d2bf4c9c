#!/bin/bash
# round 2, GPU call D (2 GPUs): PDL between the step kernels, graph ring, spawner one step ahead, sharded phase breakdown
set -u
O=gpurun_out/r02d; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "not 3600 and not bench_window and not fuzzed" 2>&1 | tail -8 ) > $O/pytest_sanity.log
( timeout 300 python bench.py --steps 200 --no-sweep 2>&1 | tail -3 ) > $O/bench_n1.log
( CITYFLOW_B200_NO_PDL=1 timeout 300 python bench.py --steps 200 --no-cpu-baseline --no-parity 2>&1 | tail -3 ) > $O/bench_n1_nopdl.log
( CITYFLOW_B200_NO_AHEAD=1 timeout 300 python bench.py --steps 200 --no-cpu-baseline --no-parity 2>&1 | tail -3 ) > $O/bench_n1_noahead.log
( timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "30x60" 2>&1 | tail -12 ) > $O/pytest_multi.log
( timeout 400 $TR --master-port 29551 bench.py --gpus 2 --steps 200 2>&1 | tail -4 ) > $O/bench_n2_weak.log
( timeout 400 $TR --master-port 29552 bench.py --gpus 2 --steps 200 --multi strong --no-parity 2>&1 | tail -4 ) > $O/bench_n2_strong.log
for f in $O/*.log; do echo "=== $f"; tail -c 1200 $f; done
