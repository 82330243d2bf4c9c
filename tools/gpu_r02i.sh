#!/bin/bash
# round 2, GPU call I (4 GPUs): final code -- single-GPU sanity + bench line, then the 4-GPU weak line with and without the
# opt-in four-thread vehicle creation (parity_check on)
set -u
O=gpurun_out/r02i; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 200 -k "not 3600 and not bench_window and not fuzzed" 2>&1 | tail -6 ) > $O/pytest_sanity.log
( timeout 240 python bench.py --steps 100 --no-sweep 2>&1 | tail -3 ) > $O/bench_n1.log
( CITYFLOW_B200_PARALLEL_SPAWN_MIN=512 timeout 400 $TR --master-port 29591 bench.py --gpus 4 --steps 100 2>&1 | tail -4 ) > $O/bench_n4_weak_parallel_creation.log
( timeout 400 $TR --master-port 29592 bench.py --gpus 4 --steps 100 --no-parity 2>&1 | tail -4 ) > $O/bench_n4_weak_default.log
for f in $O/*.log; do echo "=== $f"; tail -c 1200 $f; done
