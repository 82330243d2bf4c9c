#!/bin/bash
# round 2, GPU call F (4 GPUs): ranks with two neighbours -- 4-rank parity tests, weak / strong bench lines, RL replicas
set -u
O=gpurun_out/r02f; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -k "four" 2>&1 | tail -15 ) > $O/pytest_multi4.log
( timeout 400 $TR --master-port 29571 bench.py --gpus 4 --steps 100 2>&1 | tail -4 ) > $O/bench_n4_weak.log
( timeout 300 $TR --master-port 29572 bench.py --gpus 4 --steps 100 --multi strong 2>&1 | tail -4 ) > $O/bench_n4_strong.log
( timeout 300 $TR --master-port 29573 bench.py --gpus 4 --steps 300 --config rl 2>&1 | tail -4 ) > $O/bench_n4_rl.log
( timeout 300 python bench.py --impl reference --gpus 4 --steps 300 --config rl 2>&1 | tail -2 ) > $O/bench_n4_rl_reference.log
for f in $O/*.log; do echo "=== $f"; tail -c 1500 $f; done
