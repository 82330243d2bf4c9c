#!/bin/bash
# round 2, GPU call H (8 GPUs): the collective count over peer memory (2-rank parity test), then the 8-GPU weak and strong lines
set -u
O=gpurun_out/r02h; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 240 python -m pytest tests/test_gpu_multi.py -x -q -k "30x60" 2>&1 | tail -10 ) > $O/pytest_multi2.log
( timeout 420 $TR --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 --no-parity 2>&1 | tail -4 ) > $O/bench_n8_weak.log
( timeout 200 $TR --master-port 29582 bench.py --gpus 8 --steps 50 --multi strong --no-parity 2>&1 | tail -4 ) > $O/bench_n8_strong.log
for f in $O/*.log; do echo "=== $f"; tail -c 1500 $f; done
