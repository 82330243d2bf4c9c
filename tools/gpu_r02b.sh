#!/bin/bash
# round 2, GPU call B (2 GPUs): peer-memory seam transport -- loop-back on one GPU, 2-rank parity vs reference, weak/strong/NCCL bench lines; N=1 bench with thread sweep; CFB_CONTROL_COOP timing
set -u
O=gpurun_out/r02b; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1; nvidia-smi topo -m >> $O/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "3x3 or sharded or lane_change or 6x6_dense_vs_port" 2>&1 | tail -15 ) > $O/pytest_sanity.log
( timeout 600 python tools/shard_loopback_ref_check.py 8 12 3 400 25 2>&1 | tail -12 ) > $O/loopback_p2p_8x12.log
( timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -40 ) > $O/pytest_multi.log
( timeout 900 $TR --master-port 29541 bench.py --gpus 2 --steps 200 2>&1 | tail -4 ) > $O/bench_n2_weak_p2p.log
( timeout 900 $TR --master-port 29542 bench.py --gpus 2 --steps 200 --multi strong 2>&1 | tail -4 ) > $O/bench_n2_strong_p2p.log
( CITYFLOW_B200_SHARD_TRANSPORT=nccl timeout 900 $TR --master-port 29543 bench.py --gpus 2 --steps 200 --no-parity 2>&1 | tail -4 ) > $O/bench_n2_weak_nccl.log
( timeout 900 python bench.py --steps 200 2>&1 | tail -3 ) > $O/bench_n1.log
cp cityflow_b200/libcityflow_b200.so /tmp/lib_default.so
cp cityflow_b200/csrc/build_coop/libcityflow_b200.so cityflow_b200/libcityflow_b200.so
( timeout 600 python bench.py --steps 200 --no-cpu-baseline --no-parity 2>&1 | tail -3 ) > $O/bench_n1_coop.log
( timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "6x6_dense_vs_port or 3x3" 2>&1 | tail -5 ) > $O/pytest_coop.log
cp /tmp/lib_default.so cityflow_b200/libcityflow_b200.so
for f in $O/*.log; do echo "=== $f"; tail -c 1500 $f; done
