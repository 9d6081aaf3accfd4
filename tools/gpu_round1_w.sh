#!/bin/bash
# 8 GPUs of one box: the scaling bench (weak scaling, one NCCL reduce of the moment buffers) and C3 dataset generation sharded by scene
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "bench n8 rc=$?"; tail -c 1500 gpurun_out/bench_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_dataset.py --scenes-per-gpu 1 > gpurun_out/c3_n8.log 2>&1; echo "c3 n8 rc=$?"; tail -c 1500 gpurun_out/c3_n8.log
