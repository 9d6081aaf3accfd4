#!/bin/bash
# 2-GPU session: multi-rank bench (NCCL reduce of moments), reference arm, default bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_j.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_j.log
timeout 900 python bench.py > gpurun_out/bench_j1.log 2>&1; echo "bench1 rc=$?"; tail -c 2200 gpurun_out/bench_j1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_j2.log 2>&1; echo "bench2 rc=$?"; tail -c 1500 gpurun_out/bench_j2.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_jref.log 2>&1; echo "benchref rc=$?"; tail -c 1200 gpurun_out/bench_jref.log
