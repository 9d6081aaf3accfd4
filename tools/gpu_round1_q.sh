#!/bin/bash
# datagen (C++ host) GPU tests + first C3 dataset-generation measurements
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_q.log
timeout 600 python tools/bench_dataset.py --scenes 2 > gpurun_out/c3_a.log 2>&1; echo "c3 rc=$?"; tail -c 1800 gpurun_out/c3_a.log
timeout 600 python tools/bench_dataset.py --scenes 2 --max-threads 163840 > gpurun_out/c3_b.log 2>&1; echo "c3 rc=$?"; tail -c 1800 gpurun_out/c3_b.log
timeout 600 python tools/bench_dataset.py --scenes 2 --max-threads 655360 --launches 50 > gpurun_out/c3_c.log 2>&1; echo "c3 rc=$?"; tail -c 1800 gpurun_out/c3_c.log
