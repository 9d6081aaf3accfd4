#!/bin/bash
# v3 defaults: tests + bench line; first C3 dataset-generation measurement (2 scenes, reference collector settings, then larger thread counts)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_p.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_p.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_p.log
timeout 600 python tools/bench_dataset.py --scenes 2 > gpurun_out/c3_a.log 2>&1; echo "c3 rc=$?"; tail -c 1800 gpurun_out/c3_a.log
timeout 600 python tools/bench_dataset.py --scenes 2 --max-threads 163840 > gpurun_out/c3_b.log 2>&1; echo "c3 rc=$?"; tail -c 1800 gpurun_out/c3_b.log
timeout 600 python tools/bench_dataset.py --scenes 2 --max-threads 655360 --launches 50 > gpurun_out/c3_c.log 2>&1; echo "c3 rc=$?"; tail -c 1800 gpurun_out/c3_c.log
