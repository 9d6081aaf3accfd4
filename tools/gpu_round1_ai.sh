#!/bin/bash
# first GPU run of the radiance-predicting network: diagnostics, parity tests, device time of both kernels
mkdir -p gpurun_out
timeout 300 python tools/debug_mlp.py 300 > gpurun_out/debug_mlp_ai.log 2>&1; echo "debug rc=$?"; tail -20 gpurun_out/debug_mlp_ai.log
timeout 600 python -m pytest tests/test_disney_mlp.py -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/pytest_mlp_ai.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_mlp_ai.log
timeout 300 python tools/bench_mlp.py > gpurun_out/bench_mlp_ai.log 2>&1; echo "bench rc=$?"; cat gpurun_out/bench_mlp_ai.log | cut -c1-400
