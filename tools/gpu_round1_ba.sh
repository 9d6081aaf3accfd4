#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_disney_mlp.py -m gpu -q --timeout 240 --timeout-method thread 2>&1 | tail -12
timeout 300 python tools/bench_mlp.py > gpurun_out/bench_mlp_ba.log 2>&1; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench_mlp_ba.log
timeout 300 python tools/bench_mlp.py 262144 2>&1 | grep -v block0 | cut -c1-200
