#!/bin/bash
# cycle accounting of the tensor-core model kernel
mkdir -p gpurun_out
timeout 300 python tools/debug_mlp.py 300 2>&1 | tail -9; timeout 300 python -m pytest tests/test_disney_mlp.py -m gpu -q --timeout 240 --timeout-method thread 2>&1 | tail -5; timeout 300 python tools/bench_mlp.py > gpurun_out/bench_mlp_al.log 2>&1; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_mlp_al.log
