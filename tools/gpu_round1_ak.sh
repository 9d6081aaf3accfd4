#!/bin/bash
# reworked tensor-core kernel (K = 16 chunks, 7 weight stages, z prefetch) + frame-wide neural renderer: parity, device time, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_disney_mlp.py tests/test_gpu_parity.py -k "disney or network_input" -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/pytest_mlp_ak.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_mlp_ak.log
timeout 300 python tools/debug_mlp.py 300 > gpurun_out/debug_mlp_ak.log 2>&1; tail -12 gpurun_out/debug_mlp_ak.log
timeout 300 python tools/bench_mlp.py > gpurun_out/bench_mlp_ak.log 2>&1; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_mlp_ak.log
timeout 300 python tools/bench_mlp.py 262144 > gpurun_out/bench_mlp_ak_256k.log 2>&1; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_mlp_ak_256k.log
timeout 600 python tools/bench_disney_render.py > gpurun_out/disney_render_ak.log 2>&1; echo "render rc=$?"; cut -c1-500 gpurun_out/disney_render_ak.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_ak.csv python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_disney_ak.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_disney_ak.csv
