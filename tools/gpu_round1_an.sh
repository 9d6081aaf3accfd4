#!/bin/bash
# full GPU suite; neural renderer at 1080p (frame-wide pipeline, reworked model kernel, faded-tap skip in the descriptor gather) + launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu_an.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_an.log
timeout 600 python tools/bench_disney_render.py > gpurun_out/disney_render_an.log 2>&1; echo "render rc=$?"; cut -c1-500 gpurun_out/disney_render_an.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_an.csv python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_disney_an.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_disney_an.csv
timeout 300 python tools/bench_mlp.py 262144 > gpurun_out/bench_mlp_an_256k.log 2>&1; cut -c1-400 gpurun_out/bench_mlp_an_256k.log
