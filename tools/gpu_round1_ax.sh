#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu_ax.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_ax.log
timeout 600 python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/disney_render_ax.log 2>&1; echo "render rc=$?"; cut -c1-300 gpurun_out/disney_render_ax.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_ax.csv python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_disney_ax.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_disney_ax.csv 2>/dev/null | head -6
