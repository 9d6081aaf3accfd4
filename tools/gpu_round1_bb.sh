#!/bin/bash
# full suite after the operand-type templating; neural renderer with tf32 / bf16 operands; launch list of the bf16 frame
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_bb.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_bb.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu_bb.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_bb.log
timeout 600 python tools/bench_disney_render.py > gpurun_out/disney_render_bb.log 2>&1; echo "render rc=$?"; cut -c1-330 gpurun_out/disney_render_bb.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_bb.csv python tools/bench_disney_render.py 1920 1080 512 fast_bf16 > gpurun_out/ncu_disney_bb.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_disney_bb.csv 2>/dev/null | head -6
