#!/bin/bash
# tensor-core model kernel v4: descriptor chunks by bulk TMA from MMA-ready tiles written by the descriptor gather
mkdir -p gpurun_out
timeout 300 python tools/debug_mlp.py 300 2>&1 | tail -9
timeout 600 python -m pytest tests/test_disney_mlp.py tests/test_datagen.py -m gpu -q --timeout 240 --timeout-method thread 2>&1 | tail -6
timeout 300 python tools/bench_mlp.py > gpurun_out/bench_mlp_as.log 2>&1; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_mlp_as.log
timeout 300 python tools/bench_mlp.py 262144 > gpurun_out/bench_mlp_as_256k.log 2>&1; cut -c1-300 gpurun_out/bench_mlp_as_256k.log | head -2
timeout 600 python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/disney_render_as.log 2>&1; echo "render rc=$?"; cut -c1-400 gpurun_out/disney_render_as.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_as.csv python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_disney_as.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_disney_as.csv 2>/dev/null | head -5
