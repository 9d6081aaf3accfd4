#!/bin/bash
# descriptor gather on the texture units (FAST neural renderer): parity + frame time + launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_disney_mlp.py -k "descriptor or network_input or disney" -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/pytest_ao.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_ao.log
timeout 600 python tools/bench_disney_render.py > gpurun_out/disney_render_ao.log 2>&1; echo "render rc=$?"; cut -c1-500 gpurun_out/disney_render_ao.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_ao.csv python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_disney_ao.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_disney_ao.csv | head -6
