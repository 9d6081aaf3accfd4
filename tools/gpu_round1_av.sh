#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method thread -k "descriptor or disney or network or datagen" > gpurun_out/pytest_gpu_av.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_av.log
timeout 600 python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/disney_render_av.log 2>&1; echo "render rc=$?"; cut -c1-300 gpurun_out/disney_render_av.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_av.csv python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_disney_av.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_disney_av.csv 2>/dev/null | head -5
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_av.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_av.log | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['value'], r.get('secondary'))"
