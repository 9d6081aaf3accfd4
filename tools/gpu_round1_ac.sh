#!/bin/bash
# branch-free frame for the new direction (FAST flavour), EXR output of the render task: tests + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_ac.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_ac.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_ac.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_ac.log | cut -c1-200
python tools/sweep.py --spp 32 --reps 2 --set staging_subframes=32 --set march_keep32=14,18 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print({k:v for k,v in r['opts'].items() if k!='staging_subframes'}, 'Mpaths/s %.1f'%r['mpaths_s'], 'mean %.5f'%r['mean'])
"
