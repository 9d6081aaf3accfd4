#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep_i.jsonl gpurun_out/sweep_i.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_i.log
S="python tools/sweep.py --out gpurun_out/sweep_i.jsonl --spp 32 --reps 1 --set staging_subframes=64"
timeout 900 $S --set guide_n=4096,16384 --set block_threads=576,640 --set regen_min=2,8 --set skip_min=4,8 >> gpurun_out/sweep_i.log 2>&1
timeout 900 $S --set march_keep32=8,12 --set skip_max_iters=4,8,32 >> gpurun_out/sweep_i.log 2>&1
cat gpurun_out/sweep_i.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['opts'], 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'ms %.2f'%r['trace_ms'], 'mean %.4f'%r['mean'], 'st/p %.1f'%(r['steps_per_path']))
"
