#!/bin/bash
# pipelined march taps (prefetch on state entry) + one block of 896 threads per SM: tests + sweep
mkdir -p gpurun_out; rm -f gpurun_out/sweep_t.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread --durations=4 > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_t.log
S="python tools/sweep.py --out gpurun_out/sweep_t.jsonl --spp 32 --reps 2 --set staging_subframes=32"
timeout 600 $S --set block_threads=896,1024 --set march_unroll=2,1 > gpurun_out/sweep_t.log 2>&1
timeout 600 $S --set block_threads=896 --set march_keep32=10,12,16,20 >> gpurun_out/sweep_t.log 2>&1
timeout 600 $S --set block_threads=896 --set march_max_iters=16,32 --set regen_min=1,4 >> gpurun_out/sweep_t.log 2>&1
timeout 600 $S --set block_threads=832,960 --set zero_check_min=1,4 >> gpurun_out/sweep_t.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_t.jsonl'):
    r=json.loads(l); print({k:v for k,v in r['opts'].items() if k!='staging_subframes'}, 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gtaps/s %.1f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'mean %.5f'%r['mean'], 'nonfinite', r['nonfinite'])
PY
