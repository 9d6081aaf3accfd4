#!/bin/bash
# launch bounds (1024,1): one block of 1024 threads per SM vs two of 512; device-resident radiance collector: test + C3 comparison
mkdir -p gpurun_out; rm -f gpurun_out/sweep_s.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread --durations=4 > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_s.log
S="python tools/sweep.py --out gpurun_out/sweep_s.jsonl --spp 32 --reps 2 --set staging_subframes=32"
timeout 600 $S --set block_threads=512 --set blocks_per_sm=2 > gpurun_out/sweep_s.log 2>&1
timeout 600 $S --set block_threads=1024,896,768 --set blocks_per_sm=1 >> gpurun_out/sweep_s.log 2>&1
timeout 600 $S --set block_threads=1024 --set blocks_per_sm=1 --set march_keep32=12,16 --set regen_min=2,4 >> gpurun_out/sweep_s.log 2>&1
timeout 600 $S --set block_threads=320,256 --set blocks_per_sm=3,4 >> gpurun_out/sweep_s.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_s.jsonl'):
    r=json.loads(l); print({k:v for k,v in r['opts'].items() if k!='staging_subframes'}, 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gtaps/s %.1f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'mean %.5f'%r['mean'], 'nonfinite', r['nonfinite'])
PY
timeout 600 python tools/bench_dataset.py --scenes 2 > gpurun_out/c3_s1.log 2>&1; echo "c3 adaptive rc=$?"; tail -c 1200 gpurun_out/c3_s1.log
timeout 600 python tools/bench_dataset.py --scenes 2 --opt radiance_quota=256 > gpurun_out/c3_s2.log 2>&1; echo "c3 adaptive q256 rc=$?"; tail -c 1200 gpurun_out/c3_s2.log
