#!/bin/bash
# k_trace_fast v2 (index-based positions, deferred box/empty check, two-level CDF guide, z-pair layered textures): tests + knob sweep
mkdir -p gpurun_out; rm -f gpurun_out/sweep_m.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_m.log
S="python tools/sweep.py --out gpurun_out/sweep_m.jsonl --spp 32 --reps 2 --set staging_subframes=32"
timeout 900 $S --set tex_layout=0,1 --set march_unroll=1,2 > gpurun_out/sweep_m.log 2>&1
timeout 900 $S --set tex_layout=1 --set march_unroll=2 --set block_threads=512,640 >> gpurun_out/sweep_m.log 2>&1
timeout 900 $S --set tex_layout=1 --set march_unroll=2 --set march_keep32=8,16 --set skip_open_dist=1,2 >> gpurun_out/sweep_m.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_m.jsonl'):
    r=json.loads(l); print({k:v for k,v in r['opts'].items() if k!='staging_subframes'}, 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gtaps/s %.1f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'mean %.5f'%r['mean'], 'ev/p %.3f st/p %.2f'%(r['events_per_path'],r['steps_per_path']), 'nonfinite', r['nonfinite'])
PY
tail -3 gpurun_out/sweep_m.log | cut -c1-300
