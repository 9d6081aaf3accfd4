#!/bin/bash
# quad texture layout: tests, then C4 (1024^3) and C2 with both layouts and both march variants
mkdir -p gpurun_out; rm -f gpurun_out/sweep_z.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_z.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_z.log
S="python tools/sweep.py --out gpurun_out/sweep_z.jsonl"
timeout 900 $S --tag C4 --grid 1024 --size 12000 --sun=0.995,-0.0998,0 --spp 16 --reps 1 --set staging_subframes=16 --set tex_layout=0,1 --set march_unroll=1,2 > gpurun_out/sweep_z.log 2>&1
timeout 900 $S --tag C2 --spp 32 --reps 2 --set staging_subframes=32 --set tex_layout=0,1 --set march_unroll=2,1 >> gpurun_out/sweep_z.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_z.jsonl'):
    r=json.loads(l); print(r['tag'], {k:v for k,v in r['opts'].items() if k!='staging_subframes'}, 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gtaps/s %.1f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'mean %.5f'%r['mean'], 'nonfinite', r['nonfinite'])
PY
tail -2 gpurun_out/sweep_z.log | cut -c1-300
