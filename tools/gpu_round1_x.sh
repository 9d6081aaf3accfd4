#!/bin/bash
# collector test again; C4 (1024^3, HBM-bound) knob sweep: speculative pairs vs single taps, block shapes
mkdir -p gpurun_out; rm -f gpurun_out/sweep_x.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --timeout-method thread -k "collector" > gpurun_out/pytest_gpu_x.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_x.log
S="python tools/sweep.py --out gpurun_out/sweep_x.jsonl --tag C4 --grid 1024 --size 12000 --sun=0.995,-0.0998,0 --spp 16 --reps 1 --set staging_subframes=16"
timeout 900 $S --set march_unroll=2,1 --set block_threads=896,512 --set blocks_per_sm=1,2 > gpurun_out/sweep_x.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_x.jsonl'):
    r=json.loads(l); print({k:v for k,v in r['opts'].items() if k!='staging_subframes'}, 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gtaps/s %.1f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'mean %.5f'%r['mean'], 'nonfinite', r['nonfinite'])
PY
