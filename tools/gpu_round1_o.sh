#!/bin/bash
# k_trace_fast v3 (speculative tap pairs, deferred NEE accumulation, host-computed constants, estimator as template parameter, --use_fast_math): tests + knob sweep
mkdir -p gpurun_out; rm -f gpurun_out/sweep_o.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_o.log
S="python tools/sweep.py --out gpurun_out/sweep_o.jsonl --spp 32 --reps 2 --set staging_subframes=32"
timeout 900 $S --set march_unroll=1,2 --set zero_check_min=1,3 > gpurun_out/sweep_o.log 2>&1
timeout 900 $S --set march_unroll=2 --set zero_check_min=2,6 --set march_keep32=12,16 >> gpurun_out/sweep_o.log 2>&1
timeout 900 $S --set march_unroll=2 --set march_keep32=8,10,14 --set march_max_iters=64 >> gpurun_out/sweep_o.log 2>&1
timeout 900 $S --set march_unroll=2 --set block_threads=512,576 --set regen_min=2,4 >> gpurun_out/sweep_o.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_o.jsonl'):
    r=json.loads(l); print({k:v for k,v in r['opts'].items() if k!='staging_subframes'}, 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gtaps/s %.1f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'mean %.5f'%r['mean'], 'ev/p %.3f st/p %.2f'%(r['events_per_path'],r['steps_per_path']), 'nonfinite', r['nonfinite'])
PY
tail -3 gpurun_out/sweep_o.log | cut -c1-300
