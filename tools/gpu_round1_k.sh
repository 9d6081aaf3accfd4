#!/bin/bash
# other BASELINE configs: C1 single scatter cube, C4 thick cumulus 1024^3 grazing sun, C5 4K frame
mkdir -p gpurun_out; rm -f gpurun_out/configs_k.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_k.log
S="python tools/sweep.py --out gpurun_out/configs_k.jsonl"
timeout 600 $S --tag C1 --grid 256 --kind 1 --width 256 --height 256 --spp 64 --reps 1 --mode 2 --sun=-0.03,-0.25,0.8 --set staging_subframes=64 > gpurun_out/configs_k.log 2>&1
timeout 600 $S --tag C2 --spp 32 --reps 2 --set staging_subframes=32 >> gpurun_out/configs_k.log 2>&1
timeout 900 $S --tag C4 --grid 1024 --size 12000 --sun=0.995,-0.0998,0 --spp 16 --reps 1 --set staging_subframes=16 >> gpurun_out/configs_k.log 2>&1
timeout 900 $S --tag C5 --width 3840 --height 2160 --spp 8 --reps 1 --set staging_subframes=8 >> gpurun_out/configs_k.log 2>&1
timeout 600 $S --tag C2-exact --spp 8 --reps 1 --set precision=0 --set staging_subframes=8 >> gpurun_out/configs_k.log 2>&1
cat gpurun_out/configs_k.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['tag'], r['opts'], 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gsteps/s %.1f'%r['gsteps_s'], 'ms %.2f'%r['trace_ms'], 'algGB/s %.0f'%r['alg_gbs'], 'mean %.4f'%r['mean'], 'ev/p %.2f st/p %.1f'%(r['events_per_path'],r['steps_per_path']), 'bake %.3f'%r['bake_s'], 'nonfinite', r['nonfinite'])
"
tail -3 gpurun_out/configs_k.log | cut -c1-200
