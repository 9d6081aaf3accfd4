#!/bin/bash
# single GPU: full gpu test suite after the collector cadence fix; C3 with the device-resident collector; C4 / C5 / C1 configs with the current kernel
mkdir -p gpurun_out; rm -f gpurun_out/configs_v.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread --durations=3 > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_v.log
timeout 600 python tools/bench_dataset.py --scenes 2 > gpurun_out/c3_v.log 2>&1; echo "c3 rc=$?"; tail -c 1200 gpurun_out/c3_v.log
S="python tools/sweep.py --out gpurun_out/configs_v.jsonl"
timeout 600 $S --tag C1 --grid 256 --kind 1 --width 256 --height 256 --spp 64 --reps 1 --mode 2 --sun=-0.03,-0.25,0.8 --set staging_subframes=64 > gpurun_out/configs_v.log 2>&1
timeout 600 $S --tag C2 --spp 32 --reps 2 --set staging_subframes=32 >> gpurun_out/configs_v.log 2>&1
timeout 900 $S --tag C4 --grid 1024 --size 12000 --sun=0.995,-0.0998,0 --spp 16 --reps 1 --set staging_subframes=16 >> gpurun_out/configs_v.log 2>&1
timeout 900 $S --tag C5 --width 3840 --height 2160 --spp 8 --reps 1 --set staging_subframes=8 >> gpurun_out/configs_v.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/configs_v.jsonl'):
    r=json.loads(l); print(r['tag'], 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gsteps/s %.1f'%r['gsteps_s'], 'ms %.2f'%r['trace_ms'], 'algGB/s %.0f'%r['alg_gbs'], 'mean %.4f'%r['mean'], 'ev/p %.2f st/p %.1f'%(r['events_per_path'],r['steps_per_path']), 'bake %.3f'%r['bake_s'], 'nonfinite', r['nonfinite'])
PY
