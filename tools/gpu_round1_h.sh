#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep_h.jsonl gpurun_out/sweep_h.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_h.log
S="python tools/sweep.py --out gpurun_out/sweep_h.jsonl --spp 32 --reps 1 --set staging_subframes=64"
timeout 600 $S >> gpurun_out/sweep_h.log 2>&1
timeout 600 $S --set block_threads=512,640 >> gpurun_out/sweep_h.log 2>&1
timeout 600 $S --set skip_open_dist=1,3,4 >> gpurun_out/sweep_h.log 2>&1
timeout 600 $S --set skip_min=2,8 --set regen_min=2,8 >> gpurun_out/sweep_h.log 2>&1
timeout 600 $S --set march_keep32=6,10,12 >> gpurun_out/sweep_h.log 2>&1
cat gpurun_out/sweep_h.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['opts'], 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gsteps/s %.1f'%r['gsteps_s'], 'Gtaps/s %.2f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'algGB/s %.0f'%r['alg_gbs'], 'mean %.4f'%r['mean'], 'ev/p %.2f st/p %.1f'%(r['events_per_path'],r['steps_per_path']))
"
tail -2 gpurun_out/sweep_h.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_r1h python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline > gpurun_out/ncu_full_h.log 2>&1; echo "ncu full rc=$?"
