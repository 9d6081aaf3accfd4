#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep_c.jsonl
for spp in 1 4 16 64; do
  timeout 600 python tools/sweep.py --out gpurun_out/sweep_c.jsonl --spp $spp --reps 1 --set march_keep32=12 --set staging_subframes=64 >> gpurun_out/sweep_c.log 2>&1
done
timeout 600 python tools/sweep.py --out gpurun_out/sweep_c.jsonl --spp 32 --reps 1 --set staging_subframes=64 --set march_keep32=4,8,12,16,24 >> gpurun_out/sweep_c.log 2>&1
timeout 600 python tools/sweep.py --out gpurun_out/sweep_c.jsonl --spp 32 --reps 1 --set staging_subframes=64 --set march_keep32=12 --set block_threads=640 --set blocks_per_sm=2 >> gpurun_out/sweep_c.log 2>&1
timeout 600 python tools/sweep.py --out gpurun_out/sweep_c.jsonl --spp 32 --reps 1 --set staging_subframes=64 --set variant=1 >> gpurun_out/sweep_c.log 2>&1
cat gpurun_out/sweep_c.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['opts'], 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gsteps/s %.1f'%r['gsteps_s'], 'Gtaps/s %.2f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'algGB/s %.0f'%r['alg_gbs'], 'mean %.4f'%r['mean'])
"
tail -3 gpurun_out/sweep_c.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_r1c python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --opt march_keep32=12 > gpurun_out/ncu_full_c.log 2>&1; echo "ncu full rc=$?"
