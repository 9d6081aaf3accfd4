#!/bin/bash
# second GPU session: optimised FAST kernel (k_trace_fast) -- parity, knob sweep, bench, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_b.log
rm -f gpurun_out/sweep_b.jsonl
timeout 600 python tools/sweep.py --out gpurun_out/sweep_b.jsonl --set variant=1,0 --set skip_empty=0,1 > gpurun_out/sweep_b.log 2>&1; echo "sweep1 rc=$?"
timeout 600 python tools/sweep.py --out gpurun_out/sweep_b.jsonl --set march_keep32=0,2,4,6,8,12,16,24 >> gpurun_out/sweep_b.log 2>&1; echo "sweep2 rc=$?"
timeout 600 python tools/sweep.py --out gpurun_out/sweep_b.jsonl --set block_threads=640 --set blocks_per_sm=2 >> gpurun_out/sweep_b.log 2>&1
timeout 600 python tools/sweep.py --out gpurun_out/sweep_b.jsonl --set block_threads=416 --set blocks_per_sm=3 >> gpurun_out/sweep_b.log 2>&1
timeout 600 python tools/sweep.py --out gpurun_out/sweep_b.jsonl --set block_threads=320,256 --set blocks_per_sm=3 >> gpurun_out/sweep_b.log 2>&1
timeout 600 python tools/sweep.py --out gpurun_out/sweep_b.jsonl --set block_threads=512 --set blocks_per_sm=2 --set march_max_iters=8,16,32 >> gpurun_out/sweep_b.log 2>&1
cat gpurun_out/sweep_b.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['opts'], 'Mpaths/s %.1f'%r['mpaths_s'], 'Gev/s %.2f'%r['gevents_s'], 'Gsteps/s %.1f'%r['gsteps_s'], 'Gtaps/s %.2f'%r['gtaps_s'], 'ms %.2f'%r['trace_ms'], 'algGB/s %.0f'%r['alg_gbs'], 'mean %.4f'%r['mean'], 'ev/path %.2f steps/path %.1f'%(r['events_per_path'],r['steps_per_path']))
"
timeout 900 python bench.py --steps 5 --warmup 3 --spp 4 > gpurun_out/bench_b.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_r1b python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1; echo "ncu full rc=$?"
