#!/bin/bash
# first GPU session: parity tests, smoke, bench, knob sweep, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 --spp 2 > gpurun_out/bench_a.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_a.log
timeout 600 python tools/sweep.py --out gpurun_out/sweep_a.jsonl --set precision=1 --set skip_empty=0,1 --set march_keep_quarters=0,1,2,3 > gpurun_out/sweep_a.log 2>&1; echo "sweep rc=$?"
timeout 300 python tools/sweep.py --out gpurun_out/sweep_a.jsonl --set precision=0 --set skip_empty=1 --set march_keep_quarters=2 >> gpurun_out/sweep_a.log 2>&1
timeout 300 python tools/sweep.py --out gpurun_out/sweep_a.jsonl --set precision=1 --set block_threads=128,256,512 --set blocks_per_sm=2,4,8 >> gpurun_out/sweep_a.log 2>&1
cat gpurun_out/sweep_a.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 1 --spp 2 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_r1a python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
