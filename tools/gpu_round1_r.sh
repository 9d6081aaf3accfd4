#!/bin/bash
# v3 profile evidence: ncu launch list of the default bench command, one --set full capture of k_trace_fast; test durations
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread --durations=6 > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_r.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1r.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 1 -c 1 -o gpurun_out/prof_trace_r1r python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -5
