#!/bin/bash
# k_trace_fast v2 with the defaults picked by sweep m: tests, bench, ncu launch list, ncu full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_n.log
timeout 900 python bench.py > gpurun_out/bench_n.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_n.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 1 -c 1 -o gpurun_out/prof_trace_r1n python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_n.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
