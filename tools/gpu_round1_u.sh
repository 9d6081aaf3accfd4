#!/bin/bash
# current defaults (896 threads x 1 block/SM, pipelined tap pairs, device-resident collector): tests, bench line, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_u.log
timeout 900 python bench.py > gpurun_out/bench_u.log 2>&1; echo "bench rc=$?"; tail -c 2800 gpurun_out/bench_u.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1u.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_u.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 1 -c 1 -o gpurun_out/prof_trace_r1u python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_u.log 2>&1; echo "ncu full rc=$?"
