#!/bin/bash
# bench default = 64 subframes per step (drain tail amortised): bench line (with cpu baseline), reference arm, ncu launch list + full capture of the same command
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_ah.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_ah.log | cut -c1-220
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_ah.log 2>&1; echo "bench ref rc=$?"; grep '^{' gpurun_out/bench_ref_ah.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1ah.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_ah.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 1 -c 1 -o gpurun_out/prof_trace_r1ah python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ah.log 2>&1; echo "ncu full rc=$?"
