#!/bin/bash
# ncu full captures of k_trace_fast v2, both texture layouts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_n.log
for L in 0 1; do
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 1 -c 1 -o gpurun_out/prof_trace_r1n_layout$L python bench.py --steps 1 --warmup 1 --no-cpu-baseline --opt tex_layout=$L --opt march_unroll=2 > gpurun_out/ncu_full_n$L.log 2>&1; echo "ncu full rc=$?"
done
ls -la gpurun_out
