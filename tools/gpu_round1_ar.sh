#!/bin/bash
# final-state artefacts of this session: bench line (with cpu baseline), reference arm, ncu launch list + full capture of the default bench step
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_ar.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_ar.log | cut -c1-260
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_ar.log 2>&1; echo "bench ref rc=$?"; grep '^{' gpurun_out/bench_ref_ar.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1ar.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_ar.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r1ar.csv 2>/dev/null | head -6
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 1 -c 1 -f -o gpurun_out/prof_trace_r1ar python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ar.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_trace_r1ar.ncu-rep --page details > gpurun_out/k_trace_fast_ncu_ar.txt 2>/dev/null; wc -l gpurun_out/k_trace_fast_ncu_ar.txt
timeout 600 ncu --set full --clock-control none -k regex:k_disney_mlp_tc -s 1 -c 1 -f -o gpurun_out/prof_mlp_r1ar python tools/bench_mlp.py 262144 > gpurun_out/ncu_mlp_ar.log 2>&1; echo "ncu mlp rc=$?"
ncu -i gpurun_out/prof_mlp_r1ar.ncu-rep --page details > gpurun_out/k_disney_mlp_tc_ncu_ar.txt 2>/dev/null; wc -l gpurun_out/k_disney_mlp_tc_ncu_ar.txt
