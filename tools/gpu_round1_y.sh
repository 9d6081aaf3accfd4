#!/bin/bash
# C4 (1024^3): ncu --set full of k_trace_fast with the auto-selected single-tap march: is it DRAM-bound?
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 1 -c 1 -o gpurun_out/prof_trace_r1y_C4 python tools/sweep.py --tag C4 --grid 1024 --size 12000 --sun=0.995,-0.0998,0 --spp 16 --reps 1 --set staging_subframes=16 > gpurun_out/ncu_full_y.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_y.log | cut -c1-300
