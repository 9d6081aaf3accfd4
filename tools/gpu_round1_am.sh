#!/bin/bash
# ncu full capture (source counters + stall sampling) of the tensor-core model kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_disney_mlp_tc -s 1 -c 1 -f -o gpurun_out/prof_mlp_am python tools/bench_mlp.py > gpurun_out/ncu_mlp_am.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_mlp_am.log
ls -la gpurun_out/*.ncu-rep
