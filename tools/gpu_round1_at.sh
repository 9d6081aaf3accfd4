#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_descriptors -s 1 -c 1 -f -o gpurun_out/prof_desc_at python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_desc_at.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_desc_at.ncu-rep --page details > gpurun_out/k_descriptors_ncu_at.txt 2>/dev/null; wc -l gpurun_out/k_descriptors_ncu_at.txt
