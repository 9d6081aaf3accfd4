#!/bin/bash
# end-of-session verification: smoke, full GPU suite, default bench (with secondary numbers), ncu capture of the model kernel
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_az.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_az.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread --durations=5 > gpurun_out/pytest_gpu_az.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/pytest_gpu_az.log
timeout 900 python bench.py > gpurun_out/bench_az.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_az.log | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['value'], r['e2e']['value'], r['roofline']['frac'], r['cpu_baseline']['value'], r.get('secondary'))"
timeout 600 ncu --set full --clock-control none -k regex:k_disney_mlp_tc -s 1 -c 1 -f -o gpurun_out/prof_mlp_r1az python tools/bench_mlp.py 262144 > gpurun_out/ncu_mlp_az.log 2>&1; echo "ncu mlp rc=$?"
ncu -i gpurun_out/prof_mlp_r1az.ncu-rep --page details > gpurun_out/k_disney_mlp_tc_ncu_az.txt 2>/dev/null; grep -E "Duration|highest-utilized|Issue Slots Busy" gpurun_out/k_disney_mlp_tc_ncu_az.txt | head -4
