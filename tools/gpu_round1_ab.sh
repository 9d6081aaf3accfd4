#!/bin/bash
# 2 GPUs: smoke(), bench at N=2, C3 with the dynamic scene queue
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ab.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_ab.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2.log | cut -c1-260
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/bench_dataset.py --scenes 5 > gpurun_out/c3_n2_dynamic.log 2>&1; echo "c3 n2 rc=$?"; grep '^{' gpurun_out/c3_n2_dynamic.log | cut -c1-1500
