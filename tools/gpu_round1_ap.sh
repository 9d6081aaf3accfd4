#!/bin/bash
# smoke + full GPU suite with the neural renderer in the C++ driver
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_ap.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_ap.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu_ap.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_ap.log
