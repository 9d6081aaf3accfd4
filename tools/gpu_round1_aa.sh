#!/bin/bash
# network-input pass (f4, first half): gpu tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_aa.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_aa.log
