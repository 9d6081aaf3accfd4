#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -k "edge_cases or empty_cloud" -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_aq.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_aq.log | cut -c1-220
