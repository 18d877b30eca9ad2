#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/mma_bench > gpurun_out/r2_mma_bench.txt 2>&1; echo "mma_bench rc=$?"; cat gpurun_out/r2_mma_bench.txt
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q > gpurun_out/r2g_pytest_models.log 2>&1; echo "pytest models rc=$?"; tail -4 gpurun_out/r2g_pytest_models.log
