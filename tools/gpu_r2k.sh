#!/bin/bash
mkdir -p gpurun_out
GFLA_BWD_FUSED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "bwd or cfg2 or backward or channels_last" > gpurun_out/r2k_pytest_bwd.log 2>&1; echo "pytest fused bwd rc=$?"; tail -3 gpurun_out/r2k_pytest_bwd.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile or strip or blend" > gpurun_out/r2k_pytest_fwd.log 2>&1; echo "pytest fwd rc=$?"; tail -3 gpurun_out/r2k_pytest_fwd.log
timeout 300 python tools/ablate_bwd.py 2>&1 | tee gpurun_out/r2_bwd_ablation.txt
GFLA_BWD_FUSED=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2k_bench_fused.json 2>> gpurun_out/r2k_bench.err
GFLA_BWD_FUSED=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras --flow iid > gpurun_out/r2k_bench_fused_iid.json 2>> gpurun_out/r2k_bench.err
for f in gpurun_out/r2k_bench_fused.json gpurun_out/r2k_bench_fused_iid.json; do echo "== $f"; python -c "
import json,sys
j=json.load(open('$f')); print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'], 'nchw', j['planar_nchw']['value'] if j.get('planar_nchw') else None, j['planar_nchw']['ms_per_step'] if j.get('planar_nchw') else None, j['gpu_launches'])"; done
tail -n 5 gpurun_out/r2k_bench.err
