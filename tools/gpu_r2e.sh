#!/bin/bash
# round 2, 5th GPU visit: fused backward -- clipped reduce boxes, A-from-TMEM option; strip profile
mkdir -p gpurun_out
GFLA_BWD_FUSED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "bwd or cfg2 or backward or channels_last" > gpurun_out/r2e_pytest_bwd.log 2>&1; echo "pytest fused bwd rc=$?"; tail -4 gpurun_out/r2e_pytest_bwd.log
GFLA_BWD_FUSED=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2e_bench_fused.json 2>> gpurun_out/r2e_bench.err
GFLA_BWD_FUSED=1 GFLA_BWD_QA_TMEM=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "bwd or cfg2 or backward or channels_last" > gpurun_out/r2e_pytest_bwd_ts.log 2>&1; echo "pytest fused bwd TS rc=$?"; tail -4 gpurun_out/r2e_pytest_bwd_ts.log
GFLA_BWD_FUSED=1 GFLA_BWD_QA_TMEM=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2e_bench_fused_ts.json 2>> gpurun_out/r2e_bench.err
for f in gpurun_out/r2e_bench_fused.json gpurun_out/r2e_bench_fused_ts.json; do echo "== $f"; python -c "
import json,sys
j=json.load(open('$f')); print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'], 'nchw', j['planar_nchw']['value'] if j.get('planar_nchw') else None, j['planar_nchw']['ms_per_step'] if j.get('planar_nchw') else None, j['gpu_launches'])"; done
GFLA_BUILD_PROFILE=1 timeout 900 python -c "
import importlib.util, os
spec = importlib.util.spec_from_file_location('b', 'global-flow-local-attention_b200/build.py'); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(force=True)" > gpurun_out/r2e_profbuild.log 2>&1; echo "profile build rc=$?"
timeout 300 python tools/wait_profile.py --which 1 > gpurun_out/r2e_wait_strip.txt 2>&1; cat gpurun_out/r2e_wait_strip.txt
GFLA_BWD_FUSED=1 timeout 300 python tools/wait_profile.py --which 2 > gpurun_out/r2e_wait_fused.txt 2>&1; cat gpurun_out/r2e_wait_fused.txt
GFLA_BWD_FUSED=1 GFLA_BWD_QA_TMEM=1 timeout 300 python tools/wait_profile.py --which 2 > gpurun_out/r2e_wait_fused_ts.txt 2>&1; cat gpurun_out/r2e_wait_fused_ts.txt
tail -n 5 gpurun_out/r2e_bench.err
