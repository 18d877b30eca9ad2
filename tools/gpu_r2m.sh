#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py tests/test_gpu_models.py -m gpu -x -q -k "bwd or cfg2 or backward or channels_last or generator or nchw" > gpurun_out/r2m_pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"; tail -3 gpurun_out/r2m_pytest_bwd.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2m_bench.json 2>> gpurun_out/r2m_bench.err
GFLA_BWD_ZERO_IN_KERNEL=0 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2m_bench_memset.json 2>> gpurun_out/r2m_bench.err
for f in gpurun_out/r2m_bench.json gpurun_out/r2m_bench_memset.json; do echo "== $f"; python -c "
import json,sys
j=json.load(open('$f')); print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'], 'nchw', j['planar_nchw']['value'] if j.get('planar_nchw') else None, j['planar_nchw']['ms_per_step'] if j.get('planar_nchw') else None, j['gpu_launches'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_bwd_fused -s 2 -c 1 -o gpurun_out/r2m_bwd_fused python tools/run_fwd.py --B 16 --bwd --iters 2 > gpurun_out/r2m_ncu.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bwd_tile_vs_oracle and shape0 and smooth" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bwd_tile_vs_oracle and shape0 and smooth" 2>&1 | tail -4
tail -n 5 gpurun_out/r2m_bench.err
