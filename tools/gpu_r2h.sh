#!/bin/bash
# round 2, 8th GPU visit: conflict-free slab zeroing, fused backward with G in TMEM + 1-row Q stages
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "tile or strip or blend or cfg2 or channels_last or nchw or golden or bwd or backward" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
GFLA_BWD_FUSED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "bwd or cfg2 or backward or channels_last" > gpurun_out/r2h_pytest_bwd.log 2>&1; echo "pytest fused bwd rc=$?"; tail -3 gpurun_out/r2h_pytest_bwd.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
GFLA_BWD_FUSED=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_fused.json 2>> gpurun_out/r2h_bench.err
GFLA_BWD_FUSED=1 GFLA_BWD_KNOBS=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_fused_pf.json 2>> gpurun_out/r2h_bench.err
GFLA_BWD_FUSED=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras --flow iid > gpurun_out/r2h_bench_fused_iid.json 2>> gpurun_out/r2h_bench.err
for f in gpurun_out/r2h_bench.json gpurun_out/r2h_bench_fused.json gpurun_out/r2h_bench_fused_pf.json gpurun_out/r2h_bench_fused_iid.json; do echo "== $f"; python -c "
import json,sys
j=json.load(open('$f')); print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'], 'nchw', j['planar_nchw']['value'] if j.get('planar_nchw') else None, j['planar_nchw']['ms_per_step'] if j.get('planar_nchw') else None, j['gpu_launches'])"; done
GFLA_BWD_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_bwd_fused -s 2 -c 1 -o gpurun_out/r2h_bwd_fused python tools/run_fwd.py --B 16 --bwd --iters 2 > gpurun_out/r2h_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_fwd_strip -s 2 -c 1 -o gpurun_out/r2h_fwd_strip python tools/run_fwd.py --B 16 --iters 2 >> gpurun_out/r2h_ncu.log 2>&1
GFLA_BUILD_PROFILE=1 timeout 900 python -c "
import importlib.util, os
spec = importlib.util.spec_from_file_location('b', 'global-flow-local-attention_b200/build.py'); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(force=True)" > gpurun_out/r2h_profbuild.log 2>&1; echo "profile build rc=$?"
GFLA_BWD_FUSED=1 timeout 300 python tools/wait_profile.py --which 2 > gpurun_out/r2h_wait_fused.txt 2>&1; cat gpurun_out/r2h_wait_fused.txt
tail -n 5 gpurun_out/r2h_bench.err
