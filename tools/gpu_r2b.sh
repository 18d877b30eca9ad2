#!/bin/bash
# round 2, second GPU visit: fused backward kernel -- correctness first, then timing and ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "bwd or cfg2 or backward or channels_last" > gpurun_out/r2b_pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"; tail -25 gpurun_out/r2b_pytest_bwd.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_fused.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
GFLA_BWD_FUSED=0 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_twokernel.json 2>> gpurun_out/r2b_bench.err
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras --flow iid > gpurun_out/r2b_bench_fused_iid.json 2>> gpurun_out/r2b_bench.err
timeout 1200 python -m pytest tests -m gpu -q -rf > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2b_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_bwd_fused -s 2 -c 1 -o gpurun_out/r2b_bwd_fused python tools/run_fwd.py --B 16 --bwd --iters 2 > gpurun_out/r2b_ncu.log 2>&1
timeout 600 python bench.py --workload cfg4 --steps 5 > gpurun_out/r2b_cfg4.json 2> gpurun_out/r2b_cfg4.err; echo "cfg4 rc=$?"
timeout 600 python bench.py --workload cfg5 --steps 3 > gpurun_out/r2b_cfg5.json 2> gpurun_out/r2b_cfg5.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg5 --model-dtype fp32 --steps 3 > gpurun_out/r2b_cfg5_fp32.json 2>> gpurun_out/r2b_cfg5.err
for f in gpurun_out/r2b_bench_fused.json gpurun_out/r2b_bench_twokernel.json gpurun_out/r2b_bench_fused_iid.json; do echo "== $f"; python -c "
import json,sys
j=json.load(open('$f')); print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'], 'nchw', j['planar_nchw']['value'] if j.get('planar_nchw') else None, j['gpu_launches'])"; done
for f in gpurun_out/r2b_cfg4.json gpurun_out/r2b_cfg5.json gpurun_out/r2b_cfg5_fp32.json; do echo "== $f"; head -c 2500 $f; echo; done
tail -n 5 gpurun_out/r2b_bench.err gpurun_out/r2b_cfg4.err gpurun_out/r2b_cfg5.err
