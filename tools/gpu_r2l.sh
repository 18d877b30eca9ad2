#!/bin/bash
# round 2: validation + evidence run (fused backward as default)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --flow iid --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_bench_iid.json 2>> gpurun_out/r2_bench.err
timeout 600 python bench.py --layout nchw --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_bench_nchw.json 2>> gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_fwd_strip -s 2 -c 1 -o gpurun_out/r2_k_local_attn_fwd_strip python tools/run_fwd.py --B 16 --iters 2 > gpurun_out/r2_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_bwd_fused -s 2 -c 1 -o gpurun_out/r2_k_local_attn_bwd_fused python tools/run_fwd.py --B 16 --bwd --iters 2 >> gpurun_out/r2_ncu.log 2>&1
timeout 900 python bench.py --workload cfg4 --steps 10 > gpurun_out/r2_cfg4_n1.json 2> gpurun_out/r2_cfg4.err; echo "cfg4 rc=$?"
timeout 900 python bench.py --workload cfg4 --model-dtype fp32 --steps 10 > gpurun_out/r2_cfg4_fp32_n1.json 2>> gpurun_out/r2_cfg4.err
timeout 900 python bench.py --workload cfg5 --steps 5 > gpurun_out/r2_cfg5_n1.json 2> gpurun_out/r2_cfg5.err; echo "cfg5 rc=$?"
timeout 900 python bench.py --workload cfg5 --model-dtype fp32 --steps 5 > gpurun_out/r2_cfg5_fp32_n1.json 2>> gpurun_out/r2_cfg5.err
timeout 300 python tools/ablate_bwd.py > gpurun_out/r2_bwd_ablation.txt 2>&1
timeout 300 python tools/ablate_fwd.py > gpurun_out/r2_fwd_ablation.txt 2>&1
timeout 2400 bash tools/sanitize.sh > gpurun_out/r2_sanitizer.log 2>&1; tail -20 gpurun_out/r2_sanitizer.log
python -c "
import json
j=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value', j['value'], 'ms', j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'], 'step frac', j['step_roofline_frac'])
print('nchw', j['planar_nchw']); print('e2e', j['e2e']); print('launches', j['gpu_launches']); print('iid', j.get('iid_flow')); print('cfg3', j.get('cfg3')); print('refcuda', j.get('reference_cuda')); print('cpu', j.get('cpu_baseline'))"
tail -n 4 gpurun_out/r2_bench.err gpurun_out/r2_cfg4.err gpurun_out/r2_cfg5.err
