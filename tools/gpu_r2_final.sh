#!/bin/bash
# gpurun_ab/*.so are variant builds made on the build host before the call (git worktree of the variant -> build.py -> copy;
# GFLA_BUILD_PROFILE=1 / GFLA_BUILD_KNOBS=1 for the profile / tuning builds); the directory is git-ignored (*.so) and not kept.
# round 2, final visit with the shipped build: full GPU suite, sanitizer, the bench line, ncu captures + launch list, wait profile
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,driver_version --format=csv > gpurun_out/r2f_gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.load(open('gpurun_out/r2f_bench.json'))
print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'])
print('e2e', j['e2e'], 'launches', j['gpu_launches'], 'cpu', j.get('cpu_baseline', {}).get('value'))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench.err; echo "reference arm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_bwd_fused -s 2 -c 1 -o gpurun_out/r2f_bwd_fused python tools/run_fwd.py --B 16 --bwd --iters 2 > gpurun_out/r2f_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_fwd_strip -s 2 -c 1 -o gpurun_out/r2f_fwd_strip python tools/run_fwd.py --B 16 --iters 4 >> gpurun_out/r2f_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2f_launches_bench.log 2>&1
SEL="tile_vs_oracle and shape0 and smooth or bwd_tile_vs_oracle and shape0 and smooth or many_samples_few_groups or golden and k3 or blend_fwd and nhwc or cosine_vs_oracle and float32"
for tool in memcheck racecheck; do
  echo "== $tool =="
  timeout 600 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resample_cosine.py -m gpu -x -q -k "$SEL" 2>&1 | tail -5
done > gpurun_out/r2f_sanitizer.log 2>&1
cat gpurun_out/r2f_sanitizer.log
GFLA_LIB=$PWD/gpurun_ab/libgfla_profile.so timeout 300 python tools/wait_profile.py --which 2 > gpurun_out/r2f_wait_fused.txt 2>&1; cat gpurun_out/r2f_wait_fused.txt
GFLA_LIB=$PWD/gpurun_ab/libgfla_profile.so timeout 300 python tools/wait_profile.py --which 1 > gpurun_out/r2f_wait_strip.txt 2>&1; cat gpurun_out/r2f_wait_strip.txt
tail -n 5 gpurun_out/r2f_bench.err
