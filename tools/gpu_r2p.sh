#!/bin/bash
# round 2: cooperative launch of the fused backward, channel-sliced resample->cosine kernels: tests, sanitizer, full bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
timeout 600 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.load(open('gpurun_out/r2p_bench.json'))
print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'])
print('e2e', j['e2e'], 'launches', j['gpu_launches'])
print('f4', json.dumps(j.get('f4_resample_cosine')))
print('nchw', j.get('planar_nchw'), 'iid', j.get('iid_flow'))
PY
SEL="bwd_tile_vs_oracle and shape0 and smooth or many_samples_few_groups or cosine_vs_oracle and float32"
for tool in memcheck racecheck; do
  echo "== $tool =="
  timeout 600 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resample_cosine.py -m gpu -x -q -k "$SEL" 2>&1 | tail -5
done > gpurun_out/r2p_sanitizer.log 2>&1
cat gpurun_out/r2p_sanitizer.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r2p_gpu.txt
tail -n 5 gpurun_out/r2p_bench.err
