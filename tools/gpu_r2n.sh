#!/bin/bash
# round 2: resample2d tile / shared-box kernels + re-validation of the fused backward with the global zero-fill progress word
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_samples_few_groups" > gpurun_out/r2n_pytest_zero.log 2>&1; echo "pytest zero-fill rc=$?"; tail -3 gpurun_out/r2n_pytest_zero.log
timeout 600 python bench.py --workload cfg4 --arms fused,literal --steps 10 --warmup 3 > gpurun_out/r2n_cfg4_n1.json 2> gpurun_out/r2n_cfg4_n1.err; echo "cfg4 rc=$?"; cut -c1-600 gpurun_out/r2n_cfg4_n1.json
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "resample" > gpurun_out/r2n_pytest_rs.log 2>&1; echo "pytest resample rc=$?"; tail -3 gpurun_out/r2n_pytest_rs.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 600 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.load(open('gpurun_out/r2n_bench.json'))
print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], j['roofline_fwd']['frac'], 'bwd', j['roofline_bwd']['launch_ms'], j['roofline_bwd']['frac'])
print('e2e', j['e2e'], 'launches', j['gpu_launches'])
print('cfg3', json.dumps(j.get('cfg3'))[:1500])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample2d -c 3 -o gpurun_out/r2n_resample2d_ks4 python tools/run_resample.py --ks 4 --B 4 > gpurun_out/r2n_ncu_rs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample2d -c 3 -o gpurun_out/r2n_resample2d_ks2 python tools/run_resample.py --ks 2 --sigma 5 --B 4 >> gpurun_out/r2n_ncu_rs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local_attn_bwd_fused -s 2 -c 1 -o gpurun_out/r2n_bwd_fused python tools/run_fwd.py --B 16 --bwd --iters 2 > gpurun_out/r2n_ncu.log 2>&1
SEL="bwd_tile_vs_oracle and shape0 and smooth or shared_box_scatter and smooth"
for tool in memcheck racecheck; do
  echo "== $tool =="
  timeout 600 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -5
done > gpurun_out/r2n_sanitizer.log 2>&1
cat gpurun_out/r2n_sanitizer.log
tail -n 5 gpurun_out/r2n_bench.err
