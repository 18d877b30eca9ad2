#!/bin/bash
# round 2: (1) fused resample2d -> cosine op (f4) tests, (2) resample2d after the grad_input1 revert, (3) L2 residency variants of the fused backward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resample_cosine.py tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "resample or cosine or perceptual" > gpurun_out/r2o_pytest_rs.log 2>&1; echo "pytest resample/cosine rc=$?"; tail -15 gpurun_out/r2o_pytest_rs.log
timeout 300 python tools/l2_policy_bwd.py > gpurun_out/r2o_l2_policy.txt 2>&1; cat gpurun_out/r2o_l2_policy.txt
for kn in 0 2048 4096 8192 16384 20480 28672 30720; do
  GFLA_BWD_KNOBS=$kn timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_local_attn_bwd_fused -s 2 -c 1 --csv --log-file gpurun_out/r2o_dram_$kn.csv python tools/run_fwd.py --B 16 --bwd --iters 2 > /dev/null 2>&1
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/r2o_dram_*.csv'), key=lambda x: int(x.split('_')[-1].split('.')[0])):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    if len(rows) < 2: print(f, 'no data'); continue
    h = rows[0]; vals = {r[h.index('Metric Name')]: (r[h.index('Metric Value')], r[h.index('Metric Unit')]) for r in rows[1:]}
    print(f.split('_')[-1], vals)
PY
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.load(open('gpurun_out/r2o_bench.json'))
print(j['value'], j['ms_per_step'], 'fwd', j['roofline_fwd']['launch_ms'], 'bwd', j['roofline_bwd']['launch_ms'])
print('cfg3', json.dumps(j.get('cfg3'))[:900])
print('f4', json.dumps(j.get('f4_resample_cosine')))
PY
tail -n 5 gpurun_out/r2o_bench.err
