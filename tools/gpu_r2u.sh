#!/bin/bash
# resample2d gather kernels with / without the shared-memory source box (GFLA_RS_BOX): parity under both, cfg3 + f4 timings under both
mkdir -p gpurun_out
for box in 1 0; do
  GFLA_RS_BOX=$box timeout 600 python -m pytest tests/test_gpu_resample_cosine.py tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "resample or cosine or perceptual or cfg3" > gpurun_out/r2u_pytest_box$box.log 2>&1; echo "pytest box=$box rc=$?"; tail -1 gpurun_out/r2u_pytest_box$box.log
done
for box in 1 0 1 0; do
  GFLA_RS_BOX=$box timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 5 2>> gpurun_out/r2u.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); c=j['cfg3']; f=j['f4_resample_cosine']
print('box=$box', 'ks2 fwd/bwd', round(c['ks2']['fwd_ms'],2), round(c['ks2']['bwd_ms'],2), 'ks4 fwd/bwd', round(c['ks4']['fwd_ms'],2), round(c['ks4']['bwd_ms'],2), 'f4 relu3_1', round(f['relu3_1']['fused_ms'],3), 'relu2_1', round(f['relu2_1']['fused_ms'],3), 'unfused', round(f['relu2_1']['unfused_ms'],3))"
done 2>&1 | tee gpurun_out/r2u_ab.txt
tail -3 gpurun_out/r2u.err
