#!/bin/bash
# gpurun_ab/*.so are variant builds made on the build host before the call (git worktree of the variant -> build.py -> copy;
# GFLA_BUILD_PROFILE=1 / GFLA_BUILD_KNOBS=1 for the profile / tuning builds); the directory is git-ignored (*.so) and not kept.
# A/B on one box: micro-variants of the tile kernels, each as a prebuilt library (gpurun_ab/), timing first, then the tile tests per variant
mkdir -p gpurun_out
run() { GFLA_LIB=$PWD/gpurun_ab/libgfla_$1.so timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 20 2>> gpurun_out/r2t.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$1', j['value'], 'fwd', j['roofline_fwd']['launch_ms'], 'bwd', j['roofline_bwd']['launch_ms'], 'nchw', j['planar_nchw']['ms_per_step'])"; }
for rep in 1 2; do for v in head v1 v2 v4 v124; do run $v; done; done 2>&1 | tee gpurun_out/r2t_ab.txt
for v in v124 v1 v2 v4; do
  GFLA_LIB=$PWD/gpurun_ab/libgfla_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile or strip or irregular or blend or many_samples" > gpurun_out/r2t_pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -1 gpurun_out/r2t_pytest_$v.log
done
tail -3 gpurun_out/r2t.err
