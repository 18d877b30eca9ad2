#!/bin/bash
# gpurun_ab/*.so are variant builds made on the build host before the call (git worktree of the variant -> build.py -> copy;
# GFLA_BUILD_PROFILE=1 / GFLA_BUILD_KNOBS=1 for the profile / tuning builds); the directory is git-ignored (*.so) and not kept.
# A/B on one box: working tree vs the committed HEAD build (gpurun_ab/libgfla_head.so), then the GPU suite
mkdir -p gpurun_out
run() { timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 20 2>> gpurun_out/r2r.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$1', j['value'], 'fwd', j['roofline_fwd']['launch_ms'], 'bwd', j['roofline_bwd']['launch_ms'], 'nchw', j['planar_nchw']['ms_per_step'])"; }
for rep in 1 2; do
run "new(fastdiv group decode)"
GFLA_LIB=$PWD/gpurun_ab/libgfla_head.so run "previous(HEAD)"
done 2>&1 | tee gpurun_out/r2r_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refcuda.py -m gpu -x -q -k "bwd or backward or cfg2" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2r_pytest.log
tail -3 gpurun_out/r2r.err
