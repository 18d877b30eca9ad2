#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (run on the GPU box): the strip forward (channels-last tile tests),
# the per-tile forward (NCHW), the fused backward (default) and the two-kernel backward (GFLA_BWD_FUSED=0), the unfused ops (golden)
set -o pipefail
SEL="tile_vs_oracle and shape0 and smooth or bwd_tile_vs_oracle and shape0 and smooth or golden and k3 or blend_fwd and nhwc"
for tool in memcheck racecheck; do
  echo "== $tool (fused backward) =="
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -6
done
echo "== memcheck (two-kernel backward) =="
GFLA_BWD_FUSED=0 timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bwd_tile_vs_oracle and shape0 and smooth" 2>&1 | tail -4
