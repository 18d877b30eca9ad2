#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (run on the GPU box)
set -o pipefail
for tool in memcheck racecheck; do
  echo "== $tool =="
  timeout 600 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
      -k "tile_vs_oracle and shape0 and smooth or bwd_tile_vs_oracle and shape0 and smooth or golden and k3" 2>&1 | tail -6
done
