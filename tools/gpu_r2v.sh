#!/bin/bash
# last visit: the shipped build once more (full GPU suite + smoke), e2e chunking variants, the final bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for cfg in "4 2" "8 2" "8 4" "16 4" "2 2"; do set -- $cfg
  timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 --e2e-chunks $1 --e2e-streams $2 2>> gpurun_out/r2v.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('chunks=$1 streams=$2', 'value', round(j['value'],1), 'e2e', round(j['e2e']['value'],2), 'ms', round(j['e2e']['ms_per_step'],2))"
done 2>&1 | tee gpurun_out/r2v_e2e.txt
tail -3 gpurun_out/r2v.err
