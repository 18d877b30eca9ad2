#!/bin/bash
# round 2, first GPU visit: full GPU test suite, bench baseline (+extras), L2-chunked backward experiment, cfg4/cfg5, resample2d ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
for cb in 1 2 4; do GFLA_BWD_CHUNK=$cb timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2a_bench_chunk$cb.json 2>>gpurun_out/r2a_bench.err; done
timeout 600 python bench.py --workload cfg4 --steps 5 > gpurun_out/r2a_cfg4.json 2> gpurun_out/r2a_cfg4.err; echo "cfg4 rc=$?"
timeout 600 python bench.py --workload cfg4 --model-dtype fp32 --steps 5 > gpurun_out/r2a_cfg4_fp32.json 2>> gpurun_out/r2a_cfg4.err
timeout 600 python bench.py --workload cfg5 --steps 3 > gpurun_out/r2a_cfg5.json 2> gpurun_out/r2a_cfg5.err; echo "cfg5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resample2d -c 3 -o gpurun_out/r2_resample2d_ks4 python tools/run_resample.py --ks 4 --B 4 > gpurun_out/r2a_ncu_rs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample2d -c 3 -o gpurun_out/r2_resample2d_ks2 python tools/run_resample.py --ks 2 --sigma 5 --B 4 >> gpurun_out/r2a_ncu_rs.log 2>&1
for f in gpurun_out/r2a_bench.json gpurun_out/r2a_bench_chunk*.json gpurun_out/r2a_cfg4.json gpurun_out/r2a_cfg4_fp32.json gpurun_out/r2a_cfg5.json; do echo "== $f"; head -c 3000 $f; echo; done
tail -5 gpurun_out/r2a_bench.err gpurun_out/r2a_cfg4.err gpurun_out/r2a_cfg5.err
