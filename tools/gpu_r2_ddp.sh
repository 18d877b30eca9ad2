#!/bin/bash
# round 2, multi-GPU visit (gpurun --gpus 8): BASELINE configs 4 and 5 at 1/2/4/8 GPUs (DDP gradient all-reduce over NCCL/NVLink),
# and the cfg2 line at 8 GPUs (e2e with NUMA-bound pinned buffers)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
for n in 1 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --workload cfg4 --arms fused,literal --steps 10 --warmup 3 > gpurun_out/r2_cfg4_n$n.json 2> gpurun_out/r2_cfg4_n$n.err; echo "cfg4 n=$n rc=$?"
done
for n in 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --workload cfg5 --arms fused,literal --steps 3 --warmup 3 > gpurun_out/r2_cfg5_n$n.json 2> gpurun_out/r2_cfg5_n$n.err; echo "cfg5 n=$n rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_cfg2_n8.json 2> gpurun_out/r2_cfg2_n8.err; echo "cfg2 n=8 rc=$?"
for f in gpurun_out/r2_cfg4_n*.json gpurun_out/r2_cfg5_n*.json; do echo "== $f"; python -c "
import json
j=json.load(open('$f')); print(j['n_gpus'], j['value'], j['unit'], {k:(round(v['value'],1) if 'value' in v else v) for k,v in j['arms'].items()}, j.get('allreduce'), j['arms'].get('fused',{}).get('kernel_time_share'))"; done
python -c "
import json
j=json.load(open('gpurun_out/r2_cfg2_n8.json')); print('cfg2 n=8', j['value'], 'e2e', j['e2e'])"
tail -n 3 gpurun_out/r2_cfg4_n8.err gpurun_out/r2_cfg5_n8.err gpurun_out/r2_cfg2_n8.err
