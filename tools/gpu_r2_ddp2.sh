#!/bin/bash
# round 2, second multi-GPU visit: cfg4 (reference PoseGenerator fwd+bwd under DDP) at 2/4/8 GPUs
mkdir -p gpurun_out
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --workload cfg4 --arms fused,literal --steps 10 --warmup 3 > gpurun_out/r2_cfg4_n$n.json 2> gpurun_out/r2_cfg4_n$n.err; echo "cfg4 n=$n rc=$?"
done
for f in gpurun_out/r2_cfg4_n2.json gpurun_out/r2_cfg4_n4.json gpurun_out/r2_cfg4_n8.json; do echo "== $f"; grep '^{' $f | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['n_gpus'], j['value'], j['unit'], {k:(round(v['value'],1) if 'value' in v else v) for k,v in j['arms'].items()}, j.get('allreduce'), j['arms'].get('fused',{}).get('kernel_time_share'))"; done
tail -n 3 gpurun_out/r2_cfg4_n8.err
