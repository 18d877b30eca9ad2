#!/bin/bash
# cfg4 (reference PoseGenerator fwd+bwd, DDP over NCCL) at $1 GPUs
n=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --workload cfg4 --arms fused,literal --steps 10 --warmup 3 > gpurun_out/r2_cfg4_n$n.json 2> gpurun_out/r2_cfg4_n$n.err; echo "cfg4 n=$n rc=$?"
grep '^{' gpurun_out/r2_cfg4_n$n.json | cut -c1-1500
tail -n 3 gpurun_out/r2_cfg4_n$n.err
