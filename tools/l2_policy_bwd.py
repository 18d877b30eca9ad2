"""Fused backward: L2 residency variants of the in-kernel zero fill / reduce-add / grad_out loads (GFLA_BWD_KNOBS bits 11-14).
Timing with CUDA events here; DRAM bytes per variant come from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over
`tools/run_fwd.py --bwd` with the same knob value (tools/gpu_r2o.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from gfla_b200 import functional as F_

dev = torch.device("cuda:0")
B, C, H, W, k = 16, 256, 256, 256, 5
src, flow, logits, gout = make_inputs(torch, dev, B, C, H, W, k, 1234, "smooth")
cl = torch.channels_last
src = src.contiguous(memory_format=cl).to(dev); gout = gout.contiguous(memory_format=cl).to(dev); flow = flow.to(dev); logits = logits.to(dev)
variants = [(0, "production"), (2048, "zero fill one sample ahead"), (4096, "zero stores evict_last"), (8192, "grad_out loads evict_first"),
            (16384, "reduce-adds evict_last"), (4096 | 16384, "zero + reduce evict_last"), (4096 | 8192 | 16384, "all three hints"),
            (2048 | 4096 | 8192 | 16384, "all three hints, one sample ahead")]
for knobs, name in variants:
    os.environ["GFLA_BWD_KNOBS"] = str(knobs)
    for _ in range(3):
        F_.local_attn_bwd(src, flow, logits, gout, k, algo="tile")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        F_.local_attn_bwd(src, flow, logits, gout, k, algo="tile")
    b.record(); torch.cuda.synchronize()
    print(f"{knobs:6d}  {a.elapsed_time(b) / 20:.4f} ms  {name}", flush=True)
