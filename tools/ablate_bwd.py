"""Fused backward with parts of the pipeline switched off (GFLA_BWD_KNOBS; results are wrong, timing only): what is the
kernel's time made of?  cfg2 shape, CUDA events, 10 launches per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from gfla_b200 import functional as F_

os.environ["GFLA_BWD_FUSED"] = "1"
dev = torch.device("cuda:0")
B, C, H, W, k = 16, 256, 256, 256, 5
src, flow, logits, gout = make_inputs(torch, dev, B, C, H, W, k, 1234, sys.argv[1] if len(sys.argv) > 1 else "smooth")
cl = torch.channels_last
src = src.contiguous(memory_format=cl).to(dev); gout = gout.contiguous(memory_format=cl).to(dev); flow = flow.to(dev); logits = logits.to(dev)
variants = [(0, "production"), (2, "no Q extraction (pixel team)"), (4, "no gs staging / reduce-add"), (8, "no slab fills"), (16, "no Q MMAs"),
            (32, "no gs MMAs"), (64, "no source-row loads"), (2 | 16 | 64, "no Q chain at all"), (4 | 8 | 32, "no gs chain at all"),
            (2 | 4 | 8 | 16 | 32 | 64, "skeleton only"), (126 | 128, "skeleton, pixel team without softmax / finalize / stores"),
            (126 | 256, "skeleton, builders without softmax / window"), (126 | 512, "skeleton, gs epilogue without irregular check / TMEM reads"),
            (126 | 1024, "skeleton, no grad_out tile load / TMEM copy"), (126 | 128 | 256 | 512 | 1024, "barriers and schedule only"),
            (128, "production but pixel team without softmax / finalize"), (1, "production + L2 prefetch of next rows")]
for knobs, name in variants:
    os.environ["GFLA_BWD_KNOBS"] = str(knobs)
    for _ in range(3):
        F_.local_attn_bwd(src, flow, logits, gout, k, algo="tile")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        F_.local_attn_bwd(src, flow, logits, gout, k, algo="tile")
    b.record(); torch.cuda.synchronize()
    print(f"{knobs:4d}  {a.elapsed_time(b) / 10:.4f} ms  {name}", flush=True)
