"""Run the fused forward (and optionally backward) a few times on synthetic cfg2-like inputs -- target for ncu."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from gfla_b200 import functional as F_

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4); ap.add_argument("--C", type=int, default=256)
ap.add_argument("--H", type=int, default=256); ap.add_argument("--W", type=int, default=256)
ap.add_argument("--k", type=int, default=5); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--flow", default="smooth"); ap.add_argument("--layout", default="nhwc"); ap.add_argument("--algo", default="auto")
ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
src, flow, logits, gout = make_inputs(torch, dev, a.B, a.C, a.H, a.W, a.k, 1234, a.flow)
if a.layout == "nhwc":
    src, gout = src.contiguous(memory_format=torch.channels_last), gout.contiguous(memory_format=torch.channels_last)
src, flow, logits, gout = (t.to(dev) for t in (src, flow, logits, gout))
for _ in range(a.iters):
    out = F_.local_attn_fwd(src, flow, logits, a.k, algo=a.algo)
    if a.bwd:
        F_.local_attn_bwd(src, flow, logits, gout, a.k)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = F_.local_attn_fwd(src, flow, logits, a.k, algo=a.algo)
    e1.record(); torch.cuda.synchronize()
    print("fwd ms/iter", e0.elapsed_time(e1) / 5)
if a.bwd:
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); g = F_.local_attn_bwd(src, flow, logits, gout, a.k); e1.record()
    out = F_.local_attn_fwd(src, flow, logits, a.k, algo=a.algo); e2.record(); torch.cuda.synchronize()
    print("bwd ms", e0.elapsed_time(e1), "fwd right after bwd ms", e1.elapsed_time(e2))
