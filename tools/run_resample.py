"""resample2d forward + backward a few times at a cfg3-shaped problem -- target for ncu (`-k regex:k_resample2d`)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gfla_b200 import functional as F_

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4); ap.add_argument("--C", type=int, default=128)
ap.add_argument("--H", type=int, default=512); ap.add_argument("--W", type=int, default=512)
ap.add_argument("--ks", type=int, default=4); ap.add_argument("--sigma", type=float, default=2.0); ap.add_argument("--iters", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(5)
coarse = torch.rand(a.B, 2, a.H // 16, a.W // 16, generator=g) * 16 - 8
flow = torch.nn.functional.interpolate(coarse, size=(a.H, a.W), mode="bilinear", align_corners=True).to(dev)
in2 = torch.cat([flow, torch.full((a.B, 1, a.H, a.W), a.sigma, device=dev)], 1).contiguous()
x = torch.randn(a.B, a.C, a.H, a.W, device=dev)
go = torch.randn(a.B, a.C, a.H, a.W, device=dev)
for _ in range(a.iters):
    out = F_.resample2d_fwd(x, in2, a.ks, 1)
    g1, g2 = F_.resample2d_bwd(x, in2, go, a.ks, 1)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()), float(g1.abs().mean()), float(g2.abs().mean()))
