"""Where do the forward tile kernels' warps wait?  (gfla_debug_wait_profile; cfg2-like inputs)"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from gfla_b200 import functional as F_, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16); ap.add_argument("--C", type=int, default=256)
ap.add_argument("--H", type=int, default=256); ap.add_argument("--W", type=int, default=256)
ap.add_argument("--k", type=int, default=5); ap.add_argument("--flow", default="smooth")
ap.add_argument("--which", type=int, default=0, help="0 per-tile kernel, 1 strip kernel (set GFLA_TC_STRIP accordingly)")
a = ap.parse_args()
dev = torch.device("cuda:0")
src, flow, logits, _ = make_inputs(torch, dev, a.B, a.C, a.H, a.W, a.k, 1234, a.flow)
src = src.contiguous(memory_format=torch.channels_last).to(dev); flow = flow.to(dev); logits = logits.to(dev)
lib = _lib.lib()
for _ in range(3):
    F_.local_attn_fwd(src, flow, logits, a.k, algo="tile")
assert lib.gfla_debug_wait_profile(a.which, 1, None) == 0
iters = 5
for _ in range(iters):
    F_.local_attn_fwd(src, flow, logits, a.k, algo="tile")
out = (ctypes.c_ulonglong * 32)()
assert lib.gfla_debug_wait_profile(a.which, 0, ctypes.cast(out, ctypes.c_void_p)) == 0
v = list(out)
total = v[7]
names = {0: "producer", 1: "mma", 2: "builders(x4)", 3: "epilogue(x4)"}
kinds = {0: "full_s", 1: "full_a", 2: "empty", 3: "acc_full", 4: "acc_empty", 5: "info", 6: "region6", 7: "region7"}
print(f"kernel cycles per CTA per launch: {total / iters / 148:.0f}")
for r in range(4):
    nw = 4 if r >= 2 else 1
    parts = []
    for k in range(8):
        if r == 0 and k == 7:
            continue
        if v[r * 8 + k]:
            parts.append(f"{kinds[k]} {100.0 * v[r * 8 + k] / nw / total:.1f}%")
    print(f"{names[r]:14s} " + ", ".join(parts))
