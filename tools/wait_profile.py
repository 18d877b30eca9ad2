"""Where do the tile kernels' warps wait?  (gfla_debug_wait_profile; cfg2-like inputs; needs a GFLA_BUILD_PROFILE=1 build)

  --which 0   per-tile forward kernel (GFLA_TC_STRIP=-1)
  --which 1   strip forward kernel
  --which 2   fused backward kernel
Output: share of the kernel's cycles that lane 0 of each warp of a role spent blocked on each barrier kind / in each timed region.
"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from gfla_b200 import functional as F_, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16); ap.add_argument("--C", type=int, default=256)
ap.add_argument("--H", type=int, default=256); ap.add_argument("--W", type=int, default=256)
ap.add_argument("--k", type=int, default=5); ap.add_argument("--flow", default="smooth")
ap.add_argument("--which", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
src, flow, logits, gout = make_inputs(torch, dev, a.B, a.C, a.H, a.W, a.k, 1234, a.flow)
cl = torch.channels_last
src = src.contiguous(memory_format=cl).to(dev); gout = gout.contiguous(memory_format=cl).to(dev)
flow = flow.to(dev); logits = logits.to(dev)
lib = _lib.lib()
run = (lambda: F_.local_attn_bwd(src, flow, logits, gout, a.k, algo="tile")) if a.which == 2 else \
      (lambda: F_.local_attn_fwd(src, flow, logits, a.k, algo="tile"))
for _ in range(3):
    run()
assert lib.gfla_debug_wait_profile(a.which, 1, None) == 0
iters = 5
for _ in range(iters):
    run()
out = (ctypes.c_ulonglong * 64)()
assert lib.gfla_debug_wait_profile(a.which, 0, ctypes.cast(out, ctypes.c_void_p)) == 0
v = list(out)
total = v[7]
if a.which == 2:
    names = {0: ("producer", 1), 1: ("mma Q", 1), 5: ("mma gs", 1), 2: ("pixel team(x4)", 4), 3: ("slab builders(x4)", 4), 4: ("gs epilogue(x4)", 4)}
    kinds = {0: {2: "stage free", 6: "G free (prev group retired)"},
             1: {0: "source stage landed", 4: "Q acc drained", 5: "info", 7: "G landed"},
             5: {1: "slabs built", 5: "info", 6: "gs acc drained", 7: "G landed"},
             2: {3: "Q acc full", 5: "info", 6: "[busy] softmax/taps + next loads", 7: "[busy] finalize"},
             3: {2: "slab stage free", 5: "info", 6: "[busy] window build", 7: "[busy + wait for Q] finalize of the previous group"},
             4: {3: "gs acc full", 5: "info"}}
else:
    names = {0: ("producer", 1), 1: ("mma", 1), 2: ("builders(x8)" if a.which == 1 else "builders(x4)", 8 if a.which == 1 else 4), 3: ("epilogue(x4)", 4)}
    k = {0: "full_s", 1: "full_a", 2: "empty", 3: "acc_full", 4: "acc_empty", 5: "info", 6: "region6", 7: "region7"}
    kinds = {r: dict(k) for r in range(4)}
if total == 0:
    print("raw counters:", v)
print(f"kernel cycles per CTA per launch: {total / iters / 148:.0f}")
for r, (nm, nw) in names.items():
    parts = []
    for kk in range(8):
        if r == 0 and kk == 7:
            continue
        if v[r * 8 + kk]:
            parts.append(f"{kinds[r].get(kk, 'kind%d' % kk)} {100.0 * v[r * 8 + kk] / nw / total:.1f}%")
    print(f"{nm:18s} " + ", ".join(parts))
