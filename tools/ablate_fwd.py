"""Strip forward kernel with parts of the pipeline switched off (GFLA_TC_KNOBS; results are wrong, timing only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from gfla_b200 import functional as F_

dev = torch.device("cuda:0")
B, C, H, W, k = 16, 256, 256, 256, 5
src, flow, logits, gout = make_inputs(torch, dev, B, C, H, W, k, 1234, sys.argv[1] if len(sys.argv) > 1 else "smooth")
src = src.contiguous(memory_format=torch.channels_last).to(dev); flow = flow.to(dev); logits = logits.to(dev)
variants = [(0, "production"), (256, "no output stores"), (1024, "no weight scatter (slabs only zeroed)"), (2048, "no slab writes at all"),
            (4096, "no MMAs"), (8192, "no TMA loads"), (2048 | 4096, "no slab writes, no MMAs"), (4096 | 8192, "no MMAs, no TMA loads"),
            (2048 | 4096 | 8192, "no slab writes, no MMAs, no TMA loads"), (2048 | 4096 | 8192 | 256, "... and no output stores (skeleton)"),
            (16384, "production + L2 prefetch of upcoming rows")]
for knobs, name in variants:
    os.environ["GFLA_TC_KNOBS"] = str(knobs)
    for _ in range(3):
        F_.local_attn_fwd(src, flow, logits, k, algo="tile")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        F_.local_attn_fwd(src, flow, logits, k, algo="tile")
    b.record(); torch.cuda.synchronize()
    print(f"{knobs:6d}  {a.elapsed_time(b) / 10:.4f} ms  {name}", flush=True)
