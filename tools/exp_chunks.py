import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from gfla_b200 import functional as F_
dev = torch.device("cuda:0")
B, C, H, W, k = 16, 256, 256, 256, 5
src, flow, logits, gout = make_inputs(torch, dev, B, C, H, W, k, 1234, "smooth")
src = src.contiguous(memory_format=torch.channels_last)
src, flow, logits = (t.to(dev) for t in (src, flow, logits))
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for chunk in (16, 8, 4, 2):
    def run():
        for b0 in range(0, B, chunk):
            F_.local_attn_fwd(src[b0:b0 + chunk], flow[b0:b0 + chunk], logits[b0:b0 + chunk], k)
    print(f"B=16 in chunks of {chunk}: {timed(run):.3f} ms total")
# same 4 samples repeatedly vs 4 different slices
print("slice [0:4] x4:", timed(lambda: [F_.local_attn_fwd(src[0:4], flow[0:4], logits[0:4], k) for _ in range(4)]))
print("slice [12:16] x4:", timed(lambda: [F_.local_attn_fwd(src[12:16], flow[12:16], logits[12:16], k) for _ in range(4)]))
