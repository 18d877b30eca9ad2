"""GPU debugging aid for the tile kernels: runs one small call with the host-mapped debug
buffer armed and prints which pipeline barrier timed out (if any) and the max error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gfla_b200
from gfla_b200 import _lib, functional as F_

def main():
    B, C, H, W, k = [int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (1, 64, 32, 32, 5))]
    dbg = torch.zeros(8, dtype=torch.int64).pin_memory()
    _lib.check(_lib.lib().gfla_debug_set_buffer(dbg.data_ptr()), "debug buffer")
    torch.manual_seed(0)
    s = torch.randn(B, C, H, W, device="cuda").bfloat16()
    if os.environ.get("NHWC"):
        s = s.contiguous(memory_format=torch.channels_last)
    f = (torch.rand(B, 2, H, W, device="cuda") * 8 - 4)
    l = torch.randn(B, k * k, H, W, device="cuda").bfloat16()
    ref = F_.local_attn_fwd(s, f, l, k, algo="gather").float()
    torch.cuda.synchronize()
    try:
        out = F_.local_attn_fwd(s, f, l, k, algo="tile").float()
        torch.cuda.synchronize()
        err = (out - ref).abs()
        print("tile ok: max err", err.max().item(), "mean", err.mean().item(), "ref absmax", ref.abs().max().item())
        if err.max().item() > 3e-3:
            bad = (err > 3e-3).nonzero()
            print("bad count", bad.shape[0], "first", bad[:8].tolist())
            print("out", out.flatten()[:8].tolist(), "ref", ref.flatten()[:8].tolist())
    except Exception as e:
        print("FAILED:", str(e).splitlines()[0])
    v = dbg.numpy().astype(np.uint64)
    print("dbg: tag=0x%x parity=%d iter=%d block=%d thread=%d" % (int(v[0]) & 0xffffff, v[1], v[2], v[3], v[4]), "armed" if int(v[0]) >> 63 else "(no timeout recorded)")

if __name__ == "__main__":
    main()
