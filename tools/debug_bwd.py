"""GPU debugging aid: backward tile path vs the CUDA-core gather backward on one small channels_last call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gfla_b200 import _lib, functional as F_

def main():
    B, C, H, W, k = [int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (1, 64, 32, 32, 5))]
    dbg = torch.zeros(8, dtype=torch.int64).pin_memory()
    _lib.check(_lib.lib().gfla_debug_set_buffer(dbg.data_ptr()), "debug buffer")
    torch.manual_seed(0)
    cl = torch.channels_last
    s = torch.randn(B, C, H, W, device="cuda").bfloat16().contiguous(memory_format=cl)
    f = (torch.rand(B, 2, H, W, device="cuda") * 8 - 4)
    l = torch.randn(B, k * k, H, W, device="cuda").bfloat16()
    g = torch.randn(B, C, H, W, device="cuda").bfloat16().contiguous(memory_format=cl)
    rs, rf, rl = F_.local_attn_bwd(s, f, l, g, k, algo="gather")
    torch.cuda.synchronize()
    try:
        gs, gf, gl = F_.local_attn_bwd(s, f, l, g, k, algo="tile")
        torch.cuda.synchronize()
        for name, a, b in (("grad_source", gs, rs), ("grad_flow", gf, rf), ("grad_logits", gl, rl)):
            err = (a.float() - b.float()).abs()
            print(f"{name}: max err {err.max().item():.5f} mean {err.mean().item():.6f} ref absmax {b.float().abs().max().item():.4f}")
        import oracle.oracle as orc
        orc.build(ref=False)
        O = orc.Oracle()
        h = lambda t_: np.ascontiguousarray(t_.detach().float().cpu().numpy())
        ogs, ogf, ogl = O.local_attn_bwd(h(s), h(f), h(l), h(g), k)
        print("vs fp32 oracle: tile grad_source max err %.5f | gather grad_source max err %.5f | oracle absmax %.4f" % (
            np.abs(h(gs) - ogs).max(), np.abs(h(rs) - ogs).max(), np.abs(ogs).max()))
        e2 = np.abs(h(gs) - ogs)
        bad2 = np.argwhere(e2 > 0.02)
        print("tile-vs-oracle bad", len(bad2), bad2[:8].tolist())
        for b_, c_, y_, x_ in bad2[:4]:
            print("  at", (b_, c_, y_, x_), "tile", h(gs)[b_, c_, y_, x_], "oracle", ogs[b_, c_, y_, x_], "gather", h(rs)[b_, c_, y_, x_])
        err = (gs.float() - rs.float()).abs()
        if err.max().item() > 0.05:
            bad = (err > 0.05).nonzero()
            print("bad count", bad.shape[0], "of", err.numel(), "first", bad[:6].tolist())
    except Exception as e:
        print("FAILED:", str(e).splitlines()[0])
    v = dbg.numpy().astype(np.uint64)
    print("dbg: tag=0x%x parity=%d iter=%d block=%d thread=%d" % (int(v[0]) & 0xffffff, v[1], v[2], v[3], v[4]), "armed" if int(v[0]) >> 63 else "(no timeout recorded)")

if __name__ == "__main__":
    main()
