"""Per-op timings at BASELINE.json's configurations (CUDA events, warm-up 3, mean of N) -> JSON lines.
Secondary to bench.py (which carries the headline contract); results are copied into profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gfla_b200 import functional as F_
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
dev = "cuda:0"

def timed(fn, n=5, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def emit(name, ms, px, alg_bytes, **kw):
    print(json.dumps({"op": name, "ms": round(ms, 4), "Mpixels_per_s": round(px / ms / 1e3, 1),
                      "algorithmic_GB": round(alg_bytes / 1e9, 3), "GBps": round(alg_bytes / ms / 1e6, 1),
                      "frac_of_measured_hbm_peak": round(alg_bytes / ms / 1e6 / PEAK, 4), **kw}), flush=True)

def smooth(B, H, W):
    c = (torch.rand(B, 2, H // 16, W // 16, device=dev) * 16 - 8)
    return torch.nn.functional.interpolate(c, size=(H, W), mode="bilinear", align_corners=True).contiguous()

which = sys.argv[1:] or ["cfg3", "cfg2_unfused", "cfg2_fp32", "module"]
torch.manual_seed(0)
if "cfg3" in which:   # resample2d fwd+bwd, B=32 C=128 512x512 fp32 (BASELINE.json configs[2]); ks=2 sigma=5 and ks=4 sigma=2
    B, C, H, W = 32, 128, 512, 512
    x = torch.randn(B, C, H, W, device=dev)
    g = torch.randn(B, C, H, W, device=dev)
    # "blocky" = nearest-neighbour x16 up-sampling of a coarse flow + sub-pixel jitter, the shape PerceptualCorrectness feeds
    # (external_function.py:243: F.interpolate(flow, [h, w]), default mode 'nearest'): constant integer tap shift per block
    coarse = (torch.rand(B, 2, H // 16, W // 16, device=dev) * 16 - 8).floor() + 0.25
    blocky = (torch.nn.functional.interpolate(coarse, size=(H, W), mode="nearest") + 0.5 * torch.rand(B, 2, H, W, device=dev)).contiguous()
    for ks, sigma in ((2, 5.0), (4, 2.0)):
        px = B * H * W
        for fname, fl in (("smooth", smooth(B, H, W)), ("blocky", blocky)):
            in2 = torch.cat([fl, torch.full((B, 1, H, W), sigma, device=dev)], 1).contiguous()
            f = timed(lambda: F_.resample2d_fwd(x, in2, ks, 1))
            b = timed(lambda: F_.resample2d_bwd(x, in2, g, ks, 1), n=3, w=1)
            emit(f"resample2d_fwd ks={ks} {fname} flow", f, px, px * (2 * C * 4 + 12), config="cfg3 B=32 C=128 512x512 fp32")
            emit(f"resample2d_bwd ks={ks} {fname} flow", b, px, px * (3 * C * 4 + 24), config="cfg3")
            emit(f"resample2d_fwd+bwd ks={ks} {fname} flow", f + b, px, px * (5 * C * 4 + 36), config="cfg3")
if "cfg2_unfused" in which:   # the unfused ops at cfg2-like size (B=2: the [B,C,kH,kW] block tensor is 25x the input)
    B, C, H, W, k = 2, 256, 256, 256, 5
    s = torch.randn(B, C, H, W, device=dev).bfloat16(); fl = smooth(B, H, W)
    px = B * H * W
    t = timed(lambda: F_.block_extract_fwd(s, fl, k))
    emit("block_extract_fwd bf16 k=5", t, px, px * (C * 2 + C * 2 * k * k + 8), config="B=2 C=256 256x256")
    a = torch.randn(B, k * k, H, W, device=dev).bfloat16()
    t = timed(lambda: F_.attn_reshape_fwd(a, k))
    emit("attn_reshape_fwd bf16 k=5", t, px, px * (2 * k * k * 2), config="B=2 256x256")
if "cfg2_fp32" in which:   # fused op in fp32 (CUDA-core gather kernels), B=4
    B, C, H, W, k = 4, 256, 256, 256, 5
    s = torch.randn(B, C, H, W, device=dev); fl = smooth(B, H, W); l = torch.randn(B, k * k, H, W, device=dev); g = torch.randn(B, C, H, W, device=dev)
    px = B * H * W
    t = timed(lambda: F_.local_attn_fwd(s, fl, l, k))
    emit("local_attn_fwd fp32 (gather kernel)", t, px, px * (2 * C * 4 + 8 + k * k * 4), config="B=4 C=256 256x256 k=5")
    t = timed(lambda: F_.local_attn_bwd(s, fl, l, g, k), n=2, w=1)
    emit("local_attn_bwd fp32 (gather kernel, scalar atomics)", t, px, px * (3 * C * 4 + 16 + 2 * k * k * 4), config="B=4 C=256 256x256 k=5")

if "module" in which:   # ExtractorAttn at the shapes the pose generator uses: fused module vs the literal op chain
    import gfla_b200
    for (C, HW, k) in ((256, 32, 3), (128, 64, 5)):
        B = 8
        cl = torch.channels_last
        m = gfla_b200.ExtractorAttn(C, k, softmax=True).to(dev).bfloat16().to(memory_format=cl)
        src = torch.randn(B, C, HW, HW, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_()
        tgt = torch.randn(B, C, HW, HW, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_()
        flow = (torch.rand(B, 2, HW, HW, device=dev) * 8 - 4).requires_grad_()
        ex, rs = gfla_b200.BlockExtractor(k), gfla_b200.LocalAttnReshape()

        def ours(train):
            out = m(src, tgt, flow)
            if train: out.float().sum().backward()

        def literal(train):   # base_function.py:804-810 verbatim, on our unfused kernels
            bs = ex(src, flow); bt = ex(tgt, torch.zeros_like(flow))
            attn = m.fully_connect_layer(torch.cat((bt, bs), 1))
            out = torch.nn.functional.avg_pool2d(rs(attn, k) * bs, k, k)
            if train: out.float().sum().backward()

        px = B * HW * HW
        for name, fn in (("ExtractorAttn (fused tail, this library)", ours), ("literal reference op chain on our unfused kernels", literal)):
            with torch.no_grad():
                f = timed(lambda: fn(False))
            fb = timed(lambda: fn(True), n=3, w=2)
            print(json.dumps({"op": name, "config": f"B={B} C={C} {HW}x{HW} k={k} bf16 channels_last", "fwd_ms": round(f, 4),
                              "fwd_bwd_ms": round(fb, 4), "Mpixels_per_s_fwd_bwd": round(px / fb / 1e3, 2)}), flush=True)
