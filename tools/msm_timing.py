"""Kernel-only timing of kb_dev_msm / kb_dev_msm_ext at 2^17, 2^20, 2^22 points (A/B helper)."""
import importlib, sys, torch, numpy as np, os
sys.path.insert(0, '.')
kb = importlib.import_module("kyber-rs_b200")
ctx = kb.Context(0); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
out = []
for lg in (17, 20, 22):
    n = 1 << lg
    sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g); sc[:, 31] &= 0x0F
    ps = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g); ps[:, 31] &= 0x0F
    pts = torch.empty(n, 32, dtype=torch.uint8, device=dev); ctx.dev_point_mul_base(n, ps, pts, 1)
    raw = torch.empty(n, 128, dtype=torch.uint8, device=dev); st = torch.empty(n, dtype=torch.uint8, device=dev); ctx.dev_point_decompress(n, pts, raw, st)
    part = torch.empty(128, dtype=torch.uint8, device=dev); bad = torch.zeros(1, dtype=torch.int64, device=dev)
    for name, fn in (("enc", lambda: ctx.dev_msm(n, sc, pts, None, part, bad)), ("ext", lambda: ctx.dev_msm_ext(n, sc, raw, None, part, bad))):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        out.append(f"2^{lg} {name} {ms:.3f} ms {n / ms / 1e3:.1f} M/s")
print("msm", " | ".join(out))
