"""ncu target: one MSM of 2^LOG2N points (default 22) after a warm-up one.  Usage: python tools/profile_msm.py [log2n]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

kb = importlib.import_module("kyber-rs_b200")
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << lg
ctx = kb.Context(0)
dev = torch.device("cuda", 0)
seeds = bench.xof("kyber-b200/profile-msm", 32 * min(n, 1 << 20)).reshape(-1, 32).copy()
seeds[:, 31] &= 0x0F
d_seed = torch.from_numpy(seeds).to(dev)
d_pts = torch.empty(n, 32, dtype=torch.uint8, device=dev)
d_sc = torch.empty(n, 32, dtype=torch.uint8, device=dev)
for off in range(0, n, 1 << 20):
    k = min(1 << 20, n - off)
    tw = d_seed[:k].clone()
    tw[:, 0] = (tw[:, 0].to(torch.int32) + off // (1 << 20)).to(torch.uint8)
    ctx.dev_point_mul_base(k, tw, d_pts[off:off + k], 1)
    d_sc[off:off + k] = torch.roll(tw, 1, 0)
d_part = torch.empty(128, dtype=torch.uint8, device=dev)
d_enc = torch.empty(32, dtype=torch.uint8, device=dev)
d_bad = torch.zeros(1, dtype=torch.int64, device=dev)
for _ in range(2):
    ctx.dev_msm(n, d_sc, d_pts, d_enc, d_part, d_bad)
torch.cuda.synchronize()
print("msm", lg, d_enc.cpu().numpy().tobytes().hex(), int(d_bad.item()))
