"""ncu target: one deal-verification round (default n=256, t=171: 65 536 share checks) after a warm-up one.
Usage: python tools/profile_dkg.py [n t]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

kb = importlib.import_module("kyber-rs_b200")
n, t = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 171)
ctx = kb.Context(0)
dev = torch.device("cuda", 0)
coeff = bench.xof("kyber-b200/profile-dkg", 32 * n * t).reshape(-1, 32).copy()
coeff[:, 31] &= 0x0F
commits = torch.from_numpy(ctx.point_mul_base_batch(coeff)).to(dev)
shares = bench.xof("kyber-b200/profile-dkg/shares", 32 * n * n).reshape(-1, 32).copy()
shares[:, 31] &= 0x0F
d_sh = torch.from_numpy(shares).to(dev)
d_v = torch.zeros(n * n, dtype=torch.uint8, device=dev)
for _ in range(2):
    ctx.dev_dkg_verify_round(n, t, n, commits, d_sh, d_v)
torch.cuda.synchronize()
print("dkg profile target done", int(d_v.sum().item()))
