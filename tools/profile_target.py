"""Small, fixed workload for ncu: one launch of each hot kernel on 2^16 items (bench.py's generator).
Usage: python tools/profile_target.py [log2n]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

kb = importlib.import_module("kyber-rs_b200")
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << log2n
ctx = kb.Context(0)
pk, msg, off, sig, expect = bench.make_batch(ctx, n, 0)
dev = torch.device("cuda", 0)
d_pk, d_msg, d_sig = (torch.from_numpy(x).to(dev) for x in (pk, msg, sig))
d_off = torch.from_numpy(off.view(np.int64)).to(dev)
d_st = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(2):
    ctx.dev_verify(n, d_pk, d_msg, d_off, d_sig, d_st)
torch.cuda.synchronize()
assert (d_st.cpu().numpy() == expect).all()
d_sc = d_sig[:, 32:].contiguous()
d_o = torch.empty(n, 32, dtype=torch.uint8, device=dev)
d_pts = d_pk.clone()
d_pts[63::64] = d_pk[0]
ctx.dev_point_mul_base(n, d_sc, d_o, 0)
ctx.dev_point_mul_base(n, d_sc, d_o, 1)
ctx.dev_point_mul(n, d_sc, d_pts, d_o, d_st, 0)
ctx.dev_point_mul(n, d_sc, d_pts, d_o, d_st, 1)
d_part = torch.empty(128, dtype=torch.uint8, device=dev)
d_enc = torch.empty(32, dtype=torch.uint8, device=dev)
d_bad = torch.zeros(1, dtype=torch.int64, device=dev)
ctx.dev_msm(n, d_sc, d_pts, d_enc, d_part, d_bad)
torch.cuda.synchronize()
print("profile target done", n, ctx.launches)
