"""Kernel-only timings of the main entry points (CUDA events, inputs resident in HBM).
Usage: [KYBER_B200_LIB=path/to/variant.so] python tools/quick_bench.py [log2n] [--all]
Prints one JSON line; used to compare build variants and to fill profiles/."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

kb = importlib.import_module("kyber-rs_b200")
log2n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 18
n = 1 << log2n
ctx = kb.Context(0)
dev = torch.device("cuda", 0)
res = {"lib": os.path.basename(kb.LIB_PATH), "log2n": log2n}
for kind, name in ((1, "imad_lo32"), (0, "imad_wide"), (2, "imad_wide_carry_chain"), (3, "fe_mul_imad_eq")):
    res["probe_" + name + "_T_per_s"] = round(ctx.probe_imad(kind, 1 << 14)[0] / 1e12, 3)
pk, msg, off, sig, expect = bench.make_batch(ctx, n, 0)
d_pk, d_msg, d_sig = (torch.from_numpy(x).to(dev) for x in (pk, msg, sig))
d_off = torch.from_numpy(off.view(np.int64)).to(dev)
d_st = torch.empty(n, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


t = timed(lambda: ctx.dev_verify(n, d_pk, d_msg, d_off, d_sig, d_st))
assert (d_st.cpu().numpy() == expect).all(), "verify statuses wrong"
res["verify_eddsa_per_s"] = n / t
t = timed(lambda: ctx.dev_verify(n, d_pk, d_msg, d_off, d_sig, d_st, schnorr=True))
res["verify_schnorr_per_s"] = n / t
m = min(n, 1 << 16)
d_sc = d_sig[:m, 32:].contiguous()
d_o = torch.empty(m, 32, dtype=torch.uint8, device=dev)
d_pts = d_pk[:m].clone()
d_pts[63::64] = d_pk[0]
res["mul_base_ct_per_s"] = m / timed(lambda: ctx.dev_point_mul_base(m, d_sc, d_o, 0))
res["mul_base_vt_per_s"] = m / timed(lambda: ctx.dev_point_mul_base(m, d_sc, d_o, 1))
res["mul_var_ct_per_s"] = m / timed(lambda: ctx.dev_point_mul(m, d_sc, d_pts, d_o, d_st, 0))
res["mul_var_vt_per_s"] = m / timed(lambda: ctx.dev_point_mul(m, d_sc, d_pts, d_o, d_st, 1))
if n > m:   # the same kernels on the whole batch: 2^16 items leave a B200 partly idle (512 blocks on 148 SMs)
    d_scn = d_sig[:, 32:].contiguous()
    d_on = torch.empty(n, 32, dtype=torch.uint8, device=dev)
    d_ptn = d_pk.clone()
    d_ptn[63::64] = d_pk[0]
    res[f"mul_base_ct_2^{log2n}_per_s"] = n / timed(lambda: ctx.dev_point_mul_base(n, d_scn, d_on, 0), reps=3)
    res[f"mul_base_vt_2^{log2n}_per_s"] = n / timed(lambda: ctx.dev_point_mul_base(n, d_scn, d_on, 1), reps=3)
    res[f"mul_var_ct_2^{log2n}_per_s"] = n / timed(lambda: ctx.dev_point_mul(n, d_scn, d_ptn, d_on, d_st, 0), reps=3)
    res[f"mul_var_vt_2^{log2n}_per_s"] = n / timed(lambda: ctx.dev_point_mul(n, d_scn, d_ptn, d_on, d_st, 1), reps=3)
d_part = torch.empty(128, dtype=torch.uint8, device=dev)
d_enc = torch.empty(32, dtype=torch.uint8, device=dev)
d_bad = torch.zeros(1, dtype=torch.int64, device=dev)
d_mp = d_pk.clone()
d_mp[63::64] = d_pk[0]
d_ms = d_sig[:, 32:].contiguous()
for lg in sorted({16, min(log2n, 20), log2n}):
    k = 1 << lg
    res[f"msm_2^{lg}_points_per_s"] = k / timed(lambda: ctx.dev_msm(k, d_ms, d_mp, d_enc, d_part, d_bad), reps=3)
if "--all" in sys.argv:
    # config 3: VSS n=256, t=171 ; config 4 slice: 64 dealers of the n=1024, t=683 round
    for (nn, tt, nd, tag) in ((256, 171, 1, "cfg3_vss_n256_t171"), (1024, 683, 64, "cfg4_slice_64_dealers_n1024_t683")):
        coeff = bench.xof("kyber-b200/" + tag, 32 * nd * tt).reshape(-1, 32).copy()
        coeff[:, 31] &= 0x0F
        commits = torch.from_numpy(ctx.point_mul_base_batch(coeff)).to(dev)
        shares = torch.from_numpy(bench.xof("kyber-b200/sh" + tag, 32 * nd * nn).reshape(-1, 32).copy()).to(dev)
        verdict = torch.empty(nd * nn, dtype=torch.uint8, device=dev)
        tsec = timed(lambda: ctx.dev_dkg_verify_round(nn, tt, nd, commits, shares, verdict), reps=2)
        res[tag + "_share_checks_per_s"] = nd * nn / tsec
        res[tag + "_ms"] = tsec * 1e3
print(json.dumps(res))
