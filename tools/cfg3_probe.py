import importlib, sys, os, time, torch, numpy as np
sys.path.insert(0, '.')
import bench, bench_configs
kb = importlib.import_module("kyber-rs_b200")
ctx = kb.Context(0); dev = torch.device("cuda", 0)
env = {"torch": torch, "xof": bench.xof, "dev": dev}
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for (n, t) in ((256, 171), (1024, 683)):
    coeff, d_commits, d_shares, expect = bench_configs.build_round(env, ctx, n, t, 0, n, "probe")
    d_v = torch.zeros(n * n, dtype=torch.uint8, device=dev)
    print(n, t, "first", timed(lambda: ctx.dev_dkg_verify_round(n, t, n, d_commits, d_shares, d_v)), flush=True)
    ok = (d_v.cpu().numpy().reshape(n, n) == expect).all()
    print(" verdicts ok", ok, "again", timed(lambda: ctx.dev_dkg_verify_round(n, t, n, d_commits, d_shares, d_v)), flush=True)
    # random (dishonest) shares like tools/bench_dkg.py uses for most dealers
    d_sh2 = torch.randint(0, 256, (n * n, 32), dtype=torch.uint8, device=dev); d_sh2[:, 31] &= 0x0F
    print(" random shares", timed(lambda: ctx.dev_dkg_verify_round(n, t, n, d_commits, d_sh2, d_v)), flush=True)
