"""BASELINE config 4 (and 3): deal verification of a Pedersen DKG round, sharded by dealer.

  python tools/bench_dkg.py [--n 1024] [--t 683] [--reps 2]            # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_dkg.py

Every rank owns dealers [lo, hi): their commitments (t x 32 B each) and the n shares they sent.  It runs
kb_dev_dkg_verify_round on them; the verdict rows are all-gathered.  Dealers 0..3 get HONEST shares
(private polynomial evaluated with Python integers, 0.4 % of them corrupted), dealer 4 additionally carries a
torsion-contaminated commitment, the remaining dealers get random shares (verdict 0) — the
kernel's work does not depend on the verdict.  Prints one JSON line."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

L = bench.L_ORDER


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--t", type=int, default=683)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--shard-of", type=int, default=0, help="single GPU: run only the dealer shard rank 0 of this many ranks would own")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kb = importlib.import_module("kyber-rs_b200")
    ctx = kb.Context(local)
    dev = torch.device("cuda", local)
    n, t = a.n, a.t
    lo, hi = kb.sharding.shard_range(n, rank, world)
    if world == 1 and a.shard_of > 1:
        lo, hi = kb.sharding.shard_range(n, 0, a.shard_of)
    nd = hi - lo
    honest = min(5, n)
    # private coefficients of ALL dealers are derived per dealer, so every rank can build its own slice
    coeff = np.concatenate([bench.xof(f"kyber-b200/cfg4/dealer{d}", 32 * t).reshape(t, 32) for d in range(lo, hi)]).copy()
    coeff[:, 31] &= 0x0F
    commits = ctx.point_mul_base_batch(coeff)                                  # PriPoly::commit (poly.rs:195)
    shares = bench.xof(f"kyber-b200/cfg4/shares{rank}", 32 * nd * n).reshape(-1, 32).copy()
    shares[:, 31] &= 0x0F
    expect = np.zeros((nd, n), dtype=np.uint8)
    t0 = time.perf_counter()
    for d in range(lo, min(hi, honest)):
        c = [int.from_bytes(coeff[(d - lo) * t + j].tobytes(), "little") for j in range(t)]
        for i in range(n):
            v = 0
            for cj in reversed(c):
                v = (v * (i + 1) + cj) % L                                    # PriPoly::eval (poly.rs:133)
            shares[(d - lo) * n + i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
        expect[d - lo, :] = 1
        for i in range(d, n, 251):                                            # ~0.4 % corrupted
            shares[(d - lo) * n + i, 3] ^= 0x10
            expect[d - lo, i] = 0
    if lo <= 4 < min(hi, honest):                                             # torsion-contaminated commitment
        t8 = np.frombuffer(bytes.fromhex("26e8958fc2b227b045c3f489f2ef98f0d5dfac05d3c63339b13802886d53fc05"), dtype=np.uint8)
        k = (4 - lo) * t + 1
        commits[k] = ctx.point_add_batch(commits[k], t8)[0][0]
        expect[4 - lo, :] = 0      # x*(C1 + T8) + ... differs from the honest value unless 8 | x
        expect[4 - lo, 7::8] = 1
        for i in range(4, n, 251):
            expect[4 - lo, i] = 0
    prep_s = time.perf_counter() - t0
    d_commits = torch.from_numpy(commits).to(dev)
    d_shares = torch.from_numpy(shares).to(dev)
    d_verdict = torch.zeros(nd * n, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(n * n, dtype=torch.uint8, device=dev)

    def step():
        ctx.dev_dkg_verify_round(n, t, nd, d_commits, d_shares, d_verdict)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_verdict)                     # equal shards (n % world == 0)

    assert n % world == 0
    step()
    torch.cuda.synchronize()
    got = d_verdict.cpu().numpy().reshape(nd, n)
    ok = bool((got == expect).all())
    # every repetition is timed on its own; the MEDIAN is reported (a hiccup of the box in one repetition would
    # otherwise dominate the mean of a few 6 ms rounds), minimum and maximum next to it
    reps = max(a.reps, 3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for e0, e1 in evs:
        e0.record()
        step()
        e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    per_rep = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    ms = torch.tensor([per_rep[len(per_rep) // 2]], dtype=torch.float64, device=dev)
    okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    # end to end through the host-buffer entry point (H2D of commitments + shares, D2H of verdicts)
    v = np.zeros(n * n, dtype=np.uint8)
    full_commits = commits if world == 1 else None
    e2e_ms = None
    if world == 1 and a.shard_of <= 1:
        t0 = time.perf_counter()
        ctx.dkg_verify_round(n, t, full_commits, shares, verdict=v)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        ok = ok and bool((v.reshape(n, n) == expect).all())
    if rank == 0:
        checks = n * n if a.shard_of <= 1 else n * nd
        out = {"metric": "DKG deal-verification round" if a.shard_of <= 1 else f"DKG deal-verification, the shard of 1 rank of {a.shard_of}", "n": n, "t": t, "n_gpus": world, "round_ms": float(ms.item()), "share_checks_per_s": checks / (float(ms.item()) * 1e-3),
               "round_ms_min_max": [per_rep[0], per_rep[-1]], "reps": reps, "verdicts_match_expected": bool(okt.item()) and ok, "e2e_round_ms_host_buffers": e2e_ms, "honest_dealers_checked": honest, "prep_s": prep_s,
               "imad_eq_per_s_T": checks * t * 6800 / (float(ms.item()) * 1e-3) / 1e12}
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
