"""BASELINE config 5: Pippenger MSM sweep, total n = 2^16 .. 2^LOG2MAX points sharded by points across the
ranks; each rank reduces its shard to one partial (kb_dev_msm), the 128-byte partials are all-gathered with
NCCL and folded on every rank (kb_dev_point_sum) — the only data-path collective.  One JSON line."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2max", type=int, default=24)
    ap.add_argument("--reps", type=int, default=2)
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
    nmax = (1 << a.log2max) // world
    # points = s_i * B made on the GPU with the parity-checked fixed-base kernel; scalars random < 2^252
    base = 1 << 20
    seeds = bench.xof(f"kyber-b200/cfg5/rank{rank}", 32 * min(nmax, base)).reshape(-1, 32).copy()
    seeds[:, 31] &= 0x0F
    d_seed = torch.from_numpy(seeds).to(dev)
    d_pts = torch.empty(nmax, 32, dtype=torch.uint8, device=dev)
    d_sc = torch.empty(nmax, 32, dtype=torch.uint8, device=dev)
    for off in range(0, nmax, base):
        k = min(base, nmax - off)
        tweak = d_seed[:k].clone()
        tweak[:, 0] = (tweak[:, 0].to(torch.int32) + off // base).to(torch.uint8)     # distinct scalars per block
        ctx.dev_point_mul_base(k, tweak, d_pts[off:off + k], 1)
        d_sc[off:off + k] = torch.roll(tweak, 1, 0)
    d_part = torch.empty(128, dtype=torch.uint8, device=dev)
    d_all = torch.empty(world * 128, dtype=torch.uint8, device=dev)
    d_enc = torch.empty(32, dtype=torch.uint8, device=dev)
    d_bad = torch.zeros(1, dtype=torch.int64, device=dev)
    res = {}
    for lg in range(16, a.log2max + 1, 2):
        n_rank = (1 << lg) // world
        if n_rank == 0:
            continue

        def step():
            ctx.dev_msm(n_rank, d_sc, d_pts, None, d_part, d_bad)
            if world > 1:
                dist.all_gather_into_tensor(d_all, d_part)
                ctx.dev_point_sum(world, d_all, d_enc)
            else:
                ctx.dev_point_sum(1, d_part, d_enc)

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        enc = d_enc.cpu().numpy().tobytes().hex()
        assert int(d_bad.item()) == 0
        res[f"2^{lg}"] = {"ms": float(ms.item()), "points_per_s": (1 << lg) / (float(ms.item()) * 1e-3), "result": enc[:16]}
    if rank == 0:
        os.write(real_stdout, (json.dumps({"metric": "MSM points/sec", "n_gpus": world, "sweep": res}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
