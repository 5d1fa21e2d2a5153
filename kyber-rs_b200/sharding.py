"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed), host-side
partitioning, no point data on NVLink.

  signature / scalar-mult batches : contiguous index ranges, independent; status slices gathered
  DKG deal verification           : by dealer; verdict rows gathered
  Pippenger MSM                   : by points; each rank reduces its share to ONE uncompressed
                                    partial (128 bytes), all_gather of those, then every rank
                                    folds the <= world_size partials (point addition is not an
                                    NCCL reduce op) — the only data-path collective there is.

The compute steps are passed in as callables so the partition/gather logic can be exercised
with the `gloo` backend on CPU (tests/test_sharding_gloo.py) — the callables used in production
are the GPU entry points of binding.Context.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of n items for `rank` of `world` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def _dist():
    import torch.distributed as dist

    return dist


def gather_bytes(local: np.ndarray, counts, device=None) -> np.ndarray:
    """all_gather of variable-length uint8 slices (padded to the longest); returns the concatenation
    in rank order.  With world_size 1 or no process group this is the identity."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    import torch

    world = dist.get_world_size()
    width = max(counts)
    buf = torch.zeros(width, dtype=torch.uint8, device=device or "cpu")
    buf[: local.size] = torch.from_numpy(np.ascontiguousarray(local).reshape(-1)).to(buf.device)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([o[: counts[r]].cpu().numpy() for r, o in enumerate(outs)])


def verify_sharded(verify_fn, pk, msg, msg_off, sig, rank: int, world: int, device=None) -> np.ndarray:
    """Each rank verifies signatures [lo, hi); every rank returns the full status vector.
    verify_fn(pk, msg, msg_off, sig) -> uint8[n_local]."""
    n = pk.shape[0]
    lo, hi = shard_range(n, rank, world)
    off = np.asarray(msg_off, dtype=np.uint64)
    mlo, mhi = int(off[lo]), int(off[hi])
    local = verify_fn(pk[lo:hi], msg[mlo:mhi], off[lo:hi + 1] - off[lo], sig[lo:hi])
    counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    return gather_bytes(np.asarray(local, dtype=np.uint8), counts, device)


def dkg_round_sharded(round_fn, n: int, t: int, commits, shares, rank: int, world: int, device=None) -> np.ndarray:
    """Shard a deal-verification round by dealer.  round_fn(n, t, commits_slice, shares_slice) -> uint8[nd*n]
    for the rank's dealers; returns the full (ndealers*n) verdict vector on every rank."""
    ndealers = commits.shape[0] // t
    lo, hi = shard_range(ndealers, rank, world)
    local = round_fn(n, t, commits[lo * t:hi * t], shares[lo * n:hi * n])
    counts = [(shard_range(ndealers, r, world)[1] - shard_range(ndealers, r, world)[0]) * n for r in range(world)]
    return gather_bytes(np.asarray(local, dtype=np.uint8), counts, device)


def msm_sharded(partial_fn, fold_fn, scalars, points, rank: int, world: int, device=None) -> bytes:
    """Shard an MSM by points.  partial_fn(scalars, points) -> 128-byte uncompressed partial sum;
    fold_fn(partials[k,128]) -> 32-byte encoding of their sum.  One all_gather of 128 bytes per rank."""
    n = scalars.shape[0]
    lo, hi = shard_range(n, rank, world)
    part = np.asarray(partial_fn(scalars[lo:hi], points[lo:hi]), dtype=np.uint8).reshape(128)
    allp = gather_bytes(part, [128] * world, device).reshape(-1, 128)
    return fold_fn(allp)
