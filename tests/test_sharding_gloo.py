"""world_size-2 `gloo` run of the multi-GPU host logic (partition, gather, final MSM fold) on CPU.
The compute callables are the ORACLE here (checker role only); in production they are the GPU
entry points of kyber-rs_b200.binding.Context (see bench.py)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from helpers import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    import torch.distributed as dist

    from helpers import load_c_oracle, load_sign_input, make_sig_batch, pack_batch
    from oracle import ed25519_bigint as O

    kb = importlib.import_module("kyber-rs_b200")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    C = load_c_oracle()
    recs = load_sign_input()
    # 1. signature batch sharded by index
    pks, msgs, sigs = make_sig_batch(recs[:64], 101, bad_every=3)
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    full = kb.sharding.verify_sharded(lambda a, b, c, d: C.verify_batch(a, b, c, d), pk, flat, off, sg, rank, world)
    want = C.verify_batch(pk, flat, off, sg)
    ok1 = bool((full == want).all())
    # 2. MSM sharded by points: partial = uncompressed (x, y, 1, xy) of the oracle's partial sum
    n = 37
    rng = np.random.default_rng(5)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x7F
    good = np.frombuffer(b"".join(r[1] for r in recs[:64]), dtype=np.uint8).reshape(-1, 32)
    pts = good[:n]

    def partial_fn(s, p):
        x, y = O.msm([bytes(r) for r in s], [O.point_decode(bytes(r)) for r in p])
        words = b"".join(v.to_bytes(32, "little") for v in (x, y, 1, x * y % O.P))
        return np.frombuffer(words, dtype=np.uint8)

    def fold_fn(parts):
        acc = O.IDENTITY
        for row in parts:
            raw = row.tobytes()
            x, y = int.from_bytes(raw[:32], "little"), int.from_bytes(raw[32:64], "little")
            acc = O.point_add(acc, (x, y))
        return O.point_encode(acc)

    got = kb.sharding.msm_sharded(partial_fn, fold_fn, sc, pts, rank, world)
    ok2 = got == C.msm(sc, pts)
    # 3. DKG round sharded by dealer
    ndeal, nver, t = 5, 4, 3
    commits = good[: ndeal * t]
    shares = rng.integers(0, 256, size=(ndeal * nver, 32), dtype=np.uint8)

    def round_fn(nv, tt, cs, sh):
        nd = cs.shape[0] // tt
        out = np.zeros(nd * nv, dtype=np.uint8)
        for d in range(nd):
            out[d * nv:(d + 1) * nv] = C.vss_verify_batch(cs[d * tt:(d + 1) * tt], np.arange(nv, dtype=np.uint32), sh[d * nv:(d + 1) * nv])
        return out

    v = kb.sharding.dkg_round_sharded(round_fn, nver, t, commits, shares, rank, world)
    ok3 = bool((v == round_fn(nver, t, commits, shares)).all())
    q.put((rank, ok1, ok2, ok3))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res
