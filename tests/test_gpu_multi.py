"""The multi-device context (csrc/capi_multi.cu) against the oracle and against the single-device context, through the
C ABI.  Runs on however many GPUs are visible: with one GPU it exercises the sharding plumbing on a single shard, with
two or more the index / dealer / point partitioning and the NCCL gather of the MSM partials."""
import hashlib
import importlib

import numpy as np
import pytest

from helpers import make_sig_batch, pack_batch
from oracle import ed25519_bigint as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kb():
    return importlib.import_module("kyber-rs_b200")


@pytest.fixture(scope="module")
def ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.fixture(scope="module", params=["one", "all"])
def mctx(kb, ngpu, request):
    devs = [0] if request.param == "one" else list(range(min(ngpu, 8)))
    if request.param == "all" and ngpu < 2:
        pytest.skip("a single GPU is visible")
    m = kb.MultiContext(devs)
    yield m
    m.close()


def _scalars(tag: bytes, n: int) -> np.ndarray:
    return np.frombuffer(b"".join(O.scalar_set_bytes(hashlib.sha512(tag + b"/%d" % k).digest()) for k in range(n)), dtype=np.uint8).reshape(n, 32).copy()


def test_mctx_verify_and_mul(kb, mctx, coracle, golden_records):
    """Signature batches and scalar multiplications split by index: every status / encoding as the oracle has it, for a
    batch that does not divide evenly and one smaller than the device count."""
    for n in (1003, 1):
        pks, msgs, sigs = make_sig_batch(golden_records[:300], n, bad_every=4)
        pk, flat, off, sg = pack_batch(pks, msgs, sigs)
        for schnorr in (False, True):
            got = mctx.verify_batch(pk, flat, off, sg, schnorr=schnorr)
            assert (got == coracle.verify_batch(pk, flat, off, sg, nthreads=8, schnorr=schnorr)).all()
    s = _scalars(b"mctx-mul", 777)
    base = mctx.point_mul_base_batch(s)
    assert (base == coracle.mul_base_batch(s, nthreads=8)).all()
    out, st = mctx.point_mul_batch(s, base[::-1].copy())
    assert not st.any() and (out == coracle.mul_batch(s, base[::-1].copy(), nthreads=8)).all()
    out1, st1 = mctx.point_mul_batch(s, base[:1].copy())     # one shared point
    assert not st1.any() and (out1 == coracle.mul_batch(s, np.repeat(base[:1], 777, axis=0), nthreads=8)).all()


def test_mctx_msm(kb, mctx, coracle):
    """MSM split by points with the partials gathered over NCCL: equals the oracle's fold at 2^12 points, the
    single-device result at 2^17, and counts undecodable inputs across shards."""
    n = 1 << 12
    s = _scalars(b"mctx-msm", n)
    pts = mctx.point_mul_base_batch(_scalars(b"mctx-msm-pts", n), 1)
    enc, bad = mctx.msm(s, pts)
    assert bad == 0 and enc == coracle.msm(s, pts)
    assert mctx.msm(s[:0], pts[:0]) == ((1).to_bytes(32, "little"), 0)
    assert mctx.msm(s[:3], pts[:3]) == (coracle.msm(s[:3], pts[:3]), 0)
    big = 1 << 17
    sb = np.tile(s, (big // n, 1))
    pb = np.tile(pts, (big // n, 1))
    one = kb.Context(0)
    assert mctx.msm(sb, pb)[0] == one.msm(sb, pb)[0]
    one.close()
    pts2 = pts.copy()
    k = 0
    while coracle.point_decode_ok(bytes([k]) + b"\x13" * 31):
        k += 1
    pts2[5] = pts2[n - 2] = np.frombuffer(bytes([k]) + b"\x13" * 31, dtype=np.uint8)
    assert mctx.msm(s, pts2)[1] == 2


def test_mctx_dkg_round(kb, mctx, coracle, golden_records):
    """A DKG round split by dealer (a dealer count that does not divide evenly): verdict rows and the statuses of the
    deal / response signatures equal the single-device results, which equal the oracle's (tests/test_gpu_proto.py)."""
    n, t, nd = 20, 7, 37
    polys = [_scalars(b"mround/%d" % d, t) for d in range(nd)]
    commits = mctx.point_mul_base_batch(np.concatenate(polys), 1)
    shares = np.zeros((nd * n, 32), dtype=np.uint8)
    for d in range(nd):
        c = [int.from_bytes(x.tobytes(), "little") for x in polys[d]]
        for i in range(n):
            v = 0
            for cj in reversed(c):
                v = (v * (i + 1) + cj) % O.L
            shares[d * n + i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    want = np.ones(nd * n, dtype=np.uint8)
    for k in (0, 333, nd * n - 1):
        shares[k, 2] ^= 1
        want[k] = 0
    m = nd * n
    pks, msgs, sigs = make_sig_batch(golden_records[:200], 2 * m, bad_every=6)
    deal, resp = pack_batch(pks[:m], msgs[:m], sigs[:m]), pack_batch(pks[m:], msgs[m:], sigs[m:])
    v, ds, rs = mctx.dkg_process_round(n, t, commits, shares, deal=deal, resp=resp)
    assert (v == want).all()
    assert (ds == coracle.verify_batch(*deal, nthreads=8, schnorr=True)).all()
    assert (rs == coracle.verify_batch(*resp, nthreads=8, schnorr=True)).all()
    assert (mctx.dkg_verify_round(n, t, commits, shares) == want).all()
    row = coracle.vss_verify_batch(commits[36 * t:37 * t], np.arange(n, dtype=np.uint32), shares[36 * n:37 * n], nthreads=4)
    assert (row == v[36 * n:37 * n]).all()
