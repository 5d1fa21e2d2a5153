"""Parity of the protocol-level entry points (csrc/capi_proto.cu) against the oracle, through the C ABI, at BASELINE
config 3's shape (n = 256, t = 171) and at small / ragged sizes.  Bit-exact: encodings, digests, verdicts, statuses."""
import hashlib
import importlib

import numpy as np
import pytest

from helpers import load_sign_input, make_sig_batch, pack_batch
from oracle import ed25519_bigint as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kb():
    return importlib.import_module("kyber-rs_b200")


@pytest.fixture(scope="module")
def ctx(kb):
    c = kb.Context(0)
    yield c
    c.close()


def _scalars(tag: bytes, n: int) -> np.ndarray:
    """n scalars < L derived from a tag (SHA-512 mod L, the reference's Scalar::set_bytes)."""
    return np.frombuffer(b"".join(O.scalar_set_bytes(hashlib.sha512(tag + b"/%d" % k).digest()) for k in range(n)), dtype=np.uint8).reshape(n, 32).copy()


def _shares(coeffs: np.ndarray, n: int) -> np.ndarray:
    """PriPoly::eval (poly.rs:133) at indices 0..n-1 with Python integers."""
    c = [int.from_bytes(x.tobytes(), "little") for x in coeffs]
    out = np.zeros((n, 32), dtype=np.uint8)
    for i in range(n):
        v = 0
        for cj in reversed(c):
            v = (v * (i + 1) + cj) % O.L
        out[i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    return out


def _bad_point(coracle) -> bytes:
    k = 0
    while coracle.point_decode_ok(bytes([k]) + b"\x13" * 31):
        k += 1
    return bytes([k]) + b"\x13" * 31


NONCANON_ID = bytes([0xEE]) + b"\xff" * 30 + b"\x7f"    # y = p + 1 = 1: the identity, non-canonically encoded


@pytest.mark.parametrize("nd,n,t", [(1, 1, 1), (3, 5, 2), (7, 33, 9), (40, 256, 171)])
def test_session_ids(ctx, coracle, nd, n, t):
    """session_id (share/vss/pedersen/vss.rs:1069-1090): SHA-256 over canonical encodings, from 32-byte inputs and from
    the reference's raw limbs; a non-canonical input encoding hashes as its canonical twin, an undecodable one is flagged."""
    pts = ctx.point_mul_base_batch(_scalars(b"sid%d" % n, nd + n + nd * t), 1)
    dealers, verifiers, commits = pts[:nd].copy(), pts[nd:nd + n].copy(), pts[nd + n:].copy()
    want = [hashlib.sha256(dealers[d].tobytes() + verifiers.tobytes() + commits[d * t:(d + 1) * t].tobytes() + t.to_bytes(4, "little")).digest() for d in range(nd)]
    got, st = ctx.vss_session_ids(dealers, verifiers, commits, t)
    assert not st.any() and [g.tobytes() for g in got] == want
    assert want[0] == O.session_id(O.point_decode(dealers[0].tobytes()), [O.point_decode(v.tobytes()) for v in verifiers[:n]], [O.point_decode(c.tobytes()) for c in commits[:t]], t)
    limbs = [np.stack([coracle.point_limbs(p.tobytes()) for p in a]) for a in (dealers, verifiers, commits)]
    got_l, st_l = ctx.vss_session_ids(limbs[0], limbs[1], limbs[2], t, limbs=True)
    assert not st_l.any() and [g.tobytes() for g in got_l] == want
    if nd >= 3:
        commits2 = commits.copy()
        commits2[1 * t] = np.frombuffer(NONCANON_ID, dtype=np.uint8)
        commits2[2 * t + t - 1] = np.frombuffer(_bad_point(coracle), dtype=np.uint8)
        got2, st2 = ctx.vss_session_ids(dealers, verifiers, commits2, t)
        assert st2.tolist() == [0, 0, 1] + [0] * (nd - 3)
        canon = commits[1 * t:2 * t].copy()
        canon[0] = np.frombuffer((1).to_bytes(32, "little"), dtype=np.uint8)
        assert got2[1].tobytes() == hashlib.sha256(dealers[1].tobytes() + verifiers.tobytes() + canon.tobytes() + t.to_bytes(4, "little")).digest()
        assert got2[0].tobytes() == want[0]


def test_find_pub(ctx, coracle):
    """find_pub (share/dkg/pedersen/dkg.rs:1109-1116): first index under Point::eq, for 32-byte and raw-limb inputs."""
    n = 300
    lst = ctx.point_mul_base_batch(_scalars(b"findpub", n), 1)
    lst[17] = lst[5]                                           # a duplicate: the FIRST index wins
    lst[40] = np.frombuffer((1).to_bytes(32, "little"), dtype=np.uint8)
    others = ctx.point_mul_base_batch(_scalars(b"findpub-absent", 4), 1)
    q = np.concatenate([lst[[0, 299, 17, 5, 123]], others, np.frombuffer(NONCANON_ID + _bad_point(coracle), dtype=np.uint8).reshape(2, 32)])
    got = ctx.find_pub_batch(lst, q)
    assert got.tolist() == [0, 299, 5, 5, 123, -1, -1, -1, -1, 40, -2]
    ll = np.stack([coracle.point_limbs(p.tobytes()) for p in lst])
    ql = np.stack([coracle.point_limbs(p.tobytes()) for p in q[:10]])
    assert ctx.find_pub_batch(ll, ql, limbs=True).tolist() == got.tolist()[:10]
    assert ctx.find_pub_batch(lst[:0], q[:2]).tolist() == [-1, -1]


@pytest.mark.parametrize("n,t", [(6, 2), (33, 9), (256, 171)])
def test_rabin_verify_deals(ctx, coracle, n, t):
    """vss::rabin verify_deal (share/vss/rabin/vss.rs:889-900): f*G + g*H == eval(i) for every verifier of one dealer,
    a few shares corrupted; against the oracle's batch (the reference's t full scalar mults per check)."""
    fc, gc = _scalars(b"rabin-f%d" % t, t), _scalars(b"rabin-g%d" % t, t)
    H = ctx.point_mul_base_batch(_scalars(b"rabin-h", 1), 1)[0]
    fg = ctx.point_mul_base_batch(fc, 1)
    gh, st = ctx.point_mul_batch(gc, H.reshape(1, 32), 1)
    commits, st2 = ctx.point_add_batch(fg, gh)
    assert not st.any() and not st2.any()
    f, g = _shares(fc, n), _shares(gc, n)
    want = np.ones(n, dtype=np.uint8)
    for i in (0, n // 2, n - 1):
        (f if i % 2 else g)[i, 3] ^= 0x20
        want[i] = 0
    idx = np.arange(n, dtype=np.uint32)
    got = ctx.vss_rabin_verify_deals_batch(commits, t, H, np.zeros(n, dtype=np.uint32), idx, f, g)
    assert (got == want).all(), np.nonzero(got != want)[0]
    assert (coracle.rabin_verify_batch(commits, H.tobytes(), idx, f, g, nthreads=8) == got).all()
    # an undecodable H or commitment: nothing verifies
    assert not ctx.vss_rabin_verify_deals_batch(commits, t, np.frombuffer(_bad_point(coracle), dtype=np.uint8), np.zeros(n, dtype=np.uint32), idx, f, g).any()


@pytest.mark.parametrize("n,t", [(5, 2), (40, 13), (256, 171)])
def test_dss_verify_partials(ctx, coracle, n, t):
    """DSS::process_partial_sig (sign/dss/dss_sig.rs:244-277): hash_sig on the device, two evaluations, hash * long share,
    comparison with partial * B — plus the Schnorr check of each partial signature's own signature."""
    rc, lc = _scalars(b"dss-r%d" % t, t), _scalars(b"dss-l%d" % t, t)
    rcom, lcom = ctx.point_mul_base_batch(rc, 1), ctx.point_mul_base_batch(lc, 1)
    msg = b"message signed by the distributed key, n=%d" % n
    h = coracle.dss_hash_sig(rcom, lcom, msg)
    assert h == O.dss_hash_sig(O.point_decode(rcom[0].tobytes()), O.point_decode(lcom[0].tobytes()), msg)
    rs, ls = _shares(rc, n), _shares(lc, n)
    partials = np.frombuffer(b"".join(O.sc_add(rs[i].tobytes(), O.sc_mul(h, ls[i].tobytes())) for i in range(n)), dtype=np.uint8).reshape(n, 32).copy()
    want = np.ones(n, dtype=np.uint8)
    for i in (1, n - 1):
        partials[i, 7] ^= 1
        want[i] = 0
    idx = np.arange(n, dtype=np.uint32)
    got, hs = ctx.dss_verify_partials(rcom, lcom, msg, idx, partials)
    assert hs == h
    assert (got == want).all(), np.nonzero(got != want)[0]
    assert (coracle.dss_partial_batch(rcom, lcom, msg, idx, partials, nthreads=8) == got).all()
    # out-of-order and repeated indices, empty message
    sel = np.array([n - 1, 0, 0, 2], dtype=np.uint32)
    got2, hs2 = ctx.dss_verify_partials(rcom, lcom, b"", sel, partials[sel])
    assert hs2 == coracle.dss_hash_sig(rcom, lcom, b"")
    assert (got2 == coracle.dss_partial_batch(rcom, lcom, b"", sel, partials[sel], nthreads=2)).all()


@pytest.mark.parametrize("k,ncols", [(1, 1), (2, 3), (6, 1), (35, 8), (171, 171)])
def test_recover_commit_and_resharing(ctx, coracle, k, ncols):
    """recover_commit (share/poly.rs:566-603) column by column and resharing_key's use of it (dkg.rs:996-1031):
    interpolating the public shares of ncols polynomials gives the commitments of their secrets; one column is also
    checked against the oracle's restatement of the reference's loop."""
    n = k + 5
    rng = np.random.default_rng(k)
    pick = np.sort(rng.choice(n, size=k, replace=False)).astype(np.uint32)
    secrets = _scalars(b"recover%d" % k, ncols)
    cols, want = [], ctx.point_mul_base_batch(secrets, 1)
    for c in range(ncols):
        coeffs = np.concatenate([secrets[c:c + 1], _scalars(b"recover%d/%d" % (k, c), k - 1)]) if k > 1 else secrets[c:c + 1]
        sh = _shares(coeffs, n)[pick] if (c < 3 or k <= 35) else None
        if sh is None:   # large case: shares of the other columns from a cheaper, equally valid polynomial (degree 1)
            sh = _shares(np.concatenate([secrets[c:c + 1], _scalars(b"lin%d" % c, 1)]), n)[pick]
        cols.append(ctx.point_mul_base_batch(sh, 1))
    pts = np.concatenate(cols)
    got, st = ctx.recover_commit_batch(pick, pts, ncols=ncols)
    assert not st.any() and (got == want).all()
    assert coracle.recover_commit(pick, cols[0]) == want[0].tobytes()
    # resharing_key takes the same data node-major: coeffs[i][c]
    node_major = pts.reshape(ncols, k, 32).transpose(1, 0, 2).reshape(-1, 32)
    new_poly_share = None
    out, st2, chk = ctx.dkg_resharing_key(ncols, pick, node_major)
    assert not st2.any() and (out == want).all() and chk is None
    # the final check pub_poly.check(share): the recovered commitments are those of the polynomial `secrets`
    share = _shares(secrets, 4)[3]
    _, _, chk = ctx.dkg_resharing_key(ncols, pick, node_major, share_idx=3, share=share)
    assert chk is True
    share[0] ^= 1
    _, _, chk = ctx.dkg_resharing_key(ncols, pick, node_major, share_idx=3, share=share)
    assert chk is False
    if k >= 2:   # an undecodable share point flags its column only
        bad = pts.copy()
        bad[0 * k + 1] = np.frombuffer(_bad_point(coracle), dtype=np.uint8)
        _, st3 = ctx.recover_commit_batch(pick, bad, ncols=ncols)
        assert st3.tolist() == [1] + [0] * (ncols - 1)


@pytest.mark.parametrize("k", [1, 2, 7, 40, 171])
def test_recover_pub_poly(ctx, coracle, k):
    """recover_pub_poly (share/poly.rs:607-635): the commitments of the polynomial through k public shares; the oracle
    restates the reference's lagrange_basis / commit / add loop (cubic: compared up to k = 40), beyond that the
    known answer is the commitment vector itself."""
    coeffs = _scalars(b"pubpoly%d" % k, k)
    commits = ctx.point_mul_base_batch(coeffs, 1)
    n = k + 3
    for pick in (np.arange(k, dtype=np.uint32), np.sort(np.random.default_rng(k).choice(n, size=k, replace=False)).astype(np.uint32)):
        pub = ctx.point_mul_base_batch(_shares(coeffs, n)[pick], 1)
        got, st = ctx.recover_pub_poly(pick, pub)
        assert not st.any() and (got == commits).all()
        if k <= 40:
            assert (coracle.recover_pub_poly(pick, pub) == got).all()


def test_dkg_process_round(ctx, coracle, golden_records):
    """One deal-verification round as a whole (share/dkg/pedersen/dkg.rs:513-597, share/vss/pedersen/vss.rs:931-946):
    share checks + Schnorr verification of deal and response signatures, each held to its own oracle, for a dealer
    sub-range and for commitments given as raw limbs."""
    n, t, nd = 24, 9, 24
    polys = [_scalars(b"round/%d" % d, t) for d in range(nd)]
    commits = ctx.point_mul_base_batch(np.concatenate(polys), 1)
    shares = np.concatenate([_shares(p, n) for p in polys])
    want = np.ones(nd * n, dtype=np.uint8)
    for k in (0, 100, nd * n - 1):
        shares[k, 1] ^= 2
        want[k] = 0
    m = nd * n
    pks, msgs, sigs = make_sig_batch(golden_records[:200], 2 * m, bad_every=5)
    deal = pack_batch(pks[:m], msgs[:m], sigs[:m])
    resp = pack_batch(pks[m:], msgs[m:], sigs[m:])
    v, ds, rs = ctx.dkg_process_round(n, t, commits, shares, deal=deal, resp=resp)
    assert (v == want).all()
    assert (ds == coracle.verify_batch(*deal, nthreads=8, schnorr=True)).all() and set(ds.tolist()) >= {0, 8}
    assert (rs == coracle.verify_batch(*resp, nthreads=8, schnorr=True)).all()
    # a rank's dealer range: its verdict rows, and signature arrays that hold just its items
    lo, hi = 8, 16
    sub = lambda b: (b[0][lo * n:hi * n], b[1][int(b[2][lo * n]):int(b[2][hi * n])], b[2][lo * n:hi * n + 1] - b[2][lo * n], b[3][lo * n:hi * n])
    v2, ds2, rs2 = ctx.dkg_process_round(n, t, commits, shares, deal=sub(deal), resp=None, dealer_lo=lo, dealer_hi=hi)
    assert (v2[lo * n:hi * n] == want[lo * n:hi * n]).all() and (ds2 == ds[lo * n:hi * n]).all() and rs2 is None
    limbs = np.stack([coracle.point_limbs(c.tobytes()) for c in commits])
    v3, _, _ = ctx.dkg_process_round(n, t, limbs, shares, limbs=True)
    assert (v3 == want).all()


def test_dkg_process_round_device_resident(kb, ctx, coracle, golden_records):
    """kb_dev_dkg_process_round on CUDA tensors gives what the host-buffer call gives."""
    import torch

    n, t, nd = 16, 5, 32
    polys = [_scalars(b"devround/%d" % d, t) for d in range(nd)]
    commits = ctx.point_mul_base_batch(np.concatenate(polys), 1)
    shares = np.concatenate([_shares(p, n) for p in polys])
    shares[7, 0] ^= 1
    m = nd * n
    pks, msgs, sigs = make_sig_batch(golden_records[:100], m, bad_every=7)
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    hv, hd, _ = ctx.dkg_process_round(n, t, commits, shares, deal=(pk, flat, off, sg))
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_v = torch.zeros(m, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(m, dtype=torch.uint8, device=dev)
    ctx.dev_dkg_process_round(n, t, nd, up(commits), up(shares), d_v, deal=(up(pk), up(flat), up(off.view(np.int64)), up(sg), d_st))
    torch.cuda.synchronize()
    assert (d_v.cpu().numpy() == hv).all() and (d_st.cpu().numpy() == hd).all()


def test_pripoly_eval_batch(ctx):
    """PriPoly::eval (share/poly.rs:133-141) for whole polynomials at once: the shares a dealer hands out, against
    Python integers; coefficients that are not reduced (a raw 32-byte scalar) are taken mod L like sc_mul_add does."""
    for npoly, t, n in ((1, 1, 1), (3, 2, 5), (5, 171, 40), (2, 683, 9)):
        coeffs = _scalars(b"pripoly%d" % t, npoly * t)
        if t > 1:
            coeffs[1] = 0xFF          # 2^256 - 1: an unreduced scalar
        got = ctx.pripoly_eval_batch(coeffs, t, n)
        for d in range(npoly):
            assert (got[d * n:(d + 1) * n] == _shares(coeffs[d * t:(d + 1) * t], n)).all(), (npoly, t, n, d)
