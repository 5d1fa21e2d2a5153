"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against the oracle on
the same inputs.  Bit-exact: compressed encodings and accept/reject statuses.  Needs a B200."""
import hashlib
import importlib

import numpy as np
import pytest

import os

from helpers import load_sign_input, make_mixed_order_sigs, make_sig_batch, pack_batch, random_scalars, xof_bytes
from oracle import ed25519_bigint as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kb():
    return importlib.import_module("kyber-rs_b200")


@pytest.fixture(scope="module")
def ctx(kb):
    c = kb.Context(0)
    kb.host.set_default_context(c)
    yield c


@pytest.fixture(scope="module")
def ctx_full(kb):
    """A context that runs the full-length (253-doubling, two-launch) verify kernels instead of the default
    half-size-scalar kernel (csrc/half.cuh): both must give the reference's statuses."""
    os.environ["KB_VERIFY_FULL"] = "1"
    try:
        c = kb.Context(0)
    finally:
        del os.environ["KB_VERIFY_FULL"]
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_long(kb):
    """The half-size-scalar kernel forced to run 64 windows in every block — what an adversarial challenge whose
    lattice has no short vector with an odd u would make it do (never a hash output)."""
    os.environ["KB_VERIFY_MIN_WINDOWS"] = "64"
    try:
        c = kb.Context(0)
    finally:
        del os.environ["KB_VERIFY_MIN_WINDOWS"]
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_split(kb):
    """The preparation of the half-size-scalar verifiers as two kernels side by side (a persistent "scalars" kernel on a
    side stream beside the "points" grid, joined by k_verify_half_fix), forced for every batch size."""
    os.environ["KB_VERIFY_SPLIT"] = "2"
    os.environ["KB_VERIFY_SPLIT_BLOCKS"] = "2"
    os.environ["KB_VERIFY_SORT"] = "2"   # and the records sorted by loop length for every batch size (default: from 16384 on)
    try:
        c = kb.Context(0)
    finally:
        del os.environ["KB_VERIFY_SPLIT"]
        del os.environ["KB_VERIFY_SPLIT_BLOCKS"]
        del os.environ["KB_VERIFY_SORT"]
    yield c
    c.close()


def _golden_pks(records, n):
    return np.frombuffer(b"".join(r[1] for r in records[:n]), dtype=np.uint8).reshape(-1, 32).copy()


# ---- Point::mul ---------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, 1])
def test_mul_base_matches_oracle(ctx, coracle, flags):
    """config 1, base-point half: Point::mul(s, None) (point.rs:207 -> ge.rs:442)."""
    n = 4096
    s = random_scalars("kyber-b200/cfg1/scalars", n)
    s[:64] = np.frombuffer(xof_bytes("raw", 64 * 32), dtype=np.uint8).reshape(64, 32)   # unreduced / out-of-domain
    s[64] = 0
    s[65] = 255
    got = ctx.point_mul_base_batch(s, flags)
    want = coracle.mul_base_batch(s, nthreads=8)
    assert (got == want).all()


def test_mul_base_golden_public_keys(ctx, golden_records):
    """pk = clamp(SHA512(seed)) * B for all 1024 golden lines (tests/sign/eddsa.rs:37-94)."""
    sc = np.frombuffer(b"".join(O.clamp_key(r[0])[0] for r in golden_records), dtype=np.uint8).reshape(-1, 32)
    got = ctx.point_mul_base_batch(sc)
    assert got.tobytes() == b"".join(r[1] for r in golden_records)


@pytest.mark.parametrize("flags", [0, 1])
def test_mul_var_base_matches_oracle(ctx, coracle, golden_records, flags):
    """config 1, variable-base half: Point::mul(s, Some(p)) (ge.rs:508), distinct and shared point."""
    n = 1024
    s = random_scalars("kyber-b200/cfg1/scalars2", n)
    s[:32] = np.frombuffer(xof_bytes("raw2", 32 * 32), dtype=np.uint8).reshape(32, 32)
    pts = _golden_pks(golden_records, n)
    pts[7] = np.frombuffer(O.WEAK_KEYS[2], dtype=np.uint8)          # torsion point
    pts[8] = np.frombuffer(bytes([0xEE]) + b"\xff" * 30 + b"\x7f", dtype=np.uint8)  # y >= p is accepted by decode
    got, st = ctx.point_mul_batch(s, pts, flags)
    want = coracle.mul_batch(s, pts, nthreads=8)
    assert (got == want).all()
    ok = np.array([coracle.point_decode_ok(p.tobytes()) for p in pts])
    assert (st == (~ok).astype(np.uint8)).all()
    got1, _ = ctx.point_mul_batch(s[:256], pts[3:4], flags)
    want1 = coracle.mul_batch(s[:256], np.repeat(pts[3:4], 256, axis=0), nthreads=8)
    assert (got1 == want1).all()


def test_undecodable_point_is_flagged(ctx, coracle):
    bad = None
    k = 0
    while bad is None:
        c = hashlib.sha256(b"bad%d" % k).digest()
        if not coracle.point_decode_ok(c):
            bad = c
        k += 1
    out, st = ctx.point_mul_batch(np.ones((1, 32), dtype=np.uint8), np.frombuffer(bad, dtype=np.uint8))
    assert st[0] == 1 and not out.any()


# ---- encodings, add, checks -----------------------------------------------------------------------
def test_recode_add_check(ctx, coracle, golden_records):
    rnd = np.frombuffer(xof_bytes("recode", 2048 * 32), dtype=np.uint8).reshape(-1, 32).copy()
    rnd[:5] = np.frombuffer(b"".join(O.WEAK_KEYS), dtype=np.uint8).reshape(5, 32)
    for b0 in range(256):
        rnd[16 + b0] = np.frombuffer(bytes([b0]) + b"\xff" * 30 + b"\x7f", dtype=np.uint8)
    out, st = ctx.point_recode_batch(rnd)
    flags = ctx.point_check_batch(rnd)
    for i, r in enumerate(rnd):
        want = coracle.point_recode(r.tobytes())
        assert (st[i] == 0) == (want is not None)
        if want is not None:
            assert out[i].tobytes() == want
        assert bool(flags[i] & 1) == O.point_is_canonical(r.tobytes())
        assert bool(flags[i] & 4) == (want is not None)
        if want is not None:
            assert bool(flags[i] & 2) == bool(coracle.point_has_small_order(r.tobytes()))
    p = _golden_pks(golden_records, 512)
    q = np.roll(p, 1, axis=0)
    for sub in (False, True):
        got, st = ctx.point_add_batch(p, q, subtract=sub)
        assert not st.any()
        for i in range(0, 512, 7):
            assert got[i].tobytes() == coracle.point_add(p[i].tobytes(), q[i].tobytes(), sub)
    same, _ = ctx.point_add_batch(p[:8], p[:8])                       # doubling through the unified law
    assert same[0].tobytes() == coracle.point_add(p[0].tobytes(), p[0].tobytes())
    zero, _ = ctx.point_add_batch(p[:8], p[:8], subtract=True)
    assert zero[0].tobytes() == (1).to_bytes(32, "little")


# ---- scalars and the challenge hash ---------------------------------------------------------------
def test_decompress_compress_eq(ctx, coracle, golden_records):
    """The uncompressed chaining form: decompress -> compress is the reference's unmarshal -> marshal (re-encoding,
    so y >= p comes back reduced); Point::eq (point.rs:227-241) compares re-encodings."""
    rnd = np.random.default_rng(21)
    pts = [r[1] for r in golden_records[:300]] + list(O.WEAK_KEYS)
    pts += [bytes([0xEE]) + b"\xff" * 30 + b"\x7f", bytes([0xF0]) + b"\xff" * 30 + b"\xff", bytes([1]) + bytes(30) + b"\x80"]   # y >= p twice, identity with sign bit
    pts += [rnd.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(200)]
    arr = np.frombuffer(b"".join(pts), dtype=np.uint8).reshape(-1, 32)
    raw, st = ctx.point_decompress_batch(arr)
    enc = ctx.point_compress_batch(raw)
    for k, p in enumerate(pts):
        want = coracle.point_recode(p)
        assert (st[k] == 0) == (want is not None)
        if want is not None:
            assert enc[k].tobytes() == want
        else:
            assert enc[k].tobytes() == O.point_encode(O.IDENTITY)
    # Z = 0 (Point::default()) encodes as zeros; a scaled representative compresses to the same bytes
    z = raw[:4].copy()
    z[0, 16:24] = 0
    for c in range(4):   # multiply X, Y, Z, T of item 1 by 2 (still < 2^256 for these small limbs? use doubling via add mod p on host)
        v = int.from_bytes(z[1, 8 * c:8 * c + 8].tobytes(), "little") * 2 % O.P
        z[1, 8 * c:8 * c + 8] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint32)
    ez = ctx.point_compress_batch(z)
    assert ez[0].tobytes() == bytes(32) and ez[1].tobytes() == enc[1].tobytes() and ez[2].tobytes() == enc[2].tobytes()
    # eq: same point through a different encoding, different points, undecodable operand
    a = arr.copy()
    b = arr.copy()
    b[1] = arr[2]
    noncanon = np.frombuffer(bytes([0xEE]) + b"\xff" * 30 + b"\x7f", dtype=np.uint8)   # y = p + 1 -> the point y = 1
    a[3] = noncanon
    b[3] = np.frombuffer(O.point_encode(O.IDENTITY), dtype=np.uint8)
    got = ctx.point_eq_batch(a, b)
    for k in range(len(pts)):
        pa, pb = coracle.point_recode(a[k].tobytes()), coracle.point_recode(b[k].tobytes())
        want = (2 if (pa is None or pb is None) else 0) | (1 if (pa is not None and pa == pb) else 0)
        assert got[k] == want, k
    assert got[1] == 0 and got[3] == 1


def test_scalar_ops_and_challenge(ctx, coracle, golden_records):
    d = np.frombuffer(xof_bytes("digests", 1000 * 64), dtype=np.uint8).reshape(-1, 64).copy()
    d[0] = 255
    d[1] = 0
    got = ctx.sc_reduce64_batch(d)
    for i in range(len(d)):
        assert got[i].tobytes() == O.scalar_set_bytes(d[i].tobytes())
    abc = np.frombuffer(xof_bytes("abc", 3 * 500 * 32), dtype=np.uint8).reshape(3, -1, 32)
    got = ctx.sc_muladd_batch(abc[0], abc[1], abc[2])
    for i in range(500):
        assert got[i].tobytes() == O.sc_mul_add(abc[0][i].tobytes(), abc[1][i].tobytes(), abc[2][i].tobytes())
    recs = golden_records[::4]            # message lengths 0..1020: every SHA-512 padding case
    pk, flat, off, sg = pack_batch([r[1] for r in recs], [r[3] for r in recs], [r[2] for r in recs])
    h = ctx.challenge_batch(sg[:, :32], pk, flat, off)
    for i, r in enumerate(recs):
        assert h[i].tobytes() == O.challenge(r[2][:32], r[1], r[3])


# ---- signatures -------------------------------------------------------------------------------------
@pytest.mark.parametrize("schnorr", [False, True])
def test_verify_golden_file(ctx, golden_records, schnorr):
    """All 1024 golden signatures verify (tests/sign/eddsa.rs:37-94; schnorr_test.rs:43-51 says the
    two verifiers accept each other's signatures)."""
    pk, flat, off, sg = pack_batch([r[1] for r in golden_records], [r[3] for r in golden_records], [r[2] for r in golden_records])
    st = ctx.verify_batch(pk, flat, off, sg, schnorr=schnorr)
    assert not st.any()


@pytest.mark.parametrize("path", ["half", "full"])
@pytest.mark.parametrize("schnorr", [False, True])
def test_verify_reject_classes(ctx, ctx_full, coracle, golden_records, schnorr, path):
    """Every mutation class (eddsa_test.rs:111-272, schnorr_test.rs:6-110 and SURVEY §A cases):
    status must equal the oracle's, i.e. the reference's error variant in ITS check order."""
    c = ctx if path == "half" else ctx_full
    pks, msgs, sigs = make_sig_batch(golden_records, 4096, bad_every=2)
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    got = c.verify_batch(pk, flat, off, sg, schnorr=schnorr)
    want = coracle.verify_batch(pk, flat, off, sg, nthreads=8, schnorr=schnorr)
    assert (got == want).all(), np.nonzero(got != want)[0][:10]
    assert set(np.unique(want)) == {0, 2, 3, 4, 5, 6, 7, 8}


@pytest.mark.parametrize("schnorr", [False, True])
def test_verify_mixed_order_keys(ctx, ctx_full, coracle, golden_records, schnorr):
    """Keys and R values that carry a small-order component (they pass the reference's small-order filter): the
    reference's cofactorless equation accepts exactly those with T' + h*T = 0.  Both kernels must agree with
    the oracle on the accepted ones AND on the near misses (8*defect = 0), which is where a shortened-scalar
    verifier that reduced modulo L instead of 8L would go wrong."""
    good, bad = make_mixed_order_sigs(96, seed=5)
    items = []
    for k in range(96):
        items += [good[k], bad[k], (golden_records[k][1], golden_records[k][3], golden_records[k][2])]
    pk, flat, off, sg = pack_batch([x[0] for x in items], [x[1] for x in items], [x[2] for x in items])
    want = coracle.verify_batch(pk, flat, off, sg, nthreads=8, schnorr=schnorr)
    assert (want.reshape(-1, 3) == np.array([0, 8, 0])).all()
    for c in (ctx, ctx_full):
        got = c.verify_batch(pk, flat, off, sg, schnorr=schnorr)
        assert (got == want).all(), np.nonzero(got != want)[0][:10]


@pytest.mark.parametrize("schnorr", [False, True])
def test_verify_long_window_loops(ctx_long, coracle, golden_records, schnorr):
    """Same verdicts when every block runs the maximum number of windows (block-uniform trip count above what its
    own signatures need: the extra windows only add identities)."""
    good, bad = make_mixed_order_sigs(16, seed=11)
    pks, msgs, sigs = make_sig_batch(golden_records[300:], 700, bad_every=2)
    pks += [x[0] for x in good + bad]
    msgs += [x[1] for x in good + bad]
    sigs += [x[2] for x in good + bad]
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    got = ctx_long.verify_batch(pk, flat, off, sg, schnorr=schnorr)
    want = coracle.verify_batch(pk, flat, off, sg, nthreads=8, schnorr=schnorr)
    assert (got == want).all(), np.nonzero(got != want)[0][:10]


@pytest.mark.parametrize("schnorr", [False, True])
def test_verify_split_preparation(ctx, ctx_split, coracle, golden_records, schnorr):
    """The two-kernel preparation (KB_VERIFY_SPLIT) writes the same records: every mutation class and the mixed-order
    keys get the oracle's statuses, device-resident and through the pipelined host-buffer call."""
    good, bad = make_mixed_order_sigs(32, seed=7)
    pks, msgs, sigs = make_sig_batch(golden_records, 4096, bad_every=2)
    pks += [x[0] for x in good + bad]
    msgs += [x[1] for x in good + bad]
    sigs += [x[2] for x in good + bad]
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    want = coracle.verify_batch(pk, flat, off, sg, nthreads=8, schnorr=schnorr)
    got = ctx_split.verify_batch(pk, flat, off, sg, schnorr=schnorr)
    assert (got == want).all(), np.nonzero(got != want)[0][:10]
    for m in (1, 2, 127, 129, 1000):   # ragged sizes through the sorted order
        assert (ctx_split.verify_batch(pk[:m], flat[: int(off[m])], off[: m + 1], sg[:m], schnorr=schnorr) == want[:m]).all()
    assert (ctx.verify_batch(pk, flat, off, sg, schnorr=schnorr) == want).all()
    assert set(np.unique(want)) == {0, 2, 3, 4, 5, 6, 7, 8}


def test_verify_paths_agree_on_random_batch(kb, ctx, ctx_full, coracle):
    """2^16 fresh signatures (signed on the GPU), 1/8 damaged in s, R, A or the message: the half-size-scalar
    kernel, the full-length kernels and (on a slice) the oracle give the same statuses."""
    n = 1 << 16
    seeds = np.frombuffer(xof_bytes("kyber-b200/test/half/seed", 32 * n), dtype=np.uint8).reshape(n, 32).copy()
    flat = np.frombuffer(xof_bytes("kyber-b200/test/half/msg", 64 * n), dtype=np.uint8).copy()
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(64))
    sigs, pks = ctx.eddsa_sign_batch(seeds, flat, off)
    sigs, pks, flat = sigs.copy(), pks.copy(), flat.copy()
    rng = np.random.default_rng(9)
    for i in range(0, n, 8):
        kind = (i // 8) % 4
        if kind == 0:
            sigs[i, 32 + rng.integers(0, 31)] ^= 1 << rng.integers(0, 8)
        elif kind == 1:
            sigs[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
        elif kind == 2:
            pks[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
        else:
            flat[64 * i + rng.integers(0, 64)] ^= 1
    a = ctx.verify_batch(pks, flat, off, sigs)
    b = ctx_full.verify_batch(pks, flat, off, sigs)
    assert (a == b).all(), np.nonzero(a != b)[0][:10]
    assert (a.reshape(-1, 8)[:, 1:] == 0).all() and a.reshape(-1, 8)[:, 0].all()
    m = 4096
    want = coracle.verify_batch(pks[:m], flat[:64 * m], off[:m + 1], sigs[:m], nthreads=8)
    assert (a[:m] == want).all()


@pytest.mark.parametrize("schnorr", [False, True])
def test_device_resident_verify_and_kernel_timing(ctx, ctx_full, coracle, golden_records, schnorr):
    """kb_dev_eddsa_verify on torch tensors (the path bench.py times) gives the oracle's statuses on both kernel
    families, for batch sizes around the block / warp boundaries, and kb_verify_kernel_times reports both launches."""
    import torch

    dev = torch.device("cuda", 0)
    for n in (1, 31, 64, 129, 1000):
        pks, msgs, sigs = make_sig_batch(golden_records[5:], n, bad_every=3)
        pk, flat, off, sg = pack_batch(pks, msgs, sigs)
        want = coracle.verify_batch(pk, flat, off, sg, nthreads=4, schnorr=schnorr)
        d_pk, d_sig = torch.from_numpy(pk).to(dev), torch.from_numpy(sg).to(dev)
        d_msg = torch.from_numpy(flat if flat.size else np.zeros(1, dtype=np.uint8)).to(dev)
        d_off = torch.from_numpy(off.view(np.int64)).to(dev)
        for c in (ctx, ctx_full):
            d_st = torch.full((n,), 99, dtype=torch.uint8, device=dev)
            c.verify_kernel_timing(True)
            c.dev_verify(n, d_pk, d_msg, d_off, d_sig, d_st, schnorr=schnorr)
            a_ms, b_ms = c.last_verify_kernel_ms()
            c.verify_kernel_timing(False)
            assert a_ms > 0 and b_ms > 0
            assert (d_st.cpu().numpy() == want).all()


def test_c_abi_from_plain_c(kb, coracle, golden_records, tmp_path):
    """tests/c/abi_check.c: the boundary called from plain C (gcc, -lkyber_b200) — no Python, no torch in the
    process — on a fixture whose expected bytes come from the oracle."""
    import subprocess

    from helpers import ROOT

    n = 1500
    pks, msgs, sigs = make_sig_batch(golden_records, n, bad_every=4)
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    want = coracle.verify_batch(pk, flat, off, sg, nthreads=8)
    scalars = np.frombuffer(xof_bytes("kyber-b200/test/c-abi", 32 * n), dtype=np.uint8).reshape(n, 32).copy()
    want_mul = coracle.mul_base_batch(scalars)
    fx = tmp_path / "fixture.bin"
    with open(fx, "wb") as f:
        f.write(np.array([n, flat.size], dtype=np.uint64).tobytes())
        for a in (pk, sg, off, flat, want.astype(np.uint8), scalars, want_mul):
            f.write(np.ascontiguousarray(a).tobytes())
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(kb.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_check.c"),
                           "-L", libdir, "-lkyber_b200", "-Wl,-rpath," + libdir, "-o", str(exe)])
    r = subprocess.run([str(exe), str(fx)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches=0" in r.stdout


def test_verify_host_mirror_errors(kb, ctx, golden_records):
    """Reads like sign/eddsa/eddsa_test.rs: error strings are the reference's."""
    H = kb.host
    _, pk, sig, msg = golden_records[33]
    public = H.Point.unmarshal_binary(pk)
    H.eddsa_verify(public, msg, sig)
    H.schnorr_verify(public, msg, sig)
    s_plus_l = ((int.from_bytes(sig[32:], "little") + O.L) % 2**256).to_bytes(32, "little")
    cases = [
        (pk, msg, sig[:32] + s_plus_l, "signature is not canonical"),
        (pk, msg, bytes([0xEF]) + b"\xff" * 31 + sig[32:], "R is not canonical"),
        (bytes([0xEF]) + b"\xff" * 31, msg, sig, "public key is not canonical"),
        (pk, msg, O.WEAK_KEYS[3] + sig[32:], "R has small order"),
        (O.WEAK_KEYS[3], msg, sig, "public key has small order"),
        (pk, msg + b"!", sig, "signature is not valid"),
        (pk, msg, sig[:63], "wrong signature length"),
    ]
    for p, m, s, text in cases:
        with pytest.raises(H.SignatureError) as e:
            H.eddsa_verify_with_checks(p, m, s)
        assert str(e.value) == text
    with pytest.raises(H.MarshallingError):
        H.Point.unmarshal_binary(hashlib.sha256(b"bad0").digest() if not ctx.point_check_batch(np.frombuffer(hashlib.sha256(b"bad0").digest(), np.uint8))[0] & 4 else b"\x02" + bytes(31))


def test_bad_message_offsets_are_refused(kb, ctx, golden_records):
    """A non-monotone offset array is an argument error (KB_ERR_ARG), not an out-of-bounds read on the device — also
    when the bad entry sits in the middle of a pipelined chunk."""
    pks, msgs, sigs = make_sig_batch(golden_records[:64], 5000, bad_every=0)
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    assert not ctx.verify_batch(pk, flat, off, sg).any()
    bad = off.copy()
    bad[3333] = bad[3334] + 7
    for fn in (lambda: ctx.verify_batch(pk, flat, bad, sg), lambda: ctx.verify_batch(pk, flat, bad, sg, schnorr=True), lambda: ctx.challenge_batch(sg[:, :32], pk, flat, bad),
               lambda: ctx.eddsa_sign_batch(pk, flat, bad)):
        with pytest.raises(kb.KBError):
            fn()
    assert not ctx.verify_batch(pk, flat, off, sg).any()      # the context is still usable


@pytest.mark.parametrize("pipe,chunk", [(0, None), (1, None), (0, 1024), (1, 1024), (1, 4999), (0, 30000)])
@pytest.mark.parametrize("schnorr", [False, True])
def test_host_verify_pipeline_schedules(kb, ctx, coracle, golden_records, schnorr, pipe, chunk):
    """The host-buffer verify calls cut a batch into chunks (a third of a wave, then growing) and run them on two
    alternating lanes or, with KB_VERIFY_PIPE=1, with all kernels on one stream and the copies on a second one.  Every
    schedule and chunk cap — including caps that leave a ragged last chunk and dozens of chunks — must return the
    statuses of the single-launch device path, and a bad offset in a late chunk must be an argument error."""
    env = {"KB_VERIFY_PIPE": str(pipe)}
    if chunk is not None:
        env["KB_VERIFY_CHUNK"] = str(chunk)
    os.environ.update(env)
    try:
        c = kb.Context(0)
    finally:
        for k in env:
            del os.environ[k]
    try:
        n = 70001
        pks, msgs, sigs = make_sig_batch(golden_records[:256], n, bad_every=5)
        pk, flat, off, sg = pack_batch(pks, msgs, sigs)
        want = ctx.verify_batch(pk, flat, off, sg, schnorr=schnorr)
        # the reference's verdict on a prefix (the whole oracle pass would take minutes); the mutation classes themselves
        # are covered by test_verify_reject_classes
        m = 3000
        assert (want[:m] == coracle.verify_batch(pk[:m], flat, off[:m + 1], sg[:m], nthreads=8, schnorr=schnorr)).all()
        got = c.verify_batch(pk, flat, off, sg, schnorr=schnorr)
        assert (got == want).all()
        assert (c.verify_batch(pk[:1], flat, off[:2], sg[:1], schnorr=schnorr) == want[:1]).all()   # a single signature
        bad = off.copy()
        bad[n - 3] = bad[n - 2] + 1
        with pytest.raises(kb.KBError):
            c.verify_batch(pk, flat, bad, sg, schnorr=schnorr)
        assert (c.verify_batch(pk, flat, off, sg, schnorr=schnorr) == want).all()   # still usable afterwards
    finally:
        c.close()


def test_group_laws_host_mirror(kb, ctx):
    """util/test/group_test.rs:210-555 in miniature: 2G, -1*G + G = 0, DH commutativity, homomorphisms."""
    H = kb.host
    g = H.Point.base()
    two = H.Scalar.set_int64(2)
    assert H.Point().add(g, g) == H.Point().mul(two)
    minus1 = H.Scalar(O.scalar_set_int64(-1))
    assert H.Point().add(H.Point().mul(minus1), g) == H.Point.null()
    a = H.Scalar.set_bytes(hashlib.sha512(b"a").digest())
    b = H.Scalar.set_bytes(hashlib.sha512(b"b").digest())
    pa, pb = H.Point().mul(a), H.Point().mul(b)
    assert H.Point().mul(a, pb) == H.Point().mul(b, pa)
    assert H.Point().add(pa, pb) == H.Point().mul(a + b)
    assert H.Point().mul(a * b) == H.Point().mul(a, pb)
    assert H.Point().sub(pa, pa) == H.Point.null()
    assert H.Scalar.set_int64(0x100) + H.Scalar.one() == H.Scalar(bytes([1, 1]) + bytes(30))   # scalar_test.rs:27-35
    assert H.Scalar().neg(H.Scalar.one()).v == O.scalar_set_int64(-1)                          # scalar_test.rs:38-46
    assert H.Scalar().sub(a, b).v == O.sc_sub(a.v, b.v)
    assert H.Scalar().inv(a).v == O.sc_inv(a.v) and (H.Scalar().div(a, b) * b) == a           # group_test.rs inverse/div laws
    assert H.Scalar.set_bytes(bytes([0, 1, 2, 3])).v.hex() == "00010203" + "00" * 28            # scalar_test.rs:68-75


# ---- committed polynomials / VSS / DKG -------------------------------------------------------------
def _poly(seed, t):
    coeffs = [O.scalar_set_bytes(hashlib.sha512(b"%s/%d" % (seed, j)).digest()) for j in range(t)]
    return coeffs


def test_pubpoly_eval_and_check(kb, ctx, coracle):
    """config 3 shape at reduced size (oracle is O(t) full scalar mults per eval): t=23 commits,
    evaluation points incl. large indices; plus a torsion-contaminated polynomial."""
    t, npoly = 23, 3
    polys = [_poly(b"p%d" % k, t) for k in range(npoly)]
    commits = ctx.point_mul_base_batch(np.frombuffer(b"".join(b"".join(p) for p in polys), dtype=np.uint8).reshape(-1, 32))
    tors = ctx.point_add_batch(commits[t + 1], np.frombuffer(O.WEAK_KEYS[2], dtype=np.uint8))[0]
    commits[t + 1] = tors[0]
    idx = np.array([0, 1, 2, 5, 255, 256, 1023, 40000, 2**32 - 1] * npoly, dtype=np.uint32)
    pid = np.repeat(np.arange(npoly, dtype=np.uint32), 9)
    got, st = ctx.pubpoly_eval_batch(commits, t, pid, idx)
    assert not st.any()
    for k in range(len(idx)):
        cs = [c.tobytes() for c in commits[pid[k] * t:(pid[k] + 1) * t]]
        assert got[k].tobytes() == coracle.pubpoly_eval(cs, int(idx[k]))
    # shares: honest, corrupted
    shares = np.frombuffer(b"".join(O.pripoly_eval(polys[pid[k]], int(idx[k])) for k in range(len(idx))), dtype=np.uint8).reshape(-1, 32).copy()
    shares[4, 0] ^= 1
    verdict = ctx.vss_verify_deals_batch(commits, t, pid, idx, shares)
    want = np.array([coracle.vss_verify_deal([c.tobytes() for c in commits[pid[k] * t:(pid[k] + 1) * t]], int(idx[k]), shares[k].tobytes()) == 1 for k in range(len(idx))])
    assert (verdict.astype(bool) == want).all()
    assert verdict[0] == 1 and verdict[4] == 0
    # host mirror, poly_test.rs:121-137 style
    pp = kb.host.PubPoly([c.tobytes() for c in commits[:t]])
    assert pp.check(3, kb.host.Scalar(O.pripoly_eval(polys[0], 3)))
    assert not pp.check(3, kb.host.Scalar(O.pripoly_eval(polys[0], 4)))
    assert pp.add(pp).eval(2) == kb.host.Point().add(pp.eval(2), pp.eval(2))


def test_dkg_round_small(ctx, coracle):
    """config 4 shape at n=24, t=16: all n^2 (dealer, verifier) share checks, some corrupted."""
    n, t = 24, 16
    polys = [_poly(b"d%d" % d, t) for d in range(n)]
    commits = ctx.point_mul_base_batch(np.frombuffer(b"".join(b"".join(p) for p in polys), dtype=np.uint8).reshape(-1, 32))
    shares = np.frombuffer(b"".join(O.pripoly_eval(polys[d], i) for d in range(n) for i in range(n)), dtype=np.uint8).reshape(-1, 32).copy()
    bad = [(0, 0), (3, 17), (23, 23), (11, 5)]
    for d, i in bad:
        shares[d * n + i, 5] ^= 0x40
    verdict = ctx.dkg_verify_round(n, t, commits, shares)
    want = np.ones((n, n), dtype=np.uint8)
    for d, i in bad:
        want[d, i] = 0
    assert (verdict.reshape(n, n) == want).all()
    for d in (0, 11):   # oracle cross-check of two dealers' rows
        row = coracle.vss_verify_batch(commits[d * t:(d + 1) * t], np.arange(n, dtype=np.uint32), shares[d * n:(d + 1) * n], nthreads=8)
        assert (row == want[d]).all()
    part = ctx.dkg_verify_round(n, t, commits, shares, dealer_lo=8, dealer_hi=16)
    assert (part.reshape(n, n)[8:16] == want[8:16]).all()


@pytest.fixture(scope="module")
def ctx_fd(kb):
    """Contexts that always run DKG rounds by forward differences (csrc/dkgfd.cuh), one per number of coefficient
    blocks 1..4 (index 0: blocks chosen by cost); the default context picks FD or Horner by cost."""
    cs = []
    for parts in (0, 1, 2, 3, 4):
        os.environ["KB_DKG_FD"] = "1"
        if parts:
            os.environ["KB_FD_PARTS"] = str(parts)
        try:
            cs.append(kb.Context(0))
        finally:
            del os.environ["KB_DKG_FD"]
            os.environ.pop("KB_FD_PARTS", None)
    # the variants small rounds would not reach by default: one lane per cell / item (instead of the four-lane kernels
    # of small rounds), kernel-by-kernel launches (instead of the CUDA graph), every variant of the step kernel
    for extra in ({"KB_FD_Q4_MAX": "0", "KB_FD_CHECK_Q4_MAX": "0", "KB_FD_GRAPH": "0", "KB_FD_STEPS_MINB": "3"},
                  {"KB_FD_Q4_MAX": "100000000", "KB_FD_CHECK_Q4_MAX": "100000000", "KB_FD_STEPS_MINB": "4"},
                  {"KB_FD_Q4_MAX": "0", "KB_FD_STEPS_WIDE": "1", "KB_FD_PARTS": "2"}):
        env = dict(extra, KB_DKG_FD="1")
        os.environ.update(env)
        try:
            cs.append(kb.Context(0))
        finally:
            for k in env:
                del os.environ[k]
    yield cs
    for c in cs:
        c.close()


@pytest.fixture(scope="module")
def ctx_horner(kb):
    os.environ["KB_DKG_FD"] = "0"
    try:
        c = kb.Context(0)
    finally:
        del os.environ["KB_DKG_FD"]
    yield c
    c.close()


@pytest.mark.parametrize("n,t,nd", [(24, 16, 24), (5, 1, 3), (7, 2, 4), (9, 3, 33), (40, 40, 5), (12, 30, 7), (130, 67, 40), (20, 9, 200), (70, 130, 9), (300, 256, 3)])
def test_dkg_round_forward_differences(ctx_horner, ctx_fd, coracle, n, t, nd):
    """The forward-difference round (coefficient blocks, binomial-basis conversion, one-launch difference steps,
    Straus combination) gives the verdicts of the per-share Horner kernel and of the oracle for every number of
    blocks: honest and corrupted shares, a dealer whose commitment carries a small-order component (integer identities
    only: exact there too), a dealer with an undecodable commitment; the commitments as 32-byte encodings and as the
    reference's raw limbs."""
    polys = [_poly(b"fd%d" % d, t) for d in range(nd)]
    commits = ctx_horner.point_mul_base_batch(np.frombuffer(b"".join(b"".join(p) for p in polys), dtype=np.uint8).reshape(-1, 32)).copy()
    shares = np.frombuffer(b"".join(O.pripoly_eval(polys[d], i) for d in range(nd) for i in range(n)), dtype=np.uint8).reshape(-1, 32).copy()
    rng = np.random.default_rng(n * 1000 + t)
    for _ in range(max(2, nd * n // 10)):
        shares[rng.integers(0, nd * n), rng.integers(0, 31)] ^= 1 << rng.integers(0, 8)
    if nd > 2:   # dealer 1: torsion-contaminated commitment; dealer 2: undecodable commitment
        j = min(1, t - 1)
        c = O.point_add(O.point_decode(commits[1 * t + j].tobytes()), O.point_decode(O.WEAK_KEYS[2]))
        commits[1 * t + j] = np.frombuffer(O.point_encode(c), dtype=np.uint8)
    limbs = np.stack([coracle.point_limbs(c.tobytes()) for c in commits])
    if nd > 2:
        k = 0
        while coracle.point_decode_ok(bytes([k]) + b"\x13" * 31):
            k += 1
        commits[2 * t + (t - 1)] = np.frombuffer(bytes([k]) + b"\x13" * 31, dtype=np.uint8)
        limbs[2 * t + (t - 1), 30] += 1    # T no longer equals X Y / Z
    a = ctx_horner.dkg_verify_round(n, t, commits, shares, dealer_lo=0, dealer_hi=nd).reshape(-1)[:nd * n]
    for parts, c in enumerate(ctx_fd):
        b = c.dkg_verify_round(n, t, commits, shares, dealer_lo=0, dealer_hi=nd).reshape(-1)[:nd * n]
        assert (a == b).all(), (parts, np.nonzero(a != b)[0][:10])
    bl = ctx_fd[0].dkg_verify_round(n, t, limbs, shares, limbs=True).reshape(-1)
    assert (a == bl).all(), np.nonzero(a != bl)[0][:10]
    hl = ctx_horner.dkg_verify_round(n, t, limbs, shares, limbs=True).reshape(-1)
    assert (a == hl).all(), np.nonzero(a != hl)[0][:10]
    for d in range(min(nd, 4)):
        if nd > 2 and d == 2:
            assert not a[d * n:(d + 1) * n].any()
            continue
        row = coracle.vss_verify_batch(commits[d * t:(d + 1) * t], np.arange(n, dtype=np.uint32), shares[d * n:(d + 1) * n], nthreads=4)
        assert (row == a[d * n:(d + 1) * n]).all()
    # a sub-range of the dealers (what a rank owns when the round is sharded)
    if nd >= 4:
        part = ctx_fd[0].dkg_verify_round(n, t, commits, shares, dealer_lo=1, dealer_hi=nd - 1)
        assert (part[n:(nd - 1) * n] == a[n:(nd - 1) * n]).all()


# ---- MSM -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 33, 1000, 5000])
def test_msm_matches_oracle(ctx, coracle, golden_records, n):
    pts = np.tile(_golden_pks(golden_records, 1024), (5, 1))[:n]
    s = np.frombuffer(xof_bytes("msm%d" % n, 32 * max(n, 1)), dtype=np.uint8).reshape(-1, 32)[:n].copy()
    if n >= 1000:
        s[:, 31] &= 0x7F
    enc, bad = ctx.msm(s, pts)
    assert bad == 0
    want = coracle.msm(s, pts) if n else (1).to_bytes(32, "little")
    assert enc == want


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1000, 70000])
def test_msm_decoded_points(ctx, coracle, golden_records, n):
    """kb_dev_msm_ext: the MSM over points that are already decoded (X, Y, Z, T words, any Z) equals the MSM over their
    encodings; a point with Z = 0 or off the curve is counted as bad."""
    import torch

    dev = torch.device("cuda", 0)
    pts = _golden_pks(golden_records, min(n, 1024))
    pts = np.tile(pts, ((n + pts.shape[0] - 1) // pts.shape[0], 1))[:n]
    s = random_scalars("msm-ext%d" % n, min(n, 2048))
    s = np.tile(s, ((n + s.shape[0] - 1) // s.shape[0], 1))[:n]
    want, bad = ctx.msm(s, pts)
    assert bad == 0
    raw, st = ctx.point_decompress_batch(pts)
    assert not st.any()
    # a projective representative with Z != 1: scale through an addition with the identity-free path — (X,Y,Z,T) of 2P - P
    d_raw = torch.from_numpy(raw.view(np.uint8).reshape(n, 128)).to(dev)
    d_s = torch.from_numpy(s).to(dev)
    d_out = torch.zeros(32, dtype=torch.uint8, device=dev)
    d_part = torch.zeros(128, dtype=torch.uint8, device=dev)
    d_bad = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.dev_msm_ext(n, d_s, d_raw, d_out, d_part, d_bad)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == want and int(d_bad.item()) == 0
    # the partial of an MSM has Z != 1: feed it back as a one-point MSM with scalar 1
    one = torch.zeros(1, 32, dtype=torch.uint8, device=dev)
    one[0, 0] = 1
    ctx.dev_msm_ext(1, one, d_part.reshape(1, 128), d_out, None, d_bad)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == want
    if n >= 9:
        raw2 = raw.copy()
        raw2[3, 16:24] = 0                 # Z = 0
        raw2[8, 0] ^= 1                    # X off the curve
        ctx.dev_msm_ext(n, d_s, torch.from_numpy(raw2.view(np.uint8).reshape(n, 128)).to(dev), d_out, None, d_bad)
        torch.cuda.synchronize()
        assert int(d_bad.item()) == 2


def test_msm_skew_and_linearity(ctx, coracle, golden_records):
    """Size-independent properties at a size the oracle cannot reach quickly (2^16): linearity in the
    scalars, equal scalars (one giant bucket per window), partial + point_sum == whole."""
    n = 1 << 16
    pts = np.tile(_golden_pks(golden_records, 1024), (n // 1024, 1))
    a = random_scalars("msm/a", 1024)
    a = np.tile(a, (n // 1024, 1))
    ones = np.zeros((n, 32), dtype=np.uint8)
    ones[:, 0] = 1
    ea, _ = ctx.msm(a, pts)
    e1, _ = ctx.msm(ones, pts)
    a1 = ctx.sc_muladd_batch(a, ones, ones)           # a*1 + 1
    ea1, _ = ctx.msm(a1, pts)
    assert ctx.point_add_batch(np.frombuffer(ea, np.uint8), np.frombuffer(e1, np.uint8))[0][0].tobytes() == ea1
    # sum of the 1024 distinct points times 64, via the oracle on the small set
    small = coracle.msm(ones[:1024], pts[:1024])
    sixty4 = np.zeros((1, 32), dtype=np.uint8)
    sixty4[0, 0] = 64
    assert e1 == coracle.mul(sixty4[0].tobytes(), small)
    same = np.tile(a[:1], (n, 1))
    es, _ = ctx.msm(same, pts)
    assert es == coracle.mul(a[0].tobytes(), e1)
    # sharded: two halves -> partials -> fold
    _, p0, _ = ctx.msm(a[: n // 2], pts[: n // 2], want_partial=True)
    _, p1, _ = ctx.msm(a[n // 2:], pts[n // 2:], want_partial=True)
    assert ctx.point_sum(np.stack([p0, p1])) == ea


def test_full_size_signature_batch_properties(ctx, coracle, golden_records):
    """BASELINE config 2 size (2^20): statuses of a tiled batch must be periodic with the tile, and the
    first tile must equal the oracle's."""
    tile = 4096
    pks, msgs, sigs = make_sig_batch(golden_records[:256], tile, bad_every=8)
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    reps = (1 << 20) // tile
    pk_f = np.tile(pk, (reps, 1))
    sg_f = np.tile(sg, (reps, 1))
    flat_f = np.tile(flat, reps)
    off_f = (np.arange(reps, dtype=np.uint64)[:, None] * np.uint64(off[-1]) + off[None, :-1]).reshape(-1)
    off_f = np.concatenate([off_f, [np.uint64(reps) * off[-1]]]).astype(np.uint64)
    st = ctx.verify_batch(pk_f, flat_f, off_f, sg_f)
    want = coracle.verify_batch(pk, flat, off, sg, nthreads=8)
    assert (st.reshape(reps, tile) == want[None, :]).all()


def test_rabin_dss_and_signing_compositions(kb, ctx, coracle):
    """Rows a18 (rabin verify_deal), a20 (DSS partial signatures) and next-row f1 (batched signing), built
    from the batch primitives in host.py, against the big-int oracle."""
    H = kb.host
    t, n = 6, 10
    f = _poly(b"rabin-f", t)
    g = _poly(b"rabin-g", t)
    h_pt = O.point_mul(O.scalar_set_bytes(hashlib.sha512(b"H").digest()))
    commits = [O.point_add(O.point_mul(a), O.point_mul(b, h_pt)) for a, b in zip(f, g)]   # rabin/vss.rs:317-367
    enc = [O.point_encode(c) for c in commits]
    fs = [O.pripoly_eval(f, i) for i in range(n)]
    gs = [O.pripoly_eval(g, i) for i in range(n)]
    gs[4] = O.sc_add(gs[4], O.scalar_set_int64(1))
    got = H.vss_rabin_verify_deals_batch(enc, range(n), fs, gs, H.Point(O.point_encode(h_pt)))
    want = [O.vss_rabin_verify_deal(commits, i, fs[i], gs[i], h_pt) for i in range(n)]
    assert got.astype(bool).tolist() == want and want.count(False) == 1
    # DSS: partial_i = r_i + hash * l_i verifies; a corrupted one does not (dss_test.rs)
    rp, lp = _poly(b"dss-r", t), _poly(b"dss-l", t)
    rc, lc = O.pripoly_commit(rp), O.pripoly_commit(lp)
    hs = O.dss_hash_sig(rc[0], lc[0], b"msg")                     # hash_sig (dss_sig.rs:312-326)
    parts = [O.sc_add(O.pripoly_eval(rp, i), O.sc_mul(hs, O.pripoly_eval(lp, i))) for i in range(n)]
    parts[7] = O.sc_add(parts[7], O.scalar_set_int64(2))
    got, hgot = H.dss_verify_partials_batch([O.point_encode(c) for c in rc], [O.point_encode(c) for c in lc], range(n), parts, b"msg")
    want = [O.dss_verify_partial(rc, lc, i, parts[i], hs) for i in range(n)]
    assert hgot.v == hs and got.astype(bool).tolist() == want and want.count(False) == 1
    # batched Schnorr signing: equals the oracle's signatures and verifies under BOTH verifiers
    priv = [O.scalar_set_bytes(hashlib.sha512(b"x%d" % i).digest()) for i in range(16)]
    nonce = [O.scalar_set_bytes(hashlib.sha512(b"k%d" % i).digest()) for i in range(16)]
    msgs = [b"m" * i for i in range(16)]
    sigs, pubs = H.schnorr_sign_batch(priv, msgs, nonce)
    for i in range(16):
        assert sigs[i].tobytes() == O.schnorr_sign(priv[i], msgs[i], nonce[i])
    assert not H.eddsa_verify_batch([p.tobytes() for p in pubs], msgs, [s.tobytes() for s in sigs]).any()
    assert not H.schnorr_verify_batch([p.tobytes() for p in pubs], msgs, [s.tobytes() for s in sigs]).any()


def test_cfg1_full_size(ctx, coracle, golden_records):
    """BASELINE config 1 at its full size: 2^16 scalars, base point and variable base (2^16 distinct
    points), constant-time and vartime paths, every output byte against the oracle."""
    import os as _os

    n = 1 << 16
    threads = len(_os.sched_getaffinity(0))
    s = random_scalars("kyber-b200/cfg1/full", n)
    want_base = coracle.mul_base_batch(s, nthreads=threads)
    for flags in (0, 1):
        assert (ctx.point_mul_base_batch(s, flags) == want_base).all()
    pts = want_base                                  # 2^16 distinct prime-order points
    s2 = np.roll(s, 1, axis=0)
    want = coracle.mul_batch(s2, pts, nthreads=threads)
    for flags in (0, 1):
        got, st = ctx.point_mul_batch(s2, pts, flags)
        assert not st.any() and (got == want).all()


def test_cfg3_vss_full_size(ctx, coracle):
    """BASELINE config 3 at its full size: one polynomial with t = 171 commitments, the 256 honest shares
    plus 8 corrupted ones, every verdict against the oracle (which runs the reference's t full
    constant-time scalar mults per check)."""
    import os as _os

    n, t = 256, 171
    coeffs = _poly(b"cfg3", t)
    commits = ctx.point_mul_base_batch(np.frombuffer(b"".join(coeffs), dtype=np.uint8).reshape(-1, 32))
    idx = np.concatenate([np.arange(n), np.arange(0, n, 32)]).astype(np.uint32)
    shares = np.frombuffer(b"".join(O.pripoly_eval(coeffs, int(i)) for i in idx), dtype=np.uint8).reshape(-1, 32).copy()
    shares[n:, 7] ^= 0x20
    verdict = ctx.vss_verify_deals_batch(commits, t, np.zeros_like(idx), idx, shares)
    want = coracle.vss_verify_batch(commits, idx, shares, nthreads=len(_os.sched_getaffinity(0)))
    assert (verdict == want).all()
    assert verdict[:n].all() and not verdict[n:].any()
    ev, st = ctx.pubpoly_eval_batch(commits, t, np.zeros(4, dtype=np.uint32), np.array([0, 100, 255, 1023], dtype=np.uint32))
    for k, i in enumerate([0, 100, 255, 1023]):
        assert ev[k].tobytes() == coracle.pubpoly_eval([c.tobytes() for c in commits], i)


def test_ref10_limb_wire_format(ctx, coracle):
    """kb_point_from_limbs_batch: the reference's serde form of a Point (raw limbs) -> marshal_binary bytes,
    incl. Point::default() (all-zero limbs -> 32 zero bytes) inside a batch-inversion group."""
    n = 200
    sc = random_scalars("limbs", n)
    limbs = np.stack([coracle.mul_base_limbs(s.tobytes()) for s in sc])
    limbs[5] = 0          # Point::default()
    limbs[77, 20:30] = 0  # Z = 0 with X, Y != 0
    got = ctx.point_from_limbs_batch(limbs)
    for i in range(n):
        assert got[i].tobytes() == coracle.limbs_tobytes(limbs[i]), i
    assert not got[5].any() and not got[77].any()


def test_pubpoly_sum_dkg_key(ctx, coracle, golden_records):
    """dkg_key (dkg.rs:905-954): the distributed public polynomial is the coefficient-wise sum of the dealers'
    commitment polynomials (PubPoly::add, poly.rs:486)."""
    npoly, t = 37, 11
    pts = _golden_pks(golden_records, npoly * t)
    got, st = ctx.pubpoly_sum(pts, t)
    assert not st.any()
    for j in range(t):
        acc = (1).to_bytes(32, "little")
        for d in range(npoly):
            acc = coracle.point_add(acc, pts[d * t + j].tobytes())
        assert got[j].tobytes() == acc
    one, _ = ctx.pubpoly_sum(pts[:t], t)
    assert (one == pts[:t]).all()


@pytest.mark.parametrize("path", ["horner", "fd-3-blocks", "fd-4-blocks"])
def test_cfg4_shape_full_t_two_dealers(ctx, ctx_horner, ctx_fd, coracle, path):
    """BASELINE config 4 shape at full threshold and full verifier count (n = 1024, t = 683) for two dealers
    (the full round is 1024 dealers — CPU-days for the oracle): honest shares are computed independently with
    Python integers (PriPoly::eval, poly.rs:133), a few are corrupted, dealer 1 carries a torsion-contaminated
    commitment (the check then only passes where 8 | x); 12 verdicts are cross-checked with the
    oracle, which runs the reference's 683 full scalar mults per check."""
    n, t, nd = 1024, 683, 2
    L = O.L
    coeff = [[int.from_bytes(hashlib.sha512(b"cfg4/%d/%d" % (d, j)).digest(), "little") % L for j in range(t)] for d in range(nd)]
    commits = ctx.point_mul_base_batch(np.frombuffer(b"".join(c.to_bytes(32, "little") for row in coeff for c in row), dtype=np.uint8).reshape(-1, 32))
    shares = np.zeros((nd * n, 32), dtype=np.uint8)
    for d in range(nd):
        for i in range(n):
            v = 0
            for cj in reversed(coeff[d]):
                v = (v * (i + 1) + cj) % L
            shares[d * n + i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    want = np.ones((nd, n), dtype=np.uint8)
    for d, i in ((0, 0), (0, 511), (0, 1023), (1, 7), (1, 640)):
        shares[d * n + i, 9] ^= 4
        want[d, i] = 0
    t8 = np.frombuffer(O.WEAK_KEYS[2], dtype=np.uint8)
    commits[t + 1] = ctx.point_add_batch(commits[t + 1], t8)[0][0]
    torsion_ok = np.zeros(n, dtype=np.uint8)
    torsion_ok[7::8] = 1            # x = i + 1 divisible by 8
    want[1] &= torsion_ok
    # two dealers alone are below the cost threshold of the forward-difference round: ctx_fd forces it, so that its
    # conversion launches, the x^(q h) mod 8L table and all 1024 difference steps are held to the same expected verdicts
    c = {"horner": ctx_horner, "fd-3-blocks": ctx_fd[3], "fd-4-blocks": ctx_fd[4]}[path]
    got = c.dkg_verify_round(n, t, commits, shares).reshape(nd, n)
    assert (got == want).all(), np.argwhere(got != want)[:10]
    for d, i in ((0, 0), (0, 1), (0, 1023), (1, 7), (1, 8), (1, 15)):
        cs = commits[d * t:(d + 1) * t]
        assert coracle.vss_verify_deal([c.tobytes() for c in cs], i, shares[d * n + i].tobytes()) == int(got[d, i])


def test_cfg5_msm_oracle_2p14_and_linearity_2p20(ctx, coracle, golden_records):
    """BASELINE config 5: direct oracle comparison at 2^14 points (the oracle folds 2^14 full scalar mults),
    then at 2^20 the size-independent property msm(a, P) + msm(b, P) == msm(a + b, P) and shard-and-fold."""
    n = 1 << 14
    sc = random_scalars("cfg5/2p14", n)
    pts = ctx.point_mul_base_batch(np.roll(sc, 3, axis=0), 1)
    enc, bad = ctx.msm(sc, pts)
    assert bad == 0 and enc == coracle.msm(sc, pts)
    big = 1 << 20
    a = np.tile(sc, (big // n, 1))
    a[:, 1] ^= (np.arange(big) >> 14).astype(np.uint8)      # distinct scalars per tile
    b = np.roll(a, 5, axis=0)
    p = np.tile(pts, (big // n, 1))
    ab = ctx.sc_muladd_batch(a, np.tile(np.frombuffer((1).to_bytes(32, "little"), np.uint8), (big, 1)), b)   # a*1 + b mod L
    ea, _ = ctx.msm(a, p)
    eb, _ = ctx.msm(b, p)
    eab, _ = ctx.msm(ab, p)
    assert ctx.point_add_batch(np.frombuffer(ea, np.uint8), np.frombuffer(eb, np.uint8))[0][0].tobytes() == eab
    parts = [ctx.msm(a[k * (big // 4):(k + 1) * (big // 4)], p[k * (big // 4):(k + 1) * (big // 4)], want_partial=True)[1] for k in range(4)]
    assert ctx.point_sum(np.stack(parts)) == ea


def test_eddsa_sign_golden_file(kb, ctx, golden_records):
    """kb_eddsa_sign_batch reproduces ALL 1024 golden signatures and public keys of the reference's sign.input
    (tests/sign/eddsa.rs:37-94: message lengths 0..1023), then 10 000 random sign/verify round trips with the
    reference's `sig[63] & 0xe0 == 0` property (tests/sign/eddsa.rs:18-33)."""
    sigs, pks = kb.host.eddsa_sign_batch([r[0] for r in golden_records], [r[3] for r in golden_records])
    assert sigs.tobytes() == b"".join(r[2] for r in golden_records)
    assert pks.tobytes() == b"".join(r[1] for r in golden_records)
    n = 10000
    seeds = np.frombuffer(xof_bytes("eddsa/10000/seeds", 32 * n), dtype=np.uint8).reshape(n, 32)
    msgs = [xof_bytes("eddsa/10000/msg%d" % (i % 7), 32 + (i % 100)) for i in range(n)]
    sigs, pks = kb.host.eddsa_sign_batch([s.tobytes() for s in seeds], msgs)
    assert not (sigs[:, 63] & 0xE0).any()
    st = kb.host.eddsa_verify_batch([p.tobytes() for p in pks], msgs, [s.tobytes() for s in sigs])
    assert not st.any()
    st = kb.host.schnorr_verify_batch([p.tobytes() for p in pks], msgs, [s.tobytes() for s in sigs])
    assert not st.any()


def test_recover_commit_lagrange_msm(kb, ctx):
    """share/poly_test.rs recover tests (n = 10, t = 6 there; also a larger one): the secret commitment p(0) is
    recovered from t public shares by Lagrange interpolation in the exponent — here one kb_recover_commit_batch call —
    and equals commit[0]; with fewer than t shares the reference's error is raised."""
    H = kb.host
    for n, t, seed in ((10, 6, b"rc1"), (40, 27, b"rc2")):
        coeffs = _poly(seed, t)
        commits = O.pripoly_commit(coeffs)
        pub = H.PubPoly([O.point_encode(c) for c in commits])
        idx = list(range(n))
        shares = [(i, p) for i, p in zip(idx, pub.eval_batch(idx))]
        shares[1] = (1, None)                    # a missing share, as in poly_test.rs
        shares = shares[::-1]                    # order must not matter (xy_commit sorts)
        got = H.recover_commit(shares, t, n)
        assert got.b == O.point_encode(commits[0])
        # oracle-side Lagrange with the same t shares
        good = sorted((i for i, p in shares if p is not None))[:t]
        lam = []
        for i in good:
            num = den = 1
            for j in good:
                if j != i:
                    num = num * (j + 1) % O.L
                    den = den * ((j + 1) - (i + 1)) % O.L
            lam.append((num * pow(den, O.L - 2, O.L) % O.L).to_bytes(32, "little"))
        pts = [O.pubpoly_eval(commits, i) for i in good]
        assert got.b == O.point_encode(O.msm(lam, pts))
    with pytest.raises(ValueError):
        H.recover_commit(shares[:3], t, n)
