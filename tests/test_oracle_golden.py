"""Pins the two oracles (oracle/ed25519_bigint.py, oracle/ref10_port.c) to every golden vector /
KAT the reference's own tests hold for the hot path, and to libsodium.
CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ed25519_bigint as O

RFC8032 = [  # sign/eddsa/eddsa_test.rs:20-46 (RFC 8032 §7.1; the 1023-byte case is golden line 1024)
    ("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60", "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a", "",
     "e5564300c360ac729086e2cc806e828a84877f1eb8e5d974d873e065224901555fb8821590a33bacc61e39701cf9b46bd25bf5f0595bbe24655141438e7a100b"),
    ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb", "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c", "72",
     "92a009a9f0d4cab8720e820b5f642540a2b27b5416503f8fb3762223ebdb69da085ac1e43e15996e458f3613d0f11d8c387b2eaeb4302aeeb00d291612bb0c00"),
    ("c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7", "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025", "af82",
     "6291d657deec24024827e69c3abe01a30ce548a284743a445e3680d7db5ac3ac18ff9b538d16f290ae67f760984dc6594a7c15e9716ed28dc027beceea1ec40a"),
    ("833fe62409237b9d62ec77587520911e9a759cec1d19755b7da901b96dca3d42", "ec172b93ad5e563bf4932c70e1245034c35467ef2efd4d64ebf819683467e2bf",
     "ddaf35a193617abacc417349ae20413112e6fa4e89a97ea20a9eeee64b55d39a2192992a274fc1a836ba3c23a3feebbd454d4423643ce80e2a9ac94fa54ca49f",
     "dc2a4459e7369633a52b1bf277839a00201009a3efbf3ecb69bea2186c26b58909351fc9ac90b3ecfdfbc7c66431e0303dca179c138ac17ad9bef1177331a704"),
]

L_BYTES = O.L.to_bytes(32, "little")
NONCANON = bytes([0xEF]) + b"\xff" * 31                   # eddsa_test.rs:168-218
SMALL_ORDER = O.WEAK_KEYS[3]                               # eddsa_test.rs:222-272 (c7176a70…037a)
GO_PK = bytes.fromhex("7d4d0e7f6153a69b6242b522abbee685fda4420f8834b108c3bdae369ef549fa")   # eddsa_test.rs:144-163
GO_SIG = bytes.fromhex("7c38e026f29e14aabd059a0f2db8b0cd783040609a8be684db12f82a27774ab0"
                       "67654bce3832c2d76f8f6f5dafc08d9339d4eef676573336a5c51eb6f946b31d")


def add_l(sig: bytes) -> bytes:
    s = (int.from_bytes(sig[32:], "little") + O.L) % (1 << 256)
    return sig[:32] + s.to_bytes(32, "little")


def test_golden_file_c_oracle(golden_records, coracle):
    """tests/sign/eddsa.rs:37-94 — all 1024 lines: pk derivation (base mul + compress),
    and full verify through both verifiers."""
    for seed, pk, sig, msg in golden_records:
        a, _ = O.clamp_key(seed)
        assert coracle.mul_base(a) == pk
        assert coracle.eddsa_verify(pk, msg, sig) == O.OK
        assert coracle.schnorr_verify(pk, msg, sig) == O.OK


def test_golden_file_bigint_oracle(golden_records):
    """Same file, sampled (Python big-ints are slow): also checks deterministic signing,
    which pins SHA-512 -> mod L and (r + h*a) mod L."""
    for seed, pk, sig, msg in golden_records[::16] + golden_records[:8]:
        assert O.eddsa_public(seed) == pk
        assert O.eddsa_sign(seed, msg) == sig
        assert O.eddsa_verify(pk, msg, sig) == O.OK
        assert O.schnorr_verify(pk, msg, sig) == O.OK


def test_rfc8032_vectors(coracle):
    for sk, pk, msg, sig in RFC8032:
        sk, pk, msg, sig = map(bytes.fromhex, (sk, pk, msg, sig))
        assert O.eddsa_public(sk) == pk
        assert O.eddsa_sign(sk, msg) == sig
        for verify in (O.eddsa_verify, coracle.eddsa_verify, O.schnorr_verify, coracle.schnorr_verify):
            assert verify(pk, msg, sig) == O.OK
            assert verify(pk, msg + b"x", sig) == O.ERR_INVALID_SIGNATURE


def test_reject_vectors(golden_records, coracle):
    """eddsa_test.rs:111-272 and schnorr_test.rs:85-110."""
    seed, pk, sig, msg = golden_records[100]
    for ed, sc in ((O.eddsa_verify, O.schnorr_verify), (coracle.eddsa_verify, coracle.schnorr_verify)):
        assert ed(pk, msg, add_l(sig)) == O.ERR_SIG_NOT_CANONICAL
        assert sc(pk, msg, add_l(sig)) == O.ERR_SIG_NOT_CANONICAL
        assert ed(GO_PK, b"Test", GO_SIG) == O.ERR_SIG_NOT_CANONICAL
        assert ed(pk, msg, NONCANON + sig[32:]) == O.ERR_R_NOT_CANONICAL
        assert ed(NONCANON, msg, sig) == O.ERR_PK_NOT_CANONICAL
        assert ed(pk, msg, SMALL_ORDER + sig[32:]) == O.ERR_R_SMALL_ORDER
        assert ed(SMALL_ORDER, msg, sig) == O.ERR_PK_SMALL_ORDER
        assert ed(pk, msg, sig[:63]) == O.ERR_SIG_LENGTH
        assert sc(pk, msg, sig[:63]) == O.ERR_SIG_LENGTH
        # schnorr_test.rs:6-40: wrong R / s / pk are rejected
        bad = bytearray(sig); bad[40] ^= 1
        assert sc(pk, msg, bytes(bad)) == O.ERR_INVALID_SIGNATURE
        # check ORDER differs between the verifiers: bad s AND non-canonical R
        both = NONCANON + add_l(sig)[32:]
        assert ed(pk, msg, both) == O.ERR_SIG_NOT_CANONICAL
        assert sc(pk, msg, both) in (O.ERR_R_NOT_CANONICAL, O.ERR_MARSHALLING)


def test_weak_keys_and_decode_kat(coracle):
    for k in O.WEAK_KEYS:                                  # point_test.rs:19-25
        p = O.point_decode(k)
        assert p is not None and O.point_has_small_order(p)
        assert coracle.point_has_small_order(k) == 1
    kat = bytes([132, 100, 171, 115, 11, 183, 255, 50, 148, 134, 171, 221, 113, 152, 106, 84, 177, 153, 88,
                 19, 80, 57, 234, 7, 56, 227, 90, 220, 227, 87, 78, 223])      # ge.rs:65-73
    assert O.point_decode(kat) is not None and coracle.point_decode_ok(kat)


def test_scalar_kats(coracle):
    # scalar_test.rs:27-75
    assert O.sc_add(O.scalar_set_int64(0x100), O.scalar_set_int64(1)).hex() == "0101" + "00" * 30
    assert O.scalar_set_int64(-1).hex() == "ecd3f55c1a631258d69cf7a2def9de1400000000000000000000000000000010"
    assert O.scalar_set_bytes(bytes([0, 1, 2, 3])).hex() == "00010203" + "00" * 28
    # scalar_test.rs:89-105: L-2 .. L+1 -> [T, T, F, F]
    for k, want in zip(range(-2, 2), [True, True, False, False]):
        b = (O.L + k).to_bytes(32, "little")
        assert O.scalar_is_canonical(b) is want
        assert coracle.scalar_is_canonical(b) is want
    for _ in range(300):
        d = os.urandom(64)
        assert coracle.sc_reduce64(d) == O.scalar_set_bytes(d)
        a, b, c = os.urandom(32), os.urandom(32), os.urandom(32)
        assert coracle.sc_muladd(a, b, c) == O.sc_mul_add(a, b, c)
    for d in (b"\xff" * 64, b"\x00" * 64, L_BYTES + b"\x00" * 32, b"\x00" * 32 + L_BYTES):
        assert coracle.sc_reduce64(d) == O.scalar_set_bytes(d)


def test_sha512(coracle):
    for n in (0, 1, 55, 111, 112, 113, 127, 128, 129, 255, 256, 1000):
        m = os.urandom(n)
        assert coracle.sha512(m) == hashlib.sha512(m).digest()


def test_point_is_canonical_quirk(coracle):
    """217 canonical encodings are reported non-canonical, plus the 19 true ones."""
    bad = 0
    for b0 in range(256):
        for top in (0x7F, 0xFF, 0x7E):
            b = bytes([b0]) + b"\xff" * 30 + bytes([top])
            assert coracle.point_is_canonical(b) == O.point_is_canonical(b)
        bad += not O.point_is_canonical(bytes([b0]) + b"\xff" * 30 + b"\x7f")
    assert bad == 236
    assert O.point_is_canonical(bytes([0x13]) + b"\xff" * 30 + b"\x7f")
    assert not O.point_is_canonical(bytes([0x14]) + b"\xff" * 30 + b"\x7f")


def test_c_oracle_matches_bigint_mul(golden_records, coracle):
    """Point::mul on in-domain and out-of-domain (a[31] > 127) scalars."""
    for _, pk, _, _ in golden_records[:24]:
        pt = O.point_decode(pk)
        for s in (os.urandom(31) + bytes([os.urandom(1)[0] & 0x7F]), os.urandom(32), b"\xff" * 32,
                  bytes(31) + b"\x80", b"\x88" * 31 + b"\x88"):
            assert coracle.mul(s, pk) == O.point_encode(O.point_mul(s, pt))
            assert coracle.mul_base(s) == O.point_encode(O.point_mul(s))


def test_libsodium_cross_check(golden_records, coracle):
    nb = pytest.importorskip("nacl.bindings")
    for seed, pk, sig, msg in golden_records[:32]:
        s = (int.from_bytes(hashlib.sha512(seed).digest(), "little") % O.L).to_bytes(32, "little")
        assert nb.crypto_scalarmult_ed25519_base_noclamp(s) == coracle.mul_base(s)
        if int.from_bytes(s, "little") != 0:
            assert nb.crypto_scalarmult_ed25519_noclamp(s, pk) == coracle.mul(s, pk)
        assert nb.crypto_core_ed25519_add(pk, sig[:32]) == coracle.point_add(pk, sig[:32])
        assert nb.crypto_core_ed25519_sub(pk, sig[:32]) == coracle.point_add(pk, sig[:32], True)
        nb.crypto_sign_open(sig + msg, pk)


def test_pubpoly_eval_and_deal(golden_records, coracle):
    """No KAT exists in the reference (poly_test.rs is self-consistency only): check the two
    oracles agree and that an honest share verifies (poly_test.rs:121-137)."""
    coeffs = [O.scalar_set_bytes(hashlib.sha512(bytes([i])).digest()) for i in range(6)]
    commits = O.pripoly_commit(coeffs)
    enc = [O.point_encode(c) for c in commits]
    for i in (0, 1, 9, 700):
        share = O.pripoly_eval(coeffs, i)
        assert coracle.pubpoly_eval(enc, i) == O.point_encode(O.pubpoly_eval(commits, i))
        assert O.pubpoly_check(commits, i, share)
        assert coracle.vss_verify_deal(enc, i, share) == 1
        assert coracle.vss_verify_deal(enc, i, O.sc_add(share, O.scalar_set_int64(1))) == 0
    # torsion-contaminated commitment: Horner by the integer xi, not xi^j mod L
    t8 = O.point_decode(O.WEAK_KEYS[2])
    bad = [commits[0], O.point_add(commits[1], t8)] + commits[2:]
    benc = [O.point_encode(c) for c in bad]
    for i in (1, 2, 6):
        assert coracle.pubpoly_eval(benc, i) == O.point_encode(O.pubpoly_eval(bad, i))


def _share_setup(t, n, tag):
    import hashlib

    coeffs = [O.scalar_set_bytes(hashlib.sha512(b"%s/%d" % (tag, j)).digest()) for j in range(t)]
    commits = O.pripoly_commit(coeffs)
    shares = [O.pripoly_eval(coeffs, i) for i in range(n)]
    return coeffs, commits, shares


def test_interpolation_oracles(coracle):
    """recover_commit / recover_pub_poly (share/poly.rs:566-635) have no known answers in the reference; the two
    restatements (C: the reference's operation sequence on ref10 limbs; Python: big integers) must agree with each
    other AND with what interpolation has to return: the commitment of the secret, resp. the commitments themselves."""
    t, n = 5, 9
    coeffs, commits, shares = _share_setup(t, n, b"interp")
    for pick in ([0, 1, 2, 3, 4], [8, 2, 5, 3, 0], [4, 5, 6, 7, 8]):
        pub = [(i, O.point_mul(shares[i])) for i in pick]
        enc = np.frombuffer(b"".join(O.point_encode(p) for _, p in pub), dtype=np.uint8).reshape(-1, 32)
        want = O.point_encode(commits[0])
        assert O.point_encode(O.recover_commit(pub)) == want
        assert coracle.recover_commit(pick, enc) == want
        got = coracle.recover_pub_poly(pick, enc)
        assert [g.tobytes() for g in got] == [O.point_encode(c) for c in commits]
        assert [O.point_encode(c) for c in O.recover_pub_poly(pub)] == [O.point_encode(c) for c in commits]
    a = bytes(range(1, 33))
    assert O.sc_mul(coracle.sc_invert(a), O.scalar_set_bytes(a)) == (1).to_bytes(32, "little")


def test_rabin_and_dss_oracles(coracle):
    """vss::rabin verify_deal (rabin/vss.rs:889-900) and DSS::process_partial_sig (dss_sig.rs:263-273): C and Python
    restatements agree, honest inputs verify, corrupted ones do not."""
    t, n = 4, 6
    fc, fcom, fsh = _share_setup(t, n, b"rabin-f")
    gc, gcom, gsh = _share_setup(t, n, b"rabin-g")
    H = O.point_mul(bytes(range(7, 39)))
    commits = [O.point_add(a, O.point_mul(g, H)) for a, g in zip(fcom, gc)]     # f_j*G + g_j*H
    enc = np.frombuffer(b"".join(O.point_encode(c) for c in commits), dtype=np.uint8).reshape(-1, 32)
    f = np.frombuffer(b"".join(fsh), dtype=np.uint8).reshape(-1, 32).copy()
    g = np.frombuffer(b"".join(gsh), dtype=np.uint8).reshape(-1, 32).copy()
    g[3, 0] ^= 1
    got = coracle.rabin_verify_batch(enc, O.point_encode(H), np.arange(n), f, g, nthreads=2)
    want = [O.vss_rabin_verify_deal(commits, i, f[i].tobytes(), g[i].tobytes(), H) for i in range(n)]
    assert got.tolist() == [int(w) for w in want] == [1, 1, 1, 0, 1, 1]
    # DSS: partial_i = r_i + hash * l_i
    rc, rcom, rsh = _share_setup(t, n, b"dss-r")
    lc, lcom, lsh = _share_setup(t, n, b"dss-l")
    msg = b"dss message"
    renc = np.frombuffer(b"".join(O.point_encode(c) for c in rcom), dtype=np.uint8).reshape(-1, 32)
    lenc = np.frombuffer(b"".join(O.point_encode(c) for c in lcom), dtype=np.uint8).reshape(-1, 32)
    h = O.dss_hash_sig(rcom[0], lcom[0], msg)
    assert coracle.dss_hash_sig(renc, lenc, msg) == h
    partials = np.frombuffer(b"".join(O.sc_add(rsh[i], O.sc_mul(h, lsh[i])) for i in range(n)), dtype=np.uint8).reshape(-1, 32).copy()
    partials[1, 5] ^= 8
    got = coracle.dss_partial_batch(renc, lenc, msg, np.arange(n), partials, nthreads=2)
    want = [O.dss_verify_partial(rcom, lcom, i, partials[i].tobytes(), h) for i in range(n)]
    assert got.tolist() == [int(w) for w in want] == [1, 0, 1, 1, 1, 1]


def test_session_id_and_find_pub_oracle():
    """session_id (vss/pedersen/vss.rs:1069-1090) is SHA-256 over canonical encodings; a non-canonical input encoding
    (y >= p) of the same point gives the same id because marshal_to re-encodes."""
    import hashlib

    pts = [O.point_mul(bytes([k + 1]) + bytes(31)) for k in range(6)]
    sid = O.session_id(pts[0], pts[1:4], pts[4:6], 2)
    want = hashlib.sha256(b"".join(O.point_encode(p) for p in pts) + (2).to_bytes(4, "little")).digest()
    assert sid == want
    assert O.find_pub(pts, pts[3]) == (3, True) and O.find_pub(pts[:3], pts[4]) == (0, False)
