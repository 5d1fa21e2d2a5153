"""Big-integer restatement of kyber-rs's edwards25519 hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *semantic* oracle: every function restates what one reference
function returns (bytes, accept/reject, error variant), using Python integers.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import it; the product (``kyber-rs_b200``) never does.

Parity pinning: checked in ``tests/test_oracle_golden.py`` against
  * the 1024-case ``sign.input`` golden file of the reference
    (``src/sign/eddsa/testdata/sign.input.gz`` via ``tests/sign/eddsa.rs:37-94``;
    committed as ``tests/golden/sign_input.bin.gz`` by ``tests/golden/make_golden.py``),
  * RFC 8032 §7.1 vectors (``src/sign/eddsa/eddsa_test.rs:20-46``),
  * the reference's reject vectors (``eddsa_test.rs:111-272``),
  * ``WEAK_KEYS`` (``point_test.rs:19-25``), scalar KATs (``scalar_test.rs:27-105``),
    the decode KAT (``ge.rs:65-73``),
  * libsodium (PyNaCl) on well-formed inputs.
``PubPoly::eval`` / ``verify_deal`` / DSS have no known-answer tests in the reference
(SURVEY §8c): for those, parity is defined by this restatement of pinned primitives.

All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P          # constants.rs:60
D2 = (2 * D) % P                                   # constants.rs:65
SQRT_M1 = pow(2, (P - 1) // 4, P)                  # constants.rs:56
BY = (4 * pow(5, P - 2, P)) % P                    # base point y = 4/5 (constants.rs:70)

# status codes shared with include/kyber_b200.h (kb_sig_status); names follow
# sign/error.rs:6-25
OK = 0
ERR_SIG_LENGTH = 1
ERR_SIG_NOT_CANONICAL = 2
ERR_R_NOT_CANONICAL = 3
ERR_R_SMALL_ORDER = 4
ERR_PK_NOT_CANONICAL = 5
ERR_PK_SMALL_ORDER = 6
ERR_MARSHALLING = 7
ERR_INVALID_SIGNATURE = 8

STATUS_NAMES = {
    OK: "ok",
    ERR_SIG_LENGTH: "wrong signature length",
    ERR_SIG_NOT_CANONICAL: "signature is not canonical",
    ERR_R_NOT_CANONICAL: "R is not canonical",
    ERR_R_SMALL_ORDER: "R has small order",
    ERR_PK_NOT_CANONICAL: "public key is not canonical",
    ERR_PK_SMALL_ORDER: "public key has small order",
    ERR_MARSHALLING: "marshalling error",
    ERR_INVALID_SIGNATURE: "signature is not valid",
}

# constants.rs:3744 (values are facts about the curve: 0, 1, the two order-8 y's, p-1)
WEAK_KEYS = [
    bytes(32),
    bytes([1]) + bytes(31),
    (2707385501144840649318225287225658788936804267575313519463743609750303402022).to_bytes(32, "little"),
    (55188659117513257062467267217118295137698188065244968500265048394206261417927).to_bytes(32, "little"),
    (P - 1).to_bytes(32, "little"),
]


# --------------------------------------------------------------------------- field
def fe_from_bytes(s: bytes) -> int:
    """fe.rs:67-77 — bit 255 is dropped; values >= p are NOT rejected (taken mod p)."""
    return (int.from_bytes(s, "little") & ((1 << 255) - 1)) % P


def fe_to_bytes(x: int) -> bytes:
    """fe.rs:147 — canonical little-endian."""
    return (x % P).to_bytes(32, "little")


def fe_is_negative(x: int) -> int:
    """fe.rs:240 — lsb of the canonical encoding."""
    return (x % P) & 1


def fe_inv(x: int) -> int:
    """fe.rs:857 — z^(p-2); 0 -> 0."""
    return pow(x, P - 2, P)


# --------------------------------------------------------------------------- points
# A point is an affine pair (x, y) of ints mod p.  Internal projective coordinates of
# the reference are unobservable (point.rs:227-241 compares canonical bytes).
IDENTITY = (0, 1)


def _recover_x(y: int, sign: int):
    """ge.rs:124-179 — returns x or None.  x=0 with sign=1 is accepted (fe_neg(0)=0)."""
    u = (y * y - 1) % P
    v = (D * y * y + 1) % P
    v3 = (v * v % P) * v % P
    x = (v3 * v3 % P) * v % P * u % P           # u v^7
    x = pow(x, (P - 5) // 8, P)                 # fe_pow22523
    x = x * v3 % P * u % P                      # u v^3 (u v^7)^((p-5)/8)
    vxx = x * x % P * v % P
    if (vxx - u) % P != 0:
        if (vxx + u) % P != 0:
            return None
        x = x * SQRT_M1 % P
    if fe_is_negative(x) != sign:
        x = (-x) % P
    return x


def point_decode(s: bytes):
    """ExtendedGroupElement::set_bytes (ge.rs:124). None on failure (len != 32 or non-square)."""
    if len(s) != 32:
        return None
    y = fe_from_bytes(s)
    x = _recover_x(y, s[31] >> 7)
    if x is None:
        return None
    return (x, y)


def point_encode(pt) -> bytes:
    """ExtendedGroupElement::write_bytes (ge.rs:112-122)."""
    x, y = pt
    b = bytearray(fe_to_bytes(y))
    b[31] ^= fe_is_negative(x) << 7
    return bytes(b)


def point_add(p1, p2):
    """Point::add (point.rs:179) — complete twisted-Edwards law, a = -1."""
    x1, y1 = p1
    x2, y2 = p2
    t = D * x1 % P * x2 % P * y1 % P * y2 % P
    x3 = (x1 * y2 + x2 * y1) % P * fe_inv((1 + t) % P) % P
    y3 = (y1 * y2 + x1 * x2) % P * fe_inv((1 - t) % P) % P
    return (x3, y3)


def point_neg(p1):
    """Point::neg (point.rs:201)."""
    return ((-p1[0]) % P, p1[1])


def point_sub(p1, p2):
    """Point::sub (point.rs:190)."""
    return point_add(p1, point_neg(p2))


def _ext_add(Pt, Qt):
    X1, Y1, Z1, T1 = Pt
    X2, Y2, Z2, T2 = Qt
    A = (Y1 - X1) * (Y2 - X2) % P
    B = (Y1 + X1) * (Y2 + X2) % P
    C = T1 * D2 % P * T2 % P
    Dd = 2 * Z1 * Z2 % P
    E, F, G, H = (B - A) % P, (Dd - C) % P, (Dd + C) % P, (B + A) % P
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def _ext_dbl(Pt):
    X1, Y1, Z1, _ = Pt
    A = X1 * X1 % P
    B = Y1 * Y1 % P
    C = 2 * Z1 * Z1 % P
    H = (A + B) % P
    E = (H - (X1 + Y1) * (X1 + Y1)) % P
    G = (A - B) % P
    F = (C + G) % P
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def _to_ext(pt):
    return (pt[0], pt[1], 1, pt[0] * pt[1] % P)


def _from_ext(Pt):
    zi = fe_inv(Pt[2])
    return (Pt[0] * zi % P, Pt[1] * zi % P)


def _mul_int(k: int, pt):
    """k*pt for a Python int k (may be negative)."""
    if k < 0:
        return _mul_int(-k, point_neg(pt))
    acc = (0, 1, 1, 0)
    base = _to_ext(pt)
    for bit in bin(k)[2:] if k else "":
        acc = _ext_dbl(acc)
        if bit == "1":
            acc = _ext_add(acc, base)
    return _from_ext(acc)


BASE = (_recover_x(BY, 0), BY)


def scalar_digits_radix16(a: bytes):
    """Signed radix-16 recoding shared by ge_scalar_mult_base (ge.rs:443-458) and
    ge_scalar_mult (ge.rs:521-535): 64 digits, e[0..62] in [-8,8), e[63] = top nibble + carry."""
    e = []
    for v in a:
        e.append(v & 15)
        e.append((v >> 4) & 15)
    carry = 0
    for i in range(63):
        e[i] += carry
        carry = (e[i] + 8) >> 4
        e[i] -= carry << 4
    e[63] += carry
    return e


def scalar_effective(a: bytes) -> int:
    """The integer the reference's scalar-mult routines actually multiply by.

    For a[31] <= 127 (the documented precondition, ge.rs:440,506) this is the
    little-endian integer a.  Otherwise the top digit can be 9..16, matches no table
    entry in select_pre_computed/select_cached (ge.rs:423-434, 488-500) and contributes
    the identity — the result may then be negative.
    """
    e = scalar_digits_radix16(a)
    if not (0 <= e[63] <= 8):
        e[63] = 0
    return sum(d << (4 * i) for i, d in enumerate(e))


def point_mul(a: bytes, pt=None):
    """Point::mul (point.rs:207-225): a is the raw 32-byte scalar (NOT reduced mod L);
    pt=None means the standard base point."""
    assert len(a) == 32
    return _mul_int(scalar_effective(a), BASE if pt is None else pt)


def point_eq(p1, p2) -> bool:
    """Point::eq (point.rs:227-241) — canonical encodings compared."""
    return point_encode(p1) == point_encode(p2)


def point_is_canonical(b: bytes) -> bool:
    """Point::is_canonical (point.rs:322-337) INCLUDING its quirk: the
    low-byte test computes 0xED - (1 - b[0]) in wrapping u16 arithmetic, so encodings with
    bytes 1..30 = 0xff, b[31]&0x7f = 0x7f are reported non-canonical for b[0] >= 0x14."""
    if len(b) != 32:
        return False
    c = (b[31] & 0x7F) ^ 0x7F
    for i in range(30, 0, -1):
        c |= b[i] ^ 0xFF
    c = (((c - 1) & 0xFFFF) >> 8) & 0xFF
    d = (((0xED - ((1 - b[0]) & 0xFFFF)) & 0xFFFF) >> 8) & 0xFF
    return 1 - (c & d & 1) == 1


def point_has_small_order(pt) -> bool:
    """Point::has_small_order (point.rs:286-310): works on the RE-ENCODED point."""
    s = point_encode(pt)
    c = [0] * 5
    for j in range(31):
        for i in range(5):
            c[i] |= s[j] ^ WEAK_KEYS[i][j]
    for i in range(5):
        c[i] |= (s[31] & 0x7F) ^ WEAK_KEYS[i][31]
    k = 0
    for i in range(5):
        k |= (c[i] - 1) & 0xFFFF
    return ((k >> 8) & 1) > 0


# --------------------------------------------------------------------------- scalars
def scalar_is_canonical(sb: bytes) -> bool:
    """Scalar::is_canonical (scalar.rs:54-75): true iff the LE integer is < L."""
    if len(sb) != 32:
        return False
    return int.from_bytes(sb, "little") < L


def scalar_set_bytes(b: bytes) -> bytes:
    """Scalar::set_bytes (scalar.rs:175, integer.rs:386-396): LE integer of any length mod L."""
    return (int.from_bytes(b, "little") % L).to_bytes(32, "little")


def scalar_set_int64(v: int) -> bytes:
    """Scalar::set_int64 (scalar.rs:152); scalar_test.rs:38-46 pins -1 -> L-1."""
    return (v % L).to_bytes(32, "little")


def _sc(a: bytes) -> int:
    return int.from_bytes(a, "little")


def sc_mul_add(a: bytes, b: bytes, c: bytes) -> bytes:
    """scalar.rs:279 — (ab+c) mod L on raw 256-bit inputs."""
    return ((_sc(a) * _sc(b) + _sc(c)) % L).to_bytes(32, "little")


def sc_add(a: bytes, b: bytes) -> bytes:
    """scalar.rs:759."""
    return ((_sc(a) + _sc(b)) % L).to_bytes(32, "little")


def sc_sub(a: bytes, b: bytes) -> bytes:
    """scalar.rs:1187."""
    return ((_sc(a) - _sc(b)) % L).to_bytes(32, "little")


def sc_mul(a: bytes, b: bytes) -> bytes:
    """scalar.rs:1596."""
    return ((_sc(a) * _sc(b)) % L).to_bytes(32, "little")


def sc_neg(a: bytes) -> bytes:
    """Scalar::neg (scalar.rs:216)."""
    return ((-_sc(a)) % L).to_bytes(32, "little")


def sc_inv(a: bytes) -> bytes:
    """Scalar::inv (scalar.rs:192-214): a^(L-2)."""
    return pow(_sc(a) % L, L - 2, L).to_bytes(32, "little")


def scalar_marshal(a: bytes) -> bytes:
    """Scalar::marshal_binary (scalar.rs:91-100): reduces mod L."""
    return (_sc(a) % L).to_bytes(32, "little")


def clamp_key(seed: bytes):
    """Curve::new_key_and_seed_with_input (curve.rs:74-87): (unreduced clamped scalar, prefix)."""
    h = bytearray(hashlib.sha512(seed).digest())
    h[0] &= 0xF8
    h[31] &= 0x7F
    h[31] |= 0x40
    return bytes(h[:32]), bytes(h[32:])


# --------------------------------------------------------------------------- signatures
def challenge(r_bytes: bytes, a_bytes: bytes, msg: bytes) -> bytes:
    """H(R || A || M) -> Scalar::set_bytes (eddsa_sig.rs:195-200, schnorr_sig.rs:128-141)."""
    return scalar_set_bytes(hashlib.sha512(r_bytes + a_bytes + msg).digest())


def eddsa_sign(seed: bytes, msg: bytes) -> bytes:
    """EdDSA::sign (eddsa_sig.rs:120-152)."""
    a, prefix = clamp_key(seed)
    pk = point_encode(point_mul(a))
    r = scalar_set_bytes(hashlib.sha512(prefix + msg).digest())
    r_buf = point_encode(point_mul(r))
    h = challenge(r_buf, pk, msg)
    s = sc_add(r, sc_mul(a, h))
    return r_buf + scalar_marshal(s)


def eddsa_public(seed: bytes) -> bytes:
    a, _ = clamp_key(seed)
    return point_encode(point_mul(a))


def eddsa_verify(pk: bytes, msg: bytes, sig: bytes) -> int:
    """eddsa::verify_with_checks (eddsa_sig.rs:159-212): returns a status code in the
    reference's check order."""
    if len(sig) != 64:
        return ERR_SIG_LENGTH
    if not scalar_is_canonical(sig[32:]):
        return ERR_SIG_NOT_CANONICAL
    if not point_is_canonical(sig[:32]):
        return ERR_R_NOT_CANONICAL
    r = point_decode(sig[:32])
    if r is None:
        return ERR_MARSHALLING
    if point_has_small_order(r):
        return ERR_R_SMALL_ORDER
    if not point_is_canonical(pk):
        return ERR_PK_NOT_CANONICAL
    a = point_decode(pk)
    if a is None:
        return ERR_MARSHALLING
    if point_has_small_order(a):
        return ERR_PK_SMALL_ORDER
    h = challenge(sig[:32], pk, msg)
    s_b = point_mul(sig[32:])
    rha = point_add(r, point_mul(h, a))
    return OK if point_eq(rha, s_b) else ERR_INVALID_SIGNATURE


def schnorr_verify(pk: bytes, msg: bytes, sig: bytes) -> int:
    """schnorr::verify_with_checks (schnorr_sig.rs:53-110): same equation, different
    check order; the challenge hashes the RE-ENCODED R and A (schnorr_sig.rs:128-141)."""
    if len(sig) != 64:
        return ERR_SIG_LENGTH
    r = point_decode(sig[:32])
    if r is None:
        return ERR_MARSHALLING
    if not point_is_canonical(sig[:32]):
        return ERR_R_NOT_CANONICAL
    if point_has_small_order(r):
        return ERR_R_SMALL_ORDER
    if not scalar_is_canonical(sig[32:]):
        return ERR_SIG_NOT_CANONICAL
    a = point_decode(pk)
    if a is None:
        return ERR_MARSHALLING
    if not point_is_canonical(pk):
        return ERR_PK_NOT_CANONICAL
    if point_has_small_order(a):
        return ERR_PK_SMALL_ORDER
    h = challenge(point_encode(r), point_encode(a), msg)
    s_p = point_mul(sig[32:])
    ras = point_add(r, point_mul(h, a))
    return OK if point_eq(s_p, ras) else ERR_INVALID_SIGNATURE


def schnorr_sign(private: bytes, msg: bytes, k: bytes) -> bytes:
    """schnorr::sign (schnorr_sig.rs:25-47) with the nonce k supplied by the caller."""
    r = point_mul(k)
    public = point_mul(private)
    h = challenge(point_encode(r), point_encode(public), msg)
    s = sc_add(k, sc_mul(private, h))
    return point_encode(r) + scalar_marshal(s)


# --------------------------------------------------------------------------- polynomials
def pubpoly_eval(commits, i: int):
    """PubPoly::eval (poly.rs:457-469): Horner with FULL scalar mults by xi = 1+i."""
    xi = scalar_set_int64(1 + i)
    v = IDENTITY
    for c in reversed(commits):
        v = point_mul(xi, v)
        v = point_add(v, c)
    return v


def pubpoly_check(commits, i: int, share: bytes, base=None) -> bool:
    """PubPoly::check (poly.rs:526-530)."""
    return point_eq(pubpoly_eval(commits, i), point_mul(share, base))


def pubpoly_add(c1, c2):
    """PubPoly::add (poly.rs:486-509)."""
    assert len(c1) == len(c2)
    return [point_add(a, b) for a, b in zip(c1, c2)]


def pripoly_eval(coeffs, i: int) -> bytes:
    """PriPoly::eval (poly.rs:133-141)."""
    xi = scalar_set_int64(1 + i)
    v = bytes(32)
    for c in reversed(coeffs):
        v = sc_mul(v, xi)
        v = sc_add(v, c)
    return v


def pripoly_commit(coeffs, base=None):
    """PriPoly::commit (poly.rs:195-206)."""
    return [point_mul(c, base) for c in coeffs]


def vss_verify_deal(commits, i: int, share: bytes) -> bool:
    """Group math of vss::pedersen Aggregator::verify_deal (vss/pedersen/vss.rs:899-912):
    fi.v * B == PubPoly(commits).eval(fi.i) on canonical bytes."""
    return point_eq(point_mul(share), pubpoly_eval(commits, i))


def vss_rabin_verify_deal(commits, i: int, f_share: bytes, g_share: bytes, h_pt) -> bool:
    """vss::rabin verify_deal (vss/rabin/vss.rs:889-900): fi*G + gi*H == eval(i)."""
    ci = point_add(point_mul(f_share), point_mul(g_share, h_pt))
    return point_eq(ci, pubpoly_eval(commits, i))


def dss_verify_partial(random_commits, long_commits, i: int, partial: bytes, hash_scalar: bytes) -> bool:
    """Group math of DSS::process_partial_sig (dss_sig.rs:263-273)."""
    rand_share = pubpoly_eval(random_commits, i)
    long_share = pubpoly_eval(long_commits, i)
    right = point_add(rand_share, point_mul(hash_scalar, long_share))
    return point_eq(point_mul(partial), right)


def msm(scalars, points):
    """Sum_i Point::mul(s_i, P_i) folded with Point::add (point.rs:179,207)."""
    acc = (0, 1, 1, 0)
    for s, pt in zip(scalars, points):
        acc = _ext_add(acc, _to_ext(_mul_int(scalar_effective(s), pt)))
    return _from_ext(acc)


# --------------------------------------------------------------------------- interpolation in the exponent, session ids
def recover_commit(shares):
    """recover_commit (share/poly.rs:566-603) on the shares [(index, point)] already selected by xy_commit
    (:535-562: the first t in index order): sum_i (prod_{j!=i} x_j / (x_j - x_i)) * y_i with x = index + 1."""
    acc = IDENTITY
    xs = {i: scalar_set_int64(i + 1) for i, _ in shares}
    for i, yi in shares:
        num, den = scalar_set_int64(1), scalar_set_int64(1)
        for j, _ in shares:
            if i == j:
                continue
            num = sc_mul(num, xs[j])
            den = sc_mul(den, sc_sub(xs[j], xs[i]))
        lam = sc_mul(num, sc_inv(den))          # num.div(num, den)
        acc = point_add(acc, point_mul(lam, yi))
    return acc


def lagrange_basis(i, xs):
    """lagrange_basis (share/poly.rs:640-671): coefficients (low order first) of prod_{m!=i} (x - x_m) / (x_i - x_m)."""
    basis = [scalar_set_int64(1)]
    acc = scalar_set_int64(1)
    for m, xm in xs.items():
        if m == i:
            continue
        neg = sc_neg(xm)
        nxt = [bytes(32)] * (len(basis) + 1)
        for d, c in enumerate(basis):          # basis.mul(minus_const(xm))
            nxt[d] = sc_add(nxt[d], sc_mul(c, neg))
            nxt[d + 1] = sc_add(nxt[d + 1], c)
        basis = nxt
        acc = sc_mul(acc, sc_inv(sc_sub(xs[i], xm)))
    return [sc_mul(c, acc) for c in basis]


def recover_pub_poly(shares):
    """recover_pub_poly (share/poly.rs:607-635): sum_j L_j(x) * y_j in point space -> the commitments."""
    xs = {i: scalar_set_int64(i + 1) for i, _ in shares}
    acc = None
    for j, yj in shares:
        tmp = [point_mul(c, yj) for c in lagrange_basis(j, xs)]
        acc = tmp if acc is None else pubpoly_add(acc, tmp)
    return acc


def session_id(dealer, verifiers, commitments, t: int) -> bytes:
    """session_id (share/vss/pedersen/vss.rs:1069-1090): SHA-256 over the marshalled points and t as u32 LE
    (SuiteEd25519::hash is SHA-256, group/edwards25519/suite.rs:93-96)."""
    h = hashlib.sha256()
    h.update(point_encode(dealer))
    for v in verifiers:
        h.update(point_encode(v))
    for c in commitments:
        h.update(point_encode(c))
    h.update(int(t).to_bytes(4, "little"))
    return h.digest()


def find_pub(points, to_find):
    """find_pub (share/dkg/pedersen/dkg.rs:1109-1116): (index, found) of the first list entry equal to to_find."""
    for i, p in enumerate(points):
        if point_eq(p, to_find):
            return i, True
    return 0, False


def dss_hash_sig(random_commit0, long_commit0, msg: bytes) -> bytes:
    """hash_sig (sign/dss/dss_sig.rs:312-326): H(R || A || msg) as a scalar."""
    return challenge(point_encode(random_commit0), point_encode(long_commit0), msg)
