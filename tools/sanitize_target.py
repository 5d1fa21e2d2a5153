"""Tiny run of every entry point for compute-sanitizer (memcheck): n = 96 items each, results checked
against the oracle.  Usage: [KYBER_B200_LIB=...] python tools/sanitize_target.py"""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_c_oracle, load_sign_input, make_sig_batch, pack_batch  # noqa: E402
from oracle import ed25519_bigint as O  # noqa: E402

kb = importlib.import_module("kyber-rs_b200")
ctx = kb.Context(0)
C = load_c_oracle()
recs = load_sign_input()
n = 96
pks, msgs, sigs = make_sig_batch(recs[:200], n, bad_every=2)
pk, flat, off, sg = pack_batch(pks, msgs, sigs)
fails = []


def check(name, ok):
    print(name, "ok" if ok else "MISMATCH")
    if not ok:
        fails.append(name)


for sch in (False, True):
    check(f"verify schnorr={sch}", (ctx.verify_batch(pk, flat, off, sg, schnorr=sch) == C.verify_batch(pk, flat, off, sg, schnorr=sch)).all())
s = np.frombuffer(b"".join(r[0] for r in recs[:n]), dtype=np.uint8).reshape(-1, 32)
good = np.frombuffer(b"".join(r[1] for r in recs[:n]), dtype=np.uint8).reshape(-1, 32)
for fl in (0, 1):
    check(f"mul_base flags={fl}", (ctx.point_mul_base_batch(s, fl) == C.mul_base_batch(s)).all())
    check(f"mul flags={fl}", (ctx.point_mul_batch(s, good, fl)[0] == C.mul_batch(s, good)).all())
check("recode", (ctx.point_recode_batch(good)[0] == good).all())
check("add", ctx.point_add_batch(good[:8], good[8:16])[0][0].tobytes() == C.point_add(good[0].tobytes(), good[8].tobytes()))
ctx.point_check_batch(pk)
t = 5
coeffs = [O.scalar_set_bytes(hashlib.sha512(b"c%d" % j).digest()) for j in range(t)]
commits = ctx.point_mul_base_batch(np.frombuffer(b"".join(coeffs), dtype=np.uint8).reshape(-1, 32))
idx = np.arange(n, dtype=np.uint32)
ev, st = ctx.pubpoly_eval_batch(commits, t, np.zeros(n, dtype=np.uint32), idx)
check("eval", all(ev[k].tobytes() == C.pubpoly_eval([c.tobytes() for c in commits], k) for k in range(0, n, 7)))
shares = np.frombuffer(b"".join(O.pripoly_eval(coeffs, k) for k in range(n)), dtype=np.uint8).reshape(-1, 32).copy()
shares[5, 0] ^= 1
v = ctx.vss_verify_deals_batch(commits, t, np.zeros(n, dtype=np.uint32), idx, shares)
check("vss", v.sum() == n - 1 and v[5] == 0)
v2 = ctx.dkg_verify_round(n, t, commits, shares)
check("dkg", (v2 == v).all())
enc, bad = ctx.msm(s, good)
check("msm", bad == 0 and enc == C.msm(s, good))
print("FAILS", fails)
sys.exit(1 if fails else 0)
