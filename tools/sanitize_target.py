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
# ---- round-2 kernels: forward-difference round (all block counts), protocol-level entry points, decoded-point MSM,
# multi-device context on one device
import torch  # noqa: E402

nd, nn, tt = 40, 70, 67
polys = np.concatenate([np.frombuffer(b"".join(O.scalar_set_bytes(hashlib.sha512(b"s%d/%d" % (d, j)).digest()) for j in range(tt)), dtype=np.uint8).reshape(-1, 32) for d in range(nd)])
cm = ctx.point_mul_base_batch(polys, 1)
sh = ctx.pripoly_eval_batch(polys, tt, nn)
sh[11, 2] ^= 4
want = np.ones(nd * nn, dtype=np.uint8)
want[11] = 0
for parts in ("1", "2", "3", "4"):
    os.environ["KB_DKG_FD"], os.environ["KB_FD_PARTS"] = "1", parts
    c2 = kb.Context(0)
    del os.environ["KB_DKG_FD"], os.environ["KB_FD_PARTS"]
    check(f"dkg forward differences, {parts} blocks", (c2.dkg_verify_round(nn, tt, cm, sh) == want).all())
    limbs = np.stack([C.point_limbs(c.tobytes()) for c in cm])
    check(f"dkg forward differences, {parts} blocks, limbs", (c2.dkg_verify_round(nn, tt, limbs, sh, limbs=True) == want).all())
    c2.close()
sid, st = ctx.vss_session_ids(cm[:3], cm[3:10], cm[10:10 + 3 * 5], 5)
check("session ids", not st.any() and sid[0].tobytes() == hashlib.sha256(cm[0].tobytes() + cm[3:10].tobytes() + cm[10:15].tobytes() + (5).to_bytes(4, "little")).digest())
check("find_pub", ctx.find_pub_batch(cm[:50], cm[[7, 49, 60]]).tolist() == [7, 49, -1])
H = ctx.point_mul_base_batch(s[:1], 1)[0]
check("rabin", (ctx.vss_rabin_verify_deals_batch(cm[:tt], tt, H, np.zeros(8, dtype=np.uint32), np.arange(8, dtype=np.uint32), sh[:8], sh[8:16]) == C.rabin_verify_batch(cm[:tt], H.tobytes(), np.arange(8), sh[:8], sh[8:16])).all())
vd, hs = ctx.dss_verify_partials(cm[:tt], cm[tt:2 * tt], b"m", np.arange(6, dtype=np.uint32), sh[:6])
check("dss", hs == C.dss_hash_sig(cm[:tt], cm[tt:2 * tt], b"m") and (vd == C.dss_partial_batch(cm[:tt], cm[tt:2 * tt], b"m", np.arange(6), sh[:6])).all())
pick = np.array([0, 2, 3, 5, 9], dtype=np.uint32)
pub = ctx.point_mul_base_batch(sh[pick], 1)       # shares 0.. of dealer 0 (a degree-66 polynomial: only the oracle comparison is meaningful)
check("recover_commit", ctx.recover_commit_batch(pick, pub)[0][0].tobytes() == C.recover_commit(pick, pub))
check("recover_pub_poly", (ctx.recover_pub_poly(pick, pub)[0] == C.recover_pub_poly(pick, pub)).all())
ctx.dkg_resharing_key(3, pick, np.tile(pub, (3, 1)).reshape(3, 5, 32).transpose(1, 0, 2).reshape(-1, 32), 1, sh[1])
raw, _ = ctx.point_decompress_batch(good)
dev = torch.device("cuda", 0)
d_out = torch.zeros(32, dtype=torch.uint8, device=dev)
d_bad = torch.zeros(1, dtype=torch.int64, device=dev)
ctx.dev_msm_ext(n, torch.from_numpy(s.copy()).to(dev), torch.from_numpy(raw.view(np.uint8).reshape(n, 128).copy()).to(dev), d_out, None, d_bad)
torch.cuda.synchronize()
check("msm over decoded points", d_out.cpu().numpy().tobytes() == enc)
m = kb.MultiContext([0])
check("mctx verify", (m.verify_batch(pk, flat, off, sg) == C.verify_batch(pk, flat, off, sg)).all())
check("mctx msm", m.msm(s, good)[0] == enc)
m.close()
print("FAILS", fails)
sys.exit(1 if fails else 0)
