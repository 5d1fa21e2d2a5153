"""Out-of-bounds WRITE detection for the device-pointer entry points (compute-sanitizer is closed on the GPU pool):
every output buffer is a window inside a larger allocation filled with a sentinel; after the call the bytes on both
sides of the window must be untouched.  Sizes are deliberately odd (not multiples of the block size, of the 8-item
inversion groups or of the 32-dealer warps)."""
import hashlib
import importlib

import numpy as np
import pytest

from helpers import make_sig_batch, pack_batch
from oracle import ed25519_bigint as O

pytestmark = pytest.mark.gpu
GUARD = 4096


@pytest.fixture(scope="module")
def kb():
    return importlib.import_module("kyber-rs_b200")


@pytest.fixture(scope="module")
def ctx(kb):
    c = kb.Context(0)
    yield c
    c.close()


class Guarded:
    def __init__(self, torch, dev, nbytes):
        self.torch = torch
        pad = (-nbytes) % 16
        self.buf = torch.full((GUARD + nbytes + pad + GUARD,), 0xA5, dtype=torch.uint8, device=dev)
        self.view = self.buf[GUARD:GUARD + nbytes]
        self.n = nbytes

    def intact(self):
        self.torch.cuda.synchronize()
        b = self.buf.cpu().numpy()
        return bool((b[:GUARD] == 0xA5).all() and (b[GUARD + self.n:] == 0xA5).all())


def _scalars(tag, n):
    return np.frombuffer(b"".join(O.scalar_set_bytes(hashlib.sha512(tag + b"/%d" % k).digest()) for k in range(n)), dtype=np.uint8).reshape(n, 32).copy()


def test_outputs_stay_inside_their_buffers(kb, ctx, golden_records):
    import torch

    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    n = 1003
    pks, msgs, sigs = make_sig_batch(golden_records[:300], n, bad_every=5)
    pk, flat, off, sg = pack_batch(pks, msgs, sigs)
    for schnorr in (False, True):
        st = Guarded(torch, dev, n)
        ctx.dev_verify(n, up(pk), up(flat), up(off.view(np.int64)), up(sg), st.view, schnorr=schnorr)
        assert st.intact()
    s = _scalars(b"guard", n)
    for flags in (0, 1):
        o = Guarded(torch, dev, 32 * n)
        ctx.dev_point_mul_base(n, up(s), o.view, flags)
        assert o.intact()
        o2, s8 = Guarded(torch, dev, 32 * n), Guarded(torch, dev, n)
        ctx.dev_point_mul(n, up(s), o.view.clone().reshape(n, 32), o2.view, s8.view, flags)
        assert o2.intact() and s8.intact()
    pts = ctx.point_mul_base_batch(s, 1)
    raw, s8 = Guarded(torch, dev, 128 * n), Guarded(torch, dev, n)
    ctx.dev_point_decompress(n, up(pts), raw.view, s8.view)
    assert raw.intact() and s8.intact()
    for ext in (False, True):
        enc, part, bad = Guarded(torch, dev, 32), Guarded(torch, dev, 128), Guarded(torch, dev, 8)
        if ext:
            ctx.dev_msm_ext(n, up(s), raw.view.clone(), enc.view, part.view, bad.view)
        else:
            ctx.dev_msm(n, up(s), up(pts), enc.view, part.view, bad.view)
        assert enc.intact() and part.intact() and bad.intact()
    h = Guarded(torch, dev, 32 * n)
    ctx.dev_challenge(n, up(sg[:, :32]), up(pk), up(flat), up(off.view(np.int64)), h.view)
    assert h.intact()
    # DKG rounds: dealers and verifiers that fill neither a warp nor a block, every block count, + the whole round
    nv, t, nd = 37, 67, 45
    polys = _scalars(b"guard-poly", nd * t)
    commits = ctx.point_mul_base_batch(polys, 1)
    sh = Guarded(torch, dev, 32 * nd * nv)
    ctx.dev_pripoly_eval(nd, t, up(polys), nv, sh.view)
    assert sh.intact()
    import os

    for parts in ("0", "1", "2", "3", "4"):
        os.environ["KB_DKG_FD"] = "1"
        if parts != "0":
            os.environ["KB_FD_PARTS"] = parts
        try:
            c2 = kb.Context(0)
        finally:
            del os.environ["KB_DKG_FD"]
            os.environ.pop("KB_FD_PARTS", None)
        v = Guarded(torch, dev, nd * nv)
        c2.dev_dkg_verify_round(nv, t, nd, up(commits), sh.view, v.view)
        assert v.intact() and bool(v.view.all())
        c2.close()
    m = nd * nv
    dpk, dmsg, doff, dsg = pack_batch(*make_sig_batch(golden_records[:100], m, bad_every=9))
    v, ds = Guarded(torch, dev, m), Guarded(torch, dev, m)
    ctx.dev_dkg_process_round(nv, t, nd, up(commits), sh.view, v.view, deal=(up(dpk), up(dmsg), up(doff.view(np.int64)), up(dsg), ds.view))
    assert v.intact() and ds.intact()
