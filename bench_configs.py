"""bench_configs.py — BASELINE configs 1, 3, 4, 5 and the stage measurements of bench.py (imported by it; one process
per GPU under torchrun).  Every configuration carries its own parity check against the oracle or against verdicts
expected by construction; a mismatch on any rank ends the run with a non-zero exit code (bench.die).

Units of work and their charges are SURVEY §8(d)'s; `frac` is always against the architectural IMAD.WIDE rate of the
GPUs used (bench.py).  Strong-scaling shapes: the work of a configuration is FIXED and split over the ranks."""
import time

import numpy as np

from bench import (IMAD_EQ_PER_BASE_MUL, IMAD_EQ_PER_DECOMPRESS, IMAD_EQ_PER_EVAL_COEFF, IMAD_EQ_PER_MSM_POINT, IMAD_EQ_PER_VAR_MUL, L_ORDER, T8)


def _shard(n, rank, world):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _roof(units_per_s, charge, env):
    ach = units_per_s * charge
    return {"bound": "imad", "achieved": ach / 1e12, "peak": env["imad_peak"] * env["world"] / 1e12, "unit": "T IMAD-eq/s", "frac": ach / (env["imad_peak"] * env["world"]), "charged_imad_eq_per_unit": charge}


# ------------------------------------------------------------------------------------------------------------------------
def cfg1(env):
    """benches/ed25519.rs Point::mul: 2^16 scalars, base point and variable base (distinct points and ONE shared point —
    what util/test/group.rs:159-178 benches), constant-time and public-scalar paths; split by index over the ranks."""
    torch, ctx, dev, rank, world, C = env["torch"], env["ctx"], env["dev"], env["rank"], env["world"], env["C"]
    m_total = 1 << 16
    lo, hi = _shard(m_total, rank, world)
    m = hi - lo
    sc = env["xof"]("kyber-b200/cfg1/scalars", 32 * m_total).reshape(-1, 32).copy()
    sc[:, 31] &= 0x0F
    ps = env["xof"]("kyber-b200/cfg1/point-scalars", 32 * m_total).reshape(-1, 32).copy()
    ps[:, 31] &= 0x0F
    d_sc = torch.from_numpy(sc[lo:hi]).to(dev)
    d_ps = torch.from_numpy(ps[lo:hi]).to(dev)
    d_pts = torch.empty(m, 32, dtype=torch.uint8, device=dev)
    ctx.dev_point_mul_base(m, d_ps, d_pts, 1)
    d_o = torch.empty(m, 32, dtype=torch.uint8, device=dev)
    d_s8 = torch.empty(m, dtype=torch.uint8, device=dev)
    out = {"workload": "cfg1: Point::mul x 2^16 scalars (benches/ed25519.rs), split by index", "scaling": "strong", "unit": "mults/s", "variants": {}}
    runs = {"base_ct": (lambda: ctx.dev_point_mul_base(m, d_sc, d_o, 0), IMAD_EQ_PER_BASE_MUL),
            "base_vartime": (lambda: ctx.dev_point_mul_base(m, d_sc, d_o, 1), IMAD_EQ_PER_BASE_MUL),
            "var_ct": (lambda: ctx.dev_point_mul(m, d_sc, d_pts, d_o, d_s8, 0), IMAD_EQ_PER_VAR_MUL),
            "var_vartime": (lambda: ctx.dev_point_mul(m, d_sc, d_pts, d_o, d_s8, 1), IMAD_EQ_PER_VAR_MUL),
            "var_ct_shared_point": (lambda: ctx.dev_point_mul(m, d_sc, d_pts, d_o, d_s8, 2), IMAD_EQ_PER_VAR_MUL)}
    sample = 2048 if rank == 0 else 0
    for name, (fn, charge) in runs.items():
        s = env["timed"](fn, reps=5)
        out["variants"][name] = {"value": m_total / s, "ms": s * 1e3, "roofline": _roof(m_total / s, charge, env)}
        if rank == 0:   # parity on a sample of rank 0's slice, every variant
            got = d_o[:sample].cpu().numpy()
            pts = d_pts[:sample].cpu().numpy()
            if name.startswith("base"):
                want = C.mul_base_batch(sc[lo:lo + sample], nthreads=env["host_cores"])
            elif name.endswith("shared_point"):
                want = C.mul_batch(sc[lo:lo + sample], np.repeat(pts[:1], sample, axis=0), nthreads=env["host_cores"])
            else:
                want = C.mul_batch(sc[lo:lo + sample], pts, nthreads=env["host_cores"])
            if not (got == want).all():
                env["die"](f"cfg1 {name}: encodings differ from the oracle")
    env["parity"].append({"what": "cfg1 Point::mul encodings (5 variants) vs oracle/ref10_port.c on rank 0's slice", "items": 5 * sample})
    out["value"] = out["variants"]["base_ct"]["value"]
    out["roofline"] = out["variants"]["base_ct"]["roofline"]
    if rank == 0:
        if world == 1:   # the reference's algorithm on the host cores, the whole configuration
            th = env["host_cores"]
            t0 = time.perf_counter(); C.mul_base_batch(sc, nthreads=th); tb = time.perf_counter() - t0
            pts_all = d_pts.cpu().numpy()
            t0 = time.perf_counter(); C.mul_batch(sc, pts_all, nthreads=th); tv = time.perf_counter() - t0
            out["cpu_baseline"] = {"base_mults_per_s": m_total / tb, "var_mults_per_s": m_total / tv, "unit": "mults/s", "cores": th, "kind": "port", "sample": "all 2^16 scalars, oracle/ref10_port.c (ge_scalarmult_base / ge_scalarmult as ge.rs:442, :508)"}
        mctx = env["mctx"]
        mctx.point_mul_base_batch(sc)
        t0 = time.perf_counter()
        for _ in range(5):
            mctx.point_mul_base_batch(sc)
        tb = (time.perf_counter() - t0) / 5
        pts_all = mctx.point_mul_base_batch(ps, 1)
        mctx.point_mul_batch(sc, pts_all)
        t0 = time.perf_counter()
        for _ in range(5):
            mctx.point_mul_batch(sc, pts_all)
        tv = (time.perf_counter() - t0) / 5
        out["e2e"] = {"base_ct_mults_per_s": m_total / tb, "var_ct_mults_per_s": m_total / tv, "unit": "mults/s", "how": "kb_mctx_point_mul_base_batch / kb_mctx_point_mul_batch on host buffers (2 MB in, 2 MB out), all GPUs"}
    env["host_barrier"]()
    return out


# ------------------------------------------------------------------------------------------------------------------------
def build_round(env, ctx, n, t, d_lo, d_hi, tag):
    """Dealers [d_lo, d_hi) of a deal-verification round, on the device of ctx: commitments = PriPoly::commit of
    per-dealer coefficients (BLAKE3-XOF), shares = PriPoly::eval for every verifier (kb_dev_pripoly_eval — honest shares
    for EVERY dealer), about 0.1 % of the shares corrupted, dealer 4 with a torsion-contaminated commitment (SURVEY §7-H2:
    its checks then only pass where 8 | x).  Returns device tensors and the verdicts expected by construction."""
    torch, dev = env["torch"], torch_dev(env, ctx)
    nd = d_hi - d_lo
    coeff = np.concatenate([env["xof"](f"kyber-b200/{tag}/dealer{d}", 32 * t).reshape(t, 32) for d in range(d_lo, d_hi)]).copy()
    coeff[:, 31] &= 0x0F
    with torch.cuda.device(dev):
        d_coeff = torch.from_numpy(coeff).to(dev)
        d_commits = torch.empty(nd * t, 32, dtype=torch.uint8, device=dev)
        ctx.dev_point_mul_base(nd * t, d_coeff, d_commits, 1)
        d_shares = torch.empty(nd * n, 32, dtype=torch.uint8, device=dev)
        ctx.dev_pripoly_eval(nd, t, d_coeff, n, d_shares)
        torch.cuda.synchronize(dev)
        expect = np.ones((nd, n), dtype=np.uint8)
        flat = np.arange(d_lo * n, d_hi * n)
        bad = flat[flat % 997 == 0]
        if bad.size:
            idx = torch.from_numpy(bad - d_lo * n).to(dev)
            d_shares[idx, 3] ^= 0x10
            expect.reshape(-1)[bad - d_lo * n] = 0
        if d_lo <= 4 < d_hi and t > 1:
            k = (4 - d_lo) * t + 1
            c1 = d_commits[k].cpu().numpy()
            summed = ctx.point_add_batch(c1, np.frombuffer(T8, dtype=np.uint8))[0][0]
            d_commits[k] = torch.from_numpy(summed).to(dev)
            row = np.zeros(n, dtype=np.uint8)
            row[7::8] = 1
            expect[4 - d_lo] &= row
    return coeff, d_commits, d_shares, expect


def torch_dev(env, ctx):
    return env["torch"].device("cuda", ctx.device)


def sign_batch(env, ctx, tag, m, msg_len):
    """m Schnorr signatures (schnorr_sig.rs:25-47) over random msg_len-byte messages, made with the library's batched
    primitives; every 128th is invalid (flipped message bit -> status 8).  Returns host arrays + expected statuses."""
    x = env["xof"](f"kyber-b200/{tag}/x", 32 * m).reshape(m, 32).copy()
    x[:, 31] &= 0x0F
    k = env["xof"](f"kyber-b200/{tag}/k", 32 * m).reshape(m, 32).copy()
    k[:, 31] &= 0x0F
    msg = env["xof"](f"kyber-b200/{tag}/m", msg_len * m).copy()
    off = np.arange(m + 1, dtype=np.uint64) * np.uint64(msg_len)
    pub = ctx.point_mul_base_batch(x)
    r = ctx.point_mul_base_batch(k)
    h = ctx.challenge_batch(r, pub, msg, off)
    s = ctx.sc_muladd_batch(x, h, k)
    sig = np.concatenate([r, s], axis=1)
    expect = np.zeros(m, dtype=np.uint8)
    bad = np.arange(127, m, 128)
    msg[bad * msg_len] ^= 1
    expect[bad] = 8
    return pub, msg, off, sig, expect


def dkg_config(env, name, n, t, with_signatures):
    """BASELINE configs 3 / 4: ALL n^2 share checks of a round, the n dealers split over the ranks (strong scaling);
    with_signatures adds the 2 n^2 Schnorr verifications of deal and response signatures (the whole round of
    share/dkg/pedersen/dkg.rs:513-597 + share/vss/pedersen/vss.rs:931-946)."""
    torch, ctx, dev, rank, world, C = env["torch"], env["ctx"], env["dev"], env["rank"], env["world"], env["C"]
    lo, hi = _shard(n, rank, world)
    nd = hi - lo
    coeff, d_commits, d_shares, expect = build_round(env, ctx, n, t, lo, hi, name)
    d_v = torch.zeros(nd * n, dtype=torch.uint8, device=dev)
    ctx.dev_dkg_verify_round(n, t, nd, d_commits, d_shares, d_v)
    torch.cuda.synchronize()
    got = d_v.cpu().numpy().reshape(nd, n)
    env["all_ok"](bool((got == expect).all()), f"{name}: rank {rank} verdicts differ from the expected ones at {np.argwhere(got != expect)[:5].tolist()}")
    env["parity"].append({"what": f"{name}: all n^2 share-check verdicts vs expectation (honest shares for every dealer, ~0.1 % corrupted, a torsion-contaminated dealer)", "items": n * n, "ranks": world})
    if rank == 0:   # the reference's own evaluation (t full scalar mults per check) on a few (dealer, verifier) pairs
        commits_h = d_commits.cpu().numpy()
        shares_h = d_shares.cpu().numpy()
        picks = [(0, 0), (0, n - 1), (min(4, nd - 1), 6), (min(4, nd - 1), 7), (nd - 1, n // 2)]
        for d, i in picks:
            want = C.vss_verify_deal([c.tobytes() for c in commits_h[d * t:(d + 1) * t]], i, shares_h[d * n + i].tobytes())
            if want != int(got[d, i]):
                env["die"](f"{name}: verdict of dealer {d}, verifier {i} differs from the oracle")
        env["parity"].append({"what": f"{name}: verdicts vs oracle/ref10_port.c PubPoly::eval (t full scalar mults each)", "items": len(picks)})
    s = env["timed"](lambda: ctx.dev_dkg_verify_round(n, t, nd, d_commits, d_shares, d_v), reps=3)
    checks = n * n
    out = {"workload": f"{name}: deal verification n={n}, t={t}: all n^2 share checks (PubPoly::eval + share*B + compare), dealers split over the ranks", "scaling": "strong",
           "value": checks / s, "unit": "share checks/s", "round_ms": s * 1e3, "roofline": _roof(checks / s, t * IMAD_EQ_PER_EVAL_COEFF, env),
           "note": "charged t x 6.8k IMAD-eq per check (SURVEY 8d: one short-scalar Horner run per share); the forward-difference round (csrc/dkgfd.cuh) executes about 5x fewer, so frac can exceed 1"}
    sig_inputs = None
    if with_signatures:
        m = nd * n
        deal = sign_batch(env, ctx, f"{name}/deal-sigs/rank{rank}", m, 128)
        resp = sign_batch(env, ctx, f"{name}/resp-sigs/rank{rank}", m, 32)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dd = [up(deal[0]), up(deal[1]), up(deal[2].view(np.int64)), up(deal[3]), torch.zeros(m, dtype=torch.uint8, device=dev)]
        dr = [up(resp[0]), up(resp[1]), up(resp[2].view(np.int64)), up(resp[3]), torch.zeros(m, dtype=torch.uint8, device=dev)]
        full = lambda: ctx.dev_dkg_process_round(n, t, nd, d_commits, d_shares, d_v, deal=dd, resp=dr)
        full(); torch.cuda.synchronize()
        ok = bool((d_v.cpu().numpy().reshape(nd, n) == expect).all() and (dd[4].cpu().numpy() == deal[4]).all() and (dr[4].cpu().numpy() == resp[4]).all())
        env["all_ok"](ok, f"{name}: whole round (share checks + deal / response signatures) differs from the expected verdicts / statuses on rank {rank}")
        if rank == 0:
            mm = min(m, 4096)
            if not (C.verify_batch(deal[0][:mm], deal[1][:128 * mm], deal[2][:mm + 1], deal[3][:mm], nthreads=env["host_cores"], schnorr=True) == deal[4][:mm]).all():
                env["die"](f"{name}: deal-signature statuses differ from the oracle")
        env["parity"].append({"what": f"{name}: whole round = share checks + n^2 deal signatures (128-byte messages) + n^2 response signatures (32-byte messages) vs expectation; 4096 deal signatures vs the oracle", "items": 3 * n * n})
        s2 = env["timed"](full, reps=3)
        out["whole_round"] = {"round_ms": s2 * 1e3, "schnorr_verifies": 2 * n * n, "share_checks": n * n,
                              "note": "deal-signature messages are 128 synthetic bytes: the reference signs bincode(Deal), which at t=683 is ~110 KB (683 commitments x 161 B) per deal — hashing that is "
                                      "860 SHA-512 blocks per signature, see stages.hash for the device's SHA-512 rate; response messages are the reference's 32-byte Response::hash"}
        sig_inputs = (deal, resp)
        del dd, dr
    # e2e: the whole round from HOST buffers through the multi-device context on rank 0 (it rebuilds all n dealers)
    env["host_barrier"]()
    if rank == 0:
        mctx = env["mctx"]
        _, dc_all, ds_all, exp_all = build_round(env, ctx, n, t, 0, n, name)
        # pinned host buffers, as for the headline's end-to-end number (pageable memory would add a staging copy)
        commits_all, shares_all = dc_all.cpu().pin_memory().numpy(), ds_all.cpu().pin_memory().numpy()
        del dc_all, ds_all
        v = torch.zeros(n * n, dtype=torch.uint8).pin_memory().numpy()
        mctx.dkg_verify_round(n, t, commits_all, shares_all, verdict=v)
        samples = []
        for _ in range(3):   # wall clock around the call; the median of three
            t0 = time.perf_counter()
            mctx.dkg_verify_round(n, t, commits_all, shares_all, verdict=v)
            samples.append(time.perf_counter() - t0)
        e2e = sorted(samples)[1]
        if not (v.reshape(n, n) == exp_all).all():
            env["die"](f"{name}: e2e verdicts (multi-device context) differ from the expected ones")
        out["e2e"] = {"round_ms": e2e * 1e3, "value": checks / e2e, "unit": "share checks/s", "h2d_bytes": int(commits_all.nbytes + shares_all.nbytes), "d2h_bytes": int(v.nbytes),
                      "how": "kb_mctx_dkg_verify_round on pinned host buffers, dealers split over all GPUs inside the library"}
        env["parity"].append({"what": f"{name}: e2e kb_mctx_dkg_verify_round verdicts vs expectation", "items": n * n, "devices": world})
        if world == 1:
            # CPU baseline: the reference's per-share evaluation on the host cores, a bounded sample of the checks
            th = env["host_cores"]
            per = max(1, int(4.0 / (t * 55e-6)))          # ~4 s per thread at ~55 us per full scalar mult
            cnt = min(n, th * per)
            idx = np.arange(cnt, dtype=np.uint32) % n
            t0 = time.perf_counter()
            row = C.vss_verify_batch(commits_all[:t], idx, shares_all[idx], nthreads=th)
            dt = time.perf_counter() - t0
            if not (row == exp_all[0, idx]).all():
                env["die"](f"{name}: CPU-baseline verdicts differ")
            out["cpu_baseline"] = {"value": cnt / dt, "unit": "share checks/s", "cores": th, "kind": "port", "sample": f"{cnt} of the {checks} share checks (dealer 0), oracle/ref10_port.c: PubPoly::eval = t full constant-time scalar mults as poly.rs:457-469",
                                   "round_s_extrapolated": checks / (cnt / dt), "extrapolated": True}
    env["host_barrier"]()
    return out


# ------------------------------------------------------------------------------------------------------------------------
def cfg5(env):
    """Pippenger MSM sweep, 2^16 .. 2^26 TOTAL points split over the ranks; each rank reduces its points to one 128-byte
    partial, all_gather (NCCL) of the partials, every rank folds them."""
    torch, ctx, dev, rank, world, C, dist = env["torch"], env["ctx"], env["dev"], env["rank"], env["world"], env["C"], env["dist"]
    max_log2 = env["args"].msm_max_log2
    total_max = 1 << max_log2
    lo, hi = _shard(total_max, rank, world)
    per = hi - lo
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    CH = 1 << 22     # generated and multiplied in chunks: bounded scratch, no 2^31-element tensor op
    d_sc = torch.empty(per, 32, dtype=torch.uint8, device=dev)
    d_pts = torch.empty(per, 32, dtype=torch.uint8, device=dev)
    for c0 in range(0, per, CH):
        c1 = min(per, c0 + CH)
        d_sc[c0:c1] = torch.randint(0, 256, (c1 - c0, 32), dtype=torch.uint8, device=dev, generator=g)
        d_ps = torch.randint(0, 256, (c1 - c0, 32), dtype=torch.uint8, device=dev, generator=g)
        d_ps[:, 31] &= 0x0F
        ctx.dev_point_mul_base(c1 - c0, d_ps, d_pts[c0:c1], 1)
        torch.cuda.synchronize()
    d_sc[:, 31] &= 0x0F
    d_part = torch.empty(128, dtype=torch.uint8, device=dev)
    d_all = torch.empty(world * 128, dtype=torch.uint8, device=dev)
    d_enc = torch.empty(32, dtype=torch.uint8, device=dev)
    d_bad = torch.zeros(1, dtype=torch.int64, device=dev)

    def msm(cnt, sc=None, pts=None):
        ctx.dev_msm(cnt, d_sc if sc is None else sc, d_pts if pts is None else pts, None, d_part, d_bad)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_part)
            ctx.dev_point_sum(world, d_all, d_enc)
        else:
            ctx.dev_point_sum(1, d_part, d_enc)

    def result():
        torch.cuda.synchronize()
        return bytes(d_enc.cpu().numpy().tobytes())

    # parity: 2^14 total points, sharded, against the oracle's fold of Point::mul + Point::add on rank 0
    tot = 1 << 14
    plo, phi = _shard(tot, rank, world)
    cnt = phi - plo
    msm(cnt)
    enc = result()
    sc_h, pts_h = d_sc[:cnt].cpu().numpy(), d_pts[:cnt].cpu().numpy()
    if world > 1:
        gs, gp = [None] * world, [None] * world
        dist.all_gather_object(gs, sc_h)
        dist.all_gather_object(gp, pts_h)
        encs = [None] * world
        dist.all_gather_object(encs, enc)
        if len(set(encs)) != 1:
            env["die"]("cfg5: ranks disagree on the sharded MSM result")
        sc_h, pts_h = np.concatenate(gs), np.concatenate(gp)
    if rank == 0 and enc != C.msm(sc_h, pts_h):
        env["die"]("cfg5: sharded MSM of 2^14 points differs from the oracle")
    env["parity"].append({"what": "cfg5: MSM of 2^14 points split over the ranks (NCCL all_gather of the partials + fold) vs oracle/ref10_port.c", "items": tot, "ranks": world})
    out = {"workload": "cfg5: Pippenger MSM sweep over TOTAL points, split by points over the ranks; NCCL all_gather of 128-byte partials + fold", "scaling": "strong", "unit": "points/s", "sweep": {}}
    best = 0.0
    for lg in range(16, max_log2 + 1, 2):
        tot = 1 << lg
        plo, phi = _shard(tot, rank, world)
        cnt = phi - plo
        reps = 3 if lg <= 22 else 2
        s = env["timed"](lambda: msm(cnt), reps=reps)
        out["sweep"][f"2^{lg}"] = {"value": tot / s, "ms": s * 1e3, "frac": tot / s * IMAD_EQ_PER_MSM_POINT / (env["imad_peak"] * world)}
        best = max(best, tot / s)
    out["value"] = best
    out["roofline"] = _roof(best, IMAD_EQ_PER_MSM_POINT, env)
    # the same sweep over points that are ALREADY DECODED (kb_dev_msm_ext: what chained device-side use hands over)
    lg_raw = min(max_log2, 24)
    raw_cnt = _shard(1 << lg_raw, rank, world)
    raw_cnt = raw_cnt[1] - raw_cnt[0]
    d_raw = torch.empty(raw_cnt, 128, dtype=torch.uint8, device=dev)
    d_s8 = torch.empty(raw_cnt, dtype=torch.uint8, device=dev)
    ctx.dev_point_decompress(raw_cnt, d_pts, d_raw, d_s8)

    def msm_raw(cnt):
        ctx.dev_msm_ext(cnt, d_sc, d_raw, None, d_part, d_bad)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_part)
            ctx.dev_point_sum(world, d_all, d_enc)
        else:
            ctx.dev_point_sum(1, d_part, d_enc)

    out["sweep_decoded_points"] = {}
    for lg in range(16, lg_raw + 1, 2):
        tot = 1 << lg
        plo, phi = _shard(tot, rank, world)
        cnt = phi - plo
        msm(cnt)
        want = result()
        msm_raw(cnt)
        env["all_ok"](result() == want, f"cfg5: MSM over decoded points differs from the MSM over their encodings at 2^{lg}")
        s = env["timed"](lambda: msm_raw(cnt), reps=3 if lg <= 22 else 2)
        out["sweep_decoded_points"][f"2^{lg}"] = {"value": tot / s, "ms": s * 1e3}
    env["parity"].append({"what": "cfg5: kb_dev_msm_ext (decoded points) vs kb_dev_msm (encodings), every size of the sweep", "items": (1 << lg_raw), "ranks": world})
    del d_raw
    env["host_barrier"]()
    if rank == 0:
        mctx = env["mctx"]
        tot = 1 << min(22, max_log2)
        g0 = torch.Generator(device=dev)
        g0.manual_seed(99)
        hs = torch.randint(0, 256, (tot, 32), dtype=torch.uint8, device=dev, generator=g0)
        hs[:, 31] &= 0x0F
        hp = torch.empty(tot, 32, dtype=torch.uint8, device=dev)
        ctx.dev_point_mul_base(tot, hs, hp, 1)
        h_s, h_p = hs.cpu().pin_memory().numpy(), hp.cpu().pin_memory().numpy()
        ref_enc, _ = ctx.msm(h_s, h_p)
        e1, b1 = mctx.msm(h_s, h_p)
        t0 = time.perf_counter()
        e2, b2 = mctx.msm(h_s, h_p)
        dt = time.perf_counter() - t0
        if e1 != ref_enc or e2 != ref_enc or b1 or b2:
            env["die"]("cfg5: e2e MSM through the multi-device context differs from the single-device result")
        out["e2e"] = {"value": tot / dt, "unit": "points/s", "points": tot, "ms": dt * 1e3, "h2d_bytes": int(64 * tot), "d2h_bytes": 40,
                      "how": "kb_mctx_msm on pinned host buffers: points split over all GPUs, ncclAllGather of the partials and fold inside the library"}
        env["parity"].append({"what": "cfg5: kb_mctx_msm (NCCL inside the library) vs the single-device kb_msm", "items": tot, "devices": world})
        if world == 1:
            th = env["host_cores"]
            cnt = 1 << 14
            t0 = time.perf_counter()
            C.mul_batch(sc_h[:cnt], pts_h[:cnt], nthreads=th)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": cnt / dt, "unit": "points/s", "cores": th, "kind": "port",
                                   "sample": f"{cnt} points: the reference's fold is one Point::mul (ge_scalar_mult, ge.rs:508) and one Point::add per point (point.rs:179,207); timed as the multiplications on all host threads",
                                   "extrapolated": True}
    env["host_barrier"]()
    return out


# ------------------------------------------------------------------------------------------------------------------------
def stage_rates(env):
    """The pure decompress and hash stages (north_star): achieved HBM GB/s next to their integer work."""
    torch, ctx, dev, rank, world = env["torch"], env["ctx"], env["dev"], env["rank"], env["world"]
    n = env["n"]
    out = {}
    d_pts = env["d_pk"].clone()
    d_pts[63::64] = env["d_pk"][0]
    d_o128 = torch.empty(n, 128, dtype=torch.uint8, device=dev)
    d_s8 = torch.empty(n, dtype=torch.uint8, device=dev)
    s = env["timed"](lambda: ctx.dev_point_decompress(n, d_pts, d_o128, d_s8), reps=5)
    byt = n * (32 + 128 + 1)
    out["decompress"] = {"kernel": "k_point_decompress", "points": n * world, "ms": s * 1e3, "points_per_s": n * world / s, "bytes_per_point": 161,
                         "hbm_GBps": byt / s / 1e9, "hbm_frac": byt / s / 1e9 / env["hbm_peak"], "imad_frac": n / s * IMAD_EQ_PER_DECOMPRESS / env["imad_peak"],
                         "note": "32 B in, 128 B (X, Y, Z, T) + 1 status byte out per point; the stage is bound by the 255-squaring chain (integer pipe), not by HBM"}
    d_h = torch.empty(n, 32, dtype=torch.uint8, device=dev)
    d_r = env["d_sig"][:, :32].contiguous()
    s = env["timed"](lambda: ctx.dev_challenge(n, d_r, env["d_pk"], env["d_msg"], env["d_off"], d_h), reps=5)
    byt = n * (32 + 32 + 64 + 8 + 32)
    out["hash"] = {"kernel": "k_challenge", "items": n * world, "ms": s * 1e3, "hashes_per_s": n * world / s, "bytes_per_item": 168, "sha512_blocks_per_item": 2,
                   "hbm_GBps": byt / s / 1e9, "hbm_frac": byt / s / 1e9 / env["hbm_peak"], "sha512_GBps": n * 128 / s / 1e9,
                   "note": "SHA-512(R || A || M) for 64-byte messages (2 blocks) + reduction mod L; ALU-pipe bound"}
    # long messages: the hash rate that bounds signatures over large deals
    m, ml = 1 << 15, 8192
    d_lm = torch.randint(0, 256, (m * ml,), dtype=torch.uint8, device=dev)
    d_lo = (torch.arange(m + 1, dtype=torch.int64, device=dev) * ml)
    d_h2 = torch.empty(m, 32, dtype=torch.uint8, device=dev)
    d_r2, d_a2 = d_r[:m].contiguous(), env["d_pk"][:m].contiguous()
    s = env["timed"](lambda: ctx.dev_challenge(m, d_r2, d_a2, d_lm, d_lo, d_h2), reps=3)
    out["hash_long_messages"] = {"kernel": "k_challenge", "items": m, "message_bytes": ml, "ms": s * 1e3, "sha512_GBps": m * (ml + 64) / s / 1e9, "hbm_frac": m * (ml + 64) / s / 1e9 / env["hbm_peak"],
                                 "note": "one thread per message: each thread streams its own 8 KB"}
    return out


def run_all(env):
    configs = {}
    configs["cfg1"] = cfg1(env)
    configs["cfg3"] = dkg_config(env, "cfg3_vss_n256_t171", 256, 171, with_signatures=False)
    configs["cfg4"] = dkg_config(env, "cfg4_dkg_n1024_t683", 1024, 683, with_signatures=True)
    configs["cfg5"] = cfg5(env)
    stages = stage_rates(env)
    return configs, stages
