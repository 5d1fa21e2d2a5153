/*
 * oracle/ref10_port.c — CPU restatement of kyber-rs's edwards25519 hot path.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Nothing under kyber-rs_b200/ may link,
 * load or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * It follows the reference's ALGORITHMS (same limb schedule, same windowing, same
 * sequence of field operations per call) so that it is both a differential oracle and
 * an honest proxy for the pure-Rust path's cost (no Rust toolchain exists in this
 * image).  It is a restatement, not a copy: loops instead of the
 * reference's unrolled bodies, the BASE table computed at start-up instead of
 * transcribed, scalar arithmetic mod L done with 64-bit limbs instead of ref10's
 * 21-bit-limb sc_* bodies (result identical: fully reduced (ab+c) mod L).
 *
 * Parity pinning: tests/test_oracle_golden.py checks this library against the 1024-case
 * sign.input golden file, RFC 8032 §7.1 vectors, the reference's reject vectors,
 * WEAK_KEYS, scalar KATs, libsodium and oracle/ed25519_bigint.py.
 *
 * Citations (file:line) are relative to /root/reference/src.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t fe[10]; /* group/edwards25519/fe.rs:8 — radix 2^25.5, limbs 26/25/26/25/... bits */
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ field */

static void fe_0(fe h) { memset(h, 0, sizeof(fe)); }
static void fe_1(fe h) { fe_0(h); h[0] = 1; }
static void fe_copy(fe h, const fe f) { memcpy(h, f, sizeof(fe)); }
static void fe_add(fe h, const fe f, const fe g) { for (int i = 0; i < 10; i++) h[i] = f[i] + g[i]; }
static void fe_sub(fe h, const fe f, const fe g) { for (int i = 0; i < 10; i++) h[i] = f[i] - g[i]; }
static void fe_neg(fe h, const fe f) { for (int i = 0; i < 10; i++) h[i] = -f[i]; }

/* fe.rs:41 fe_c_move */
static void fe_cmov(fe f, const fe g, int32_t b)
{
    int32_t m = -b;
    for (int i = 0; i < 10; i++) f[i] ^= m & (f[i] ^ g[i]);
}

static uint64_t load3(const uint8_t *in) { return (uint64_t)in[0] | ((uint64_t)in[1] << 8) | ((uint64_t)in[2] << 16); }
static uint64_t load4(const uint8_t *in) { return load3(in) | ((uint64_t)in[3] << 24); }

/* The 12-step carry chain closing fe_mul / fe_square (fe.rs:299ff, 544ff). */
static inline void fe_carry_store(fe out, int64_t h[10])
{
    int64_t c;
#define CARRY26(i, j) c = (h[i] + ((int64_t)1 << 25)) >> 26; h[j] += c; h[i] -= c << 26
#define CARRY25(i, j) c = (h[i] + ((int64_t)1 << 24)) >> 25; h[j] += c; h[i] -= c << 25
    CARRY26(0, 1); CARRY26(4, 5);
    CARRY25(1, 2); CARRY25(5, 6);
    CARRY26(2, 3); CARRY26(6, 7);
    CARRY25(3, 4); CARRY25(7, 8);
    CARRY26(4, 5); CARRY26(8, 9);
    c = (h[9] + ((int64_t)1 << 24)) >> 25; h[0] += c * 19; h[9] -= c << 25;
    CARRY26(0, 1);
#undef CARRY26
#undef CARRY25
    for (int i = 0; i < 10; i++) out[i] = (int32_t)h[i];
}

/* fe.rs:67 fe_from_bytes — bit 255 ignored, no range check */
static void fe_frombytes(fe h, const uint8_t *s)
{
    int64_t t[10];
    t[0] = load4(s);
    t[1] = load3(s + 4) << 6;
    t[2] = load3(s + 7) << 5;
    t[3] = load3(s + 10) << 3;
    t[4] = load3(s + 13) << 2;
    t[5] = load4(s + 16);
    t[6] = load3(s + 20) << 7;
    t[7] = load3(s + 23) << 5;
    t[8] = load3(s + 26) << 4;
    t[9] = (load3(s + 29) & 8388607) << 2;
    int64_t c;
    c = (t[9] + ((int64_t)1 << 24)) >> 25; t[0] += c * 19; t[9] -= c << 25;
    c = (t[1] + ((int64_t)1 << 24)) >> 25; t[2] += c; t[1] -= c << 25;
    c = (t[3] + ((int64_t)1 << 24)) >> 25; t[4] += c; t[3] -= c << 25;
    c = (t[5] + ((int64_t)1 << 24)) >> 25; t[6] += c; t[5] -= c << 25;
    c = (t[7] + ((int64_t)1 << 24)) >> 25; t[8] += c; t[7] -= c << 25;
    c = (t[0] + ((int64_t)1 << 25)) >> 26; t[1] += c; t[0] -= c << 26;
    c = (t[2] + ((int64_t)1 << 25)) >> 26; t[3] += c; t[2] -= c << 26;
    c = (t[4] + ((int64_t)1 << 25)) >> 26; t[5] += c; t[4] -= c << 26;
    c = (t[6] + ((int64_t)1 << 25)) >> 26; t[7] += c; t[6] -= c << 26;
    c = (t[8] + ((int64_t)1 << 25)) >> 26; t[9] += c; t[8] -= c << 26;
    for (int i = 0; i < 10; i++) h[i] = (int32_t)t[i];
}

/* fe.rs:147 fe_to_bytes — fully reduced canonical encoding */
static void fe_tobytes(uint8_t *s, const fe f)
{
    int32_t h[10];
    memcpy(h, f, sizeof(h));
    int32_t q = (19 * h[9] + ((int32_t)1 << 24)) >> 25;
    for (int i = 0; i < 10; i++) q = (h[i] + q) >> ((i & 1) ? 25 : 26);
    h[0] += 19 * q;
    for (int i = 0; i < 9; i++) {
        int sh = (i & 1) ? 25 : 26;
        int32_t c = h[i] >> sh;
        h[i + 1] += c;
        h[i] -= c << sh;
    }
    h[9] -= (h[9] >> 25) << 25;
    /* pack 26/25-bit limbs */
    static const int off[10] = {0, 26, 51, 77, 102, 128, 153, 179, 204, 230};
    memset(s, 0, 32);
    for (int i = 0; i < 10; i++) {
        uint64_t v = (uint32_t)h[i];
        int byte = off[i] >> 3, bit = off[i] & 7;
        v <<= bit;
        for (int k = 0; k < 5 && byte + k < 32; k++) s[byte + k] |= (uint8_t)(v >> (8 * k));
    }
}

static int fe_isnegative(const fe f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }   /* fe.rs:240 */
static int fe_isnonzero(const fe f)                                                            /* fe.rs:246 */
{
    uint8_t s[32], x = 0;
    fe_tobytes(s, f);
    for (int i = 0; i < 32; i++) x |= s[i];
    return x != 0;
}

/* fe.rs:299 fe_mul (100 products) and fe.rs:544 fe_square (55 products): straight-line product
 * sums emitted by oracle/gen_fe.py, closed by the 12-step carry chain above. */
#include "fe_gen.inc"
static void fe_mul(fe out, const fe f, const fe g) { int64_t h[10]; fe_mul_raw(h, f, g); fe_carry_store(out, h); }
static void fe_sq(fe out, const fe f) { int64_t h[10]; fe_sq_raw(h, f); fe_carry_store(out, h); }
/* fe.rs:700 fe_square2 */
static void fe_sq2(fe out, const fe f)
{
    int64_t h[10];
    fe_sq_raw(h, f);
    for (int i = 0; i < 10; i++) h[i] += h[i];
    fe_carry_store(out, h);
}

static void fe_sqn(fe out, const fe f, int n) { fe_sq(out, f); for (int i = 1; i < n; i++) fe_sq(out, out); }

/* shared prefix of fe_invert / fe_pow22523: returns z^(2^250-1) in t, z^11 in z11 */
static void fe_pow_2_250_1(fe t, fe z11, const fe z)
{
    fe t0, t1, t2, t3;
    fe_sq(t0, z);              /* 2 */
    fe_sqn(t1, t0, 2);         /* 8 */
    fe_mul(t1, z, t1);         /* 9 */
    fe_mul(t0, t0, t1);        /* 11 */
    fe_copy(z11, t0);
    fe_sq(t2, t0);             /* 22 */
    fe_mul(t1, t1, t2);        /* 2^5-1 */
    fe_sqn(t2, t1, 5); fe_mul(t1, t2, t1);     /* 2^10-1 */
    fe_sqn(t2, t1, 10); fe_mul(t2, t2, t1);    /* 2^20-1 */
    fe_sqn(t3, t2, 20); fe_mul(t2, t3, t2);    /* 2^40-1 */
    fe_sqn(t2, t2, 10); fe_mul(t1, t2, t1);    /* 2^50-1 */
    fe_sqn(t2, t1, 50); fe_mul(t2, t2, t1);    /* 2^100-1 */
    fe_sqn(t3, t2, 100); fe_mul(t2, t3, t2);   /* 2^200-1 */
    fe_sqn(t2, t2, 50); fe_mul(t, t2, t1);     /* 2^250-1 */
}

/* fe.rs:857 fe_invert: z^(p-2) = z^(2^255-21) */
static void fe_invert(fe out, const fe z)
{
    fe t, z11;
    fe_pow_2_250_1(t, z11, z);
    fe_sqn(t, t, 5);
    fe_mul(out, t, z11);
}

/* fe.rs:946 fe_pow22523: z^((p-5)/8) = z^(2^252-3) */
static void fe_pow22523(fe out, const fe z)
{
    fe t, z11;
    fe_pow_2_250_1(t, z11, z);
    fe_sqn(t, t, 2);
    fe_mul(out, t, z);
}

/* ------------------------------------------------------------------ curve constants (facts about the curve) */
static fe FE_D, FE_D2, FE_SQRTM1;
static void hex32(uint8_t out[32], const char *h)
{
    for (int i = 0; i < 32; i++) {
        unsigned v = 0;
        for (int k = 0; k < 2; k++) {
            char c = h[2 * i + k];
            v = v * 16 + (unsigned)(c <= '9' ? c - '0' : c - 'a' + 10);
        }
        out[i] = (uint8_t)v;
    }
}

/* ------------------------------------------------------------------ group elements (ge.rs) */
typedef struct { fe X, Y, Z; } ge_p2;                 /* ProjectiveGroupElement */
typedef struct { fe X, Y, Z, T; } ge_p3;              /* ExtendedGroupElement   */
typedef struct { fe X, Y, Z, T; } ge_p1p1;            /* CompletedGroupElement  */
typedef struct { fe yplusx, yminusx, xy2d; } ge_precomp; /* PreComputedGroupElement */
typedef struct { fe YplusX, YminusX, Z, T2d; } ge_cached; /* CachedGroupElement */

static void ge_p3_0(ge_p3 *h) { fe_0(h->X); fe_1(h->Y); fe_1(h->Z); fe_0(h->T); }
static void ge_p1p1_to_p2(ge_p2 *r, const ge_p1p1 *p) { fe_mul(r->X, p->X, p->T); fe_mul(r->Y, p->Y, p->Z); fe_mul(r->Z, p->Z, p->T); }           /* ge.rs:211 */
static void ge_p1p1_to_p3(ge_p3 *r, const ge_p1p1 *p) { fe_mul(r->X, p->X, p->T); fe_mul(r->Y, p->Y, p->Z); fe_mul(r->Z, p->Z, p->T); fe_mul(r->T, p->X, p->Y); } /* ge.rs:292 */
static void ge_p3_to_cached(ge_cached *r, const ge_p3 *p) { fe_add(r->YplusX, p->Y, p->X); fe_sub(r->YminusX, p->Y, p->X); fe_copy(r->Z, p->Z); fe_mul(r->T2d, p->T, FE_D2); } /* ge.rs:99 */

/* ge.rs:35 ProjectiveGroupElement::double */
static void ge_p2_dbl(ge_p1p1 *r, const ge_p2 *p)
{
    fe t0;
    fe_sq(r->X, p->X);
    fe_sq(r->Z, p->Y);
    fe_sq2(r->T, p->Z);
    fe_add(r->Y, p->X, p->Y);
    fe_sq(t0, r->Y);
    fe_add(r->Y, r->Z, r->X);
    fe_sub(r->Z, r->Z, r->X);
    fe_sub(r->X, t0, r->Y);
    fe_sub(r->T, r->T, r->Z);
}
static void ge_p3_dbl(ge_p1p1 *r, const ge_p3 *p) { ge_p2 q; fe_copy(q.X, p->X); fe_copy(q.Y, p->Y); fe_copy(q.Z, p->Z); ge_p2_dbl(r, &q); }

/* ge.rs:274 mixed_add */
static void ge_madd(ge_p1p1 *r, const ge_p3 *p, const ge_precomp *q)
{
    fe t0;
    fe_add(r->X, p->Y, p->X);
    fe_sub(r->Y, p->Y, p->X);
    fe_mul(r->Z, r->X, q->yplusx);
    fe_mul(r->Y, r->Y, q->yminusx);
    fe_mul(r->T, q->xy2d, p->T);
    fe_add(t0, p->Z, p->Z);
    fe_sub(r->X, r->Z, r->Y);
    fe_add(r->Y, r->Z, r->Y);
    fe_add(r->Z, t0, r->T);
    fe_sub(r->T, t0, r->T);
}
/* ge.rs:217 add */
static void ge_add(ge_p1p1 *r, const ge_p3 *p, const ge_cached *q)
{
    fe t0;
    fe_add(r->X, p->Y, p->X);
    fe_sub(r->Y, p->Y, p->X);
    fe_mul(r->Z, r->X, q->YplusX);
    fe_mul(r->Y, r->Y, q->YminusX);
    fe_mul(r->T, q->T2d, p->T);
    fe_mul(r->X, p->Z, q->Z);
    fe_add(t0, r->X, r->X);
    fe_sub(r->X, r->Z, r->Y);
    fe_add(r->Y, r->Z, r->Y);
    fe_add(r->Z, t0, r->T);
    fe_sub(r->T, t0, r->T);
}
/* ge.rs:236 sub */
static void ge_sub(ge_p1p1 *r, const ge_p3 *p, const ge_cached *q)
{
    fe t0;
    fe_add(r->X, p->Y, p->X);
    fe_sub(r->Y, p->Y, p->X);
    fe_mul(r->Z, r->X, q->YminusX);
    fe_mul(r->Y, r->Y, q->YplusX);
    fe_mul(r->T, q->T2d, p->T);
    fe_mul(r->X, p->Z, q->Z);
    fe_add(t0, r->X, r->X);
    fe_sub(r->X, r->Z, r->Y);
    fe_add(r->Y, r->Z, r->Y);
    fe_sub(r->Z, t0, r->T);
    fe_add(r->T, t0, r->T);
}

/* ge.rs:112 write_bytes */
static void ge_p3_tobytes(uint8_t *s, const ge_p3 *h)
{
    fe recip, x, y;
    fe_invert(recip, h->Z);
    fe_mul(x, h->X, recip);
    fe_mul(y, h->Y, recip);
    fe_tobytes(s, y);
    s[31] ^= (uint8_t)(fe_isnegative(x) << 7);
}

/* ge.rs:124 set_bytes — returns 1 on success */
static int ge_frombytes(ge_p3 *h, const uint8_t *s)
{
    fe u, v, v3, vxx, check;
    fe_frombytes(h->Y, s);
    fe_1(h->Z);
    fe_sq(u, h->Y);
    fe_mul(v, u, FE_D);
    fe_sub(u, u, h->Z);
    fe_add(v, v, h->Z);
    fe_sq(v3, v);
    fe_mul(v3, v3, v);
    fe_sq(h->X, v3);
    fe_mul(h->X, h->X, v);
    fe_mul(h->X, h->X, u);
    fe_pow22523(h->X, h->X);
    fe_mul(h->X, h->X, v3);
    fe_mul(h->X, h->X, u);
    fe_sq(vxx, h->X);
    fe_mul(vxx, vxx, v);
    fe_sub(check, vxx, u);
    if (fe_isnonzero(check)) {
        fe_add(check, vxx, u);
        if (fe_isnonzero(check)) return 0;
        fe_mul(h->X, h->X, FE_SQRTM1);
    }
    if (fe_isnegative(h->X) != (s[31] >> 7)) fe_neg(h->X, h->X);
    fe_mul(h->T, h->X, h->Y);
    return 1;
}

/* Point::add / Point::sub (point.rs:179,190) */
static void pt_add(ge_p3 *r, const ge_p3 *a, const ge_p3 *b) { ge_cached c; ge_p1p1 t; ge_p3_to_cached(&c, b); ge_add(&t, a, &c); ge_p1p1_to_p3(r, &t); }
static void pt_sub(ge_p3 *r, const ge_p3 *a, const ge_p3 *b) { ge_cached c; ge_p1p1 t; ge_p3_to_cached(&c, b); ge_sub(&t, a, &c); ge_p1p1_to_p3(r, &t); }

/* BASE[32][8] (constants.rs:89): BASE[pos][j] = (j+1) * 256^pos * B in (y+x, y-x, 2dxy) form.
 * Computed once instead of transcribed. */
static ge_precomp BASE_TABLE[32][8];
static ge_p3 BASE_P3;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void to_precomp(ge_precomp *r, const ge_p3 *p)
{
    fe recip, x, y, xy;
    fe_invert(recip, p->Z);
    fe_mul(x, p->X, recip);
    fe_mul(y, p->Y, recip);
    fe_add(r->yplusx, y, x);
    fe_sub(r->yminusx, y, x);
    fe_mul(xy, x, y);
    fe_mul(r->xy2d, xy, FE_D2);
    /* normalise limbs through a byte round trip so the table is in reduced form */
    uint8_t b[32];
    fe_tobytes(b, r->yplusx); fe_frombytes(r->yplusx, b);
    fe_tobytes(b, r->yminusx); fe_frombytes(r->yminusx, b);
    fe_tobytes(b, r->xy2d); fe_frombytes(r->xy2d, b);
}

static void oracle_init_once(void)
{
    uint8_t b[32];
    hex32(b, "a3785913ca4deb75abd841414d0a700098e879777940c78c73fe6f2bee6c0352"); fe_frombytes(FE_D, b);      /* constants.rs:60 */
    hex32(b, "59f1b226949bd6eb56b183829a14e00030d1f3eef2808e19e7fcdf56dcd90624"); fe_frombytes(FE_D2, b);     /* constants.rs:65 */
    hex32(b, "b0a00e4a271beec478e42fad0618432fa7d7fb3d99004d2b0bdfc14f8024832b"); fe_frombytes(FE_SQRTM1, b); /* constants.rs:56 */
    hex32(b, "5866666666666666666666666666666666666666666666666666666666666666");                               /* y = 4/5, x positive */
    ge_frombytes(&BASE_P3, b);
    ge_p3 pos = BASE_P3;
    for (int i = 0; i < 32; i++) {
        ge_p3 m = pos;
        for (int j = 0; j < 8; j++) {
            to_precomp(&BASE_TABLE[i][j], &m);
            pt_add(&m, &m, &pos);
        }
        for (int k = 0; k < 8; k++) { ge_p1p1 t; ge_p3_dbl(&t, &pos); ge_p1p1_to_p3(&pos, &t); }
    }
}
void oracle_init(void) { pthread_once(&g_once, oracle_init_once); }

static int32_t ct_equal(int32_t b, int32_t c) { uint32_t x = (uint32_t)(b ^ c); x -= 1; return (int32_t)(x >> 31); }
static int32_t ct_negative(int32_t b) { return (b >> 31) & 1; }

/* ge.rs:423 select_pre_computed */
static void select_precomp(ge_precomp *t, int pos, int32_t b)
{
    ge_precomp minus;
    int32_t bneg = ct_negative(b);
    int32_t babs = b - (((-bneg) & b) << 1);
    fe_1(t->yplusx); fe_1(t->yminusx); fe_0(t->xy2d);
    for (int i = 0; i < 8; i++) {
        int32_t e = ct_equal(babs, i + 1);
        fe_cmov(t->yplusx, BASE_TABLE[pos][i].yplusx, e);
        fe_cmov(t->yminusx, BASE_TABLE[pos][i].yminusx, e);
        fe_cmov(t->xy2d, BASE_TABLE[pos][i].xy2d, e);
    }
    fe_copy(minus.yplusx, t->yminusx);
    fe_copy(minus.yminusx, t->yplusx);
    fe_neg(minus.xy2d, t->xy2d);
    fe_cmov(t->yplusx, minus.yplusx, bneg);
    fe_cmov(t->yminusx, minus.yminusx, bneg);
    fe_cmov(t->xy2d, minus.xy2d, bneg);
}

/* signed radix-16 recoding, ge.rs:443-458 / 521-535 */
static void recode16(int8_t e[64], const uint8_t a[32])
{
    for (int i = 0; i < 32; i++) { e[2 * i] = a[i] & 15; e[2 * i + 1] = (a[i] >> 4) & 15; }
    int8_t carry = 0;
    for (int i = 0; i < 63; i++) {
        e[i] += carry;
        carry = (int8_t)((e[i] + 8) >> 4);
        e[i] -= (int8_t)(carry << 4);
    }
    e[63] += carry;
}

/* ge.rs:442 ge_scalar_mult_base */
static void ge_scalarmult_base(ge_p3 *h, const uint8_t a[32])
{
    int8_t e[64];
    ge_p1p1 r; ge_p2 s; ge_precomp t;
    recode16(e, a);
    ge_p3_0(h);
    for (int i = 1; i < 64; i += 2) {
        select_precomp(&t, i / 2, e[i]);
        ge_madd(&r, h, &t); ge_p1p1_to_p3(h, &r);
    }
    ge_p3_dbl(&r, h); ge_p1p1_to_p2(&s, &r);
    ge_p2_dbl(&r, &s); ge_p1p1_to_p2(&s, &r);
    ge_p2_dbl(&r, &s); ge_p1p1_to_p2(&s, &r);
    ge_p2_dbl(&r, &s); ge_p1p1_to_p3(h, &r);
    for (int i = 0; i < 64; i += 2) {
        select_precomp(&t, i / 2, e[i]);
        ge_madd(&r, h, &t); ge_p1p1_to_p3(h, &r);
    }
}

/* ge.rs:488 select_cached */
static void select_cached(ge_cached *c, const ge_cached ai[8], int32_t b)
{
    ge_cached minus;
    int32_t bneg = ct_negative(b);
    int32_t babs = b - (((-bneg) & b) << 1);
    fe_1(c->YplusX); fe_1(c->YminusX); fe_1(c->Z); fe_0(c->T2d);
    for (int i = 0; i < 8; i++) {
        int32_t e = ct_equal(babs, i + 1);
        fe_cmov(c->YplusX, ai[i].YplusX, e);
        fe_cmov(c->YminusX, ai[i].YminusX, e);
        fe_cmov(c->Z, ai[i].Z, e);
        fe_cmov(c->T2d, ai[i].T2d, e);
    }
    fe_copy(minus.YplusX, c->YminusX);
    fe_copy(minus.YminusX, c->YplusX);
    fe_copy(minus.Z, c->Z);
    fe_neg(minus.T2d, c->T2d);
    fe_cmov(c->YplusX, minus.YplusX, bneg);
    fe_cmov(c->YminusX, minus.YminusX, bneg);
    fe_cmov(c->Z, minus.Z, bneg);
    fe_cmov(c->T2d, minus.T2d, bneg);
}

/* ge.rs:508 ge_scalar_mult — constant-time fixed window 4 */
static void ge_scalarmult(ge_p3 *h, const uint8_t a[32], const ge_p3 *A)
{
    int8_t e[64];
    ge_p1p1 t; ge_p3 u; ge_p2 r; ge_cached c, ai[8];
    recode16(e, a);
    ge_p3_to_cached(&ai[0], A);
    for (int i = 0; i < 7; i++) { ge_add(&t, A, &ai[i]); ge_p1p1_to_p3(&u, &t); ge_p3_to_cached(&ai[i + 1], &u); }
    ge_p3_0(&u);
    select_cached(&c, ai, e[63]);
    ge_add(&t, &u, &c);
    for (int i = 62; i >= 0; i--) {
        ge_p1p1_to_p2(&r, &t); ge_p2_dbl(&t, &r);
        ge_p1p1_to_p2(&r, &t); ge_p2_dbl(&t, &r);
        ge_p1p1_to_p2(&r, &t); ge_p2_dbl(&t, &r);
        ge_p1p1_to_p2(&r, &t); ge_p2_dbl(&t, &r);
        ge_p1p1_to_p3(&u, &t);
        select_cached(&c, ai, e[i]);
        ge_add(&t, &u, &c);
    }
    ge_p1p1_to_p3(h, &t);
}

/* ------------------------------------------------------------------ SHA-512 (FIPS 180-4; the reference uses the sha2 crate ^0.10.6, Cargo.toml:27) */
static const uint64_t K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL,
    0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL,
    0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL, 0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, 0x06ca6351e003826fULL, 0x142929670a0e6e70ULL,
    0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL,
    0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL, 0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL,
    0xca273eceea26619cULL, 0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};

typedef struct { uint64_t h[8]; uint8_t buf[128]; uint64_t len; } sha512_ctx;
static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static void sha512_block(uint64_t h[8], const uint8_t *p)
{
    uint64_t w[80];
    for (int i = 0; i < 16; i++) { uint64_t v = 0; for (int k = 0; k < 8; k++) v = (v << 8) | p[8 * i + k]; w[i] = v; }
    for (int i = 16; i < 80; i++) {
        uint64_t s0 = rotr64(w[i - 15], 1) ^ rotr64(w[i - 15], 8) ^ (w[i - 15] >> 7);
        uint64_t s1 = rotr64(w[i - 2], 19) ^ rotr64(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 80; i++) {
        uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
        uint64_t ch = (e & f) ^ (~e & g);
        uint64_t t1 = hh + S1 + ch + K512[i] + w[i];
        uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
        uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint64_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha512_init(sha512_ctx *c)
{
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    memcpy(c->h, iv, sizeof(iv));
    c->len = 0;
}
static void sha512_update(sha512_ctx *c, const uint8_t *p, size_t n)
{
    size_t fill = (size_t)(c->len & 127);
    c->len += n;
    if (fill) {
        size_t take = 128 - fill; if (take > n) take = n;
        memcpy(c->buf + fill, p, take); p += take; n -= take; fill += take;
        if (fill < 128) return;
        sha512_block(c->h, c->buf);
    }
    while (n >= 128) { sha512_block(c->h, p); p += 128; n -= 128; }
    if (n) memcpy(c->buf, p, n);
}
static void sha512_final(sha512_ctx *c, uint8_t out[64])
{
    size_t fill = (size_t)(c->len & 127);
    uint64_t bits = c->len * 8;
    c->buf[fill++] = 0x80;
    if (fill > 112) { memset(c->buf + fill, 0, 128 - fill); sha512_block(c->h, c->buf); fill = 0; }
    memset(c->buf + fill, 0, 120 - fill);
    for (int k = 0; k < 8; k++) c->buf[120 + k] = (uint8_t)(bits >> (56 - 8 * k));
    sha512_block(c->h, c->buf);
    for (int i = 0; i < 8; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(c->h[i] >> (56 - 8 * k));
}
void oracle_sha512(const uint8_t *msg, size_t n, uint8_t out[64]) { sha512_ctx c; sha512_init(&c); sha512_update(&c, msg, n); sha512_final(&c, out); }

/* ------------------------------------------------------------------ scalars mod L (scalar.rs) */
static const uint64_t LQ[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};
static const uint64_t LC[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL}; /* L - 2^252 */

/* r[0..n) = a - b, returns borrow */
static uint64_t sub_n(uint64_t *r, const uint64_t *a, const uint64_t *b, int n)
{
    uint64_t br = 0;
    for (int i = 0; i < n; i++) { u128 d = (u128)a[i] - b[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    return br;
}
static uint64_t add_n(uint64_t *r, const uint64_t *a, const uint64_t *b, int n)
{
    uint64_t c = 0;
    for (int i = 0; i < n; i++) { u128 s = (u128)a[i] + b[i] + c; r[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    return c;
}
/* split x (n limbs) at bit 252: lo[4] (252 bits), hi[n-3] = x >> 252 */
static void split252(uint64_t lo[4], uint64_t *hi, const uint64_t *x, int n)
{
    for (int i = 0; i < 4; i++) lo[i] = x[i];
    lo[3] &= 0x0fffffffffffffffULL;
    for (int i = 3; i < n; i++) hi[i - 3] = (x[i] >> 60) | ((i + 1 < n) ? (x[i + 1] << 4) : 0);
}
/* r[na+2] = a[na] * LC[2] */
static void mul_c(uint64_t *r, const uint64_t *a, int na)
{
    memset(r, 0, (size_t)(na + 2) * 8);
    for (int i = 0; i < na; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 2; j++) { u128 t = (u128)a[i] * LC[j] + r[i + j] + carry; r[i + j] = (uint64_t)t; carry = (uint64_t)(t >> 64); }
        r[i + 2] += carry;
    }
}
/* x: 8 limbs (512 bits) -> x mod L, 4 limbs.  2^252 = -c (mod L): fold three times. */
static void mod_l_512(uint64_t r[4], const uint64_t x[8])
{
    uint64_t lo1[4], hi1[5], y[7], lo2[4], hi2[4], z[6], lo3[4], hi3[3], w[5], t[4];
    split252(lo1, hi1, x, 8);            /* hi1 < 2^260 */
    mul_c(y, hi1, 5);                    /* y = hi1*c < 2^385 */
    split252(lo2, hi2, y, 7);            /* hi2 < 2^133 */
    hi2[3] = 0;
    mul_c(z, hi2, 3);                    /* z = hi2*c < 2^258 */
    split252(lo3, hi3, z, 5);            /* hi3 < 2^6 */
    mul_c(w, hi3, 1);                    /* w = hi3*c < 2^131 */
    w[3] = 0;
    /* x = lo1 - (lo2 - (lo3 - w))  (mod L), every term in [0, L) after fix-up */
    if (sub_n(t, lo3, w, 4)) add_n(t, t, LQ, 4);
    if (sub_n(t, lo2, t, 4)) add_n(t, t, LQ, 4);
    if (sub_n(t, lo1, t, 4)) add_n(t, t, LQ, 4);
    memcpy(r, t, 32);
}
static void load256(uint64_t r[4], const uint8_t b[32]) { for (int i = 0; i < 4; i++) { uint64_t v = 0; for (int k = 7; k >= 0; k--) v = (v << 8) | b[8 * i + k]; r[i] = v; } }
static void store256(uint8_t b[32], const uint64_t r[4]) { for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) b[8 * i + k] = (uint8_t)(r[i] >> (8 * k)); }

/* Scalar::set_bytes for a 64-byte digest (scalar.rs:175, integer_field/integer.rs:386) */
void oracle_sc_reduce64(uint8_t out[32], const uint8_t in[64])
{
    uint64_t x[8], r[4];
    load256(x, in); load256(x + 4, in + 32);
    mod_l_512(r, x);
    store256(out, r);
}
/* scalar.rs:279 sc_mul_add: (ab+c) mod L */
void oracle_sc_muladd(uint8_t s[32], const uint8_t a[32], const uint8_t b[32], const uint8_t c[32])
{
    uint64_t A[4], B[4], C[4], x[8] = {0}, r[4];
    load256(A, a); load256(B, b); load256(C, c);
    for (int i = 0; i < 4; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 4; j++) { u128 t = (u128)A[i] * B[j] + x[i + j] + carry; x[i + j] = (uint64_t)t; carry = (uint64_t)(t >> 64); }
        x[i + 4] = carry;
    }
    uint64_t cy = 0;
    for (int i = 0; i < 8; i++) { u128 t = (u128)x[i] + (i < 4 ? C[i] : 0) + cy; x[i] = (uint64_t)t; cy = (uint64_t)(t >> 64); }
    mod_l_512(r, x);
    store256(s, r);
}
/* Scalar::is_canonical (scalar.rs:54) */
int oracle_scalar_is_canonical(const uint8_t sb[32])
{
    if ((sb[31] & 0xf0) == 0) return 1;
    uint64_t v[4], t[4];
    load256(v, sb);
    return (int)sub_n(t, v, LQ, 4);
}

/* ------------------------------------------------------------------ Point-level API (point.rs) */
/* Point::is_canonical (point.rs:322) with its u16-wrapping quirk */
int oracle_point_is_canonical(const uint8_t b[32])
{
    uint8_t c = (uint8_t)((b[31] & 0x7f) ^ 0x7f);
    for (int i = 30; i >= 1; i--) c |= b[i] ^ 0xff;
    c = (uint8_t)((((uint16_t)c - 1) & 0xffff) >> 8);
    uint16_t one_minus = (uint16_t)(1 - (uint16_t)b[0]);
    uint8_t d = (uint8_t)(((uint16_t)(0xED - one_minus)) >> 8);
    return 1 - (c & d & 1);
}
static const uint8_t WEAK_Y[5][32] = {
    {0},
    {1},
    {0x26, 0xe8, 0x95, 0x8f, 0xc2, 0xb2, 0x27, 0xb0, 0x45, 0xc3, 0xf4, 0x89, 0xf2, 0xef, 0x98, 0xf0, 0xd5, 0xdf, 0xac, 0x05, 0xd3, 0xc6, 0x33, 0x39, 0xb1, 0x38, 0x02, 0x88, 0x6d, 0x53, 0xfc, 0x05},
    {0xc7, 0x17, 0x6a, 0x70, 0x3d, 0x4d, 0xd8, 0x4f, 0xba, 0x3c, 0x0b, 0x76, 0x0d, 0x10, 0x67, 0x0f, 0x2a, 0x20, 0x53, 0xfa, 0x2c, 0x39, 0xcc, 0xc6, 0x4e, 0xc7, 0xfd, 0x77, 0x92, 0xac, 0x03, 0x7a},
    {0xec, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x7f}};
/* Point::has_small_order (point.rs:286): re-encodes, then compares with WEAK_KEYS (constants.rs:3744) */
static int has_small_order(const ge_p3 *p)
{
    uint8_t s[32], c[5] = {0};
    ge_p3_tobytes(s, p);
    for (int j = 0; j < 31; j++) for (int i = 0; i < 5; i++) c[i] |= s[j] ^ WEAK_Y[i][j];
    for (int i = 0; i < 5; i++) c[i] |= (s[31] & 0x7f) ^ WEAK_Y[i][31];
    uint16_t k = 0;
    for (int i = 0; i < 5; i++) k |= (uint16_t)((uint16_t)c[i] - 1);
    return (k >> 8) & 1;
}
static int pt_eq(const ge_p3 *a, const ge_p3 *b) { uint8_t x[32], y[32]; ge_p3_tobytes(x, a); ge_p3_tobytes(y, b); return memcmp(x, y, 32) == 0; } /* point.rs:227 */

int oracle_point_decode_ok(const uint8_t s[32]) { ge_p3 p; oracle_init(); return ge_frombytes(&p, s); }
/* decode then re-encode; returns 0 if decode fails */
int oracle_point_recode(uint8_t out[32], const uint8_t s[32]) { ge_p3 p; oracle_init(); if (!ge_frombytes(&p, s)) return 0; ge_p3_tobytes(out, &p); return 1; }
int oracle_point_has_small_order(const uint8_t s[32]) { ge_p3 p; oracle_init(); if (!ge_frombytes(&p, s)) return -1; return has_small_order(&p); }
/* Point::mul(s, None) (point.rs:207) -> compressed */
void oracle_mul_base(uint8_t out[32], const uint8_t a[32]) { ge_p3 h; oracle_init(); ge_scalarmult_base(&h, a); ge_p3_tobytes(out, &h); }
/* Point::mul(s, Some(p)) -> compressed; returns 0 if p fails to decode */
int oracle_mul(uint8_t out[32], const uint8_t a[32], const uint8_t p[32])
{
    ge_p3 A, h; oracle_init();
    if (!ge_frombytes(&A, p)) return 0;
    ge_scalarmult(&h, a, &A); ge_p3_tobytes(out, &h); return 1;
}
int oracle_point_add(uint8_t out[32], const uint8_t p[32], const uint8_t q[32], int subtract)
{
    ge_p3 A, B, R; oracle_init();
    if (!ge_frombytes(&A, p) || !ge_frombytes(&B, q)) return 0;
    if (subtract) pt_sub(&R, &A, &B); else pt_add(&R, &A, &B);
    ge_p3_tobytes(out, &R); return 1;
}

/* raw ExtendedGroupElement limbs (X,Y,Z,T: 4 x 10 i32 — the reference's serde wire format, ge.rs:75-83) of a*B */
void oracle_mul_base_limbs(int32_t out[40], const uint8_t a[32])
{
    ge_p3 h; oracle_init(); ge_scalarmult_base(&h, a);
    memcpy(out, h.X, 40); memcpy(out + 10, h.Y, 40); memcpy(out + 20, h.Z, 40); memcpy(out + 30, h.T, 40);
}
/* the limbs ExtendedGroupElement::set_bytes (ge.rs:124-179) leaves for an encoding (Z = 1); returns 0 if it does not decode */
int oracle_point_limbs(int32_t out[40], const uint8_t s[32])
{
    ge_p3 h; oracle_init();
    if (!ge_frombytes(&h, s)) return 0;
    memcpy(out, h.X, 40); memcpy(out + 10, h.Y, 40); memcpy(out + 20, h.Z, 40); memcpy(out + 30, h.T, 40);
    return 1;
}
/* write_bytes (ge.rs:112-122) of an element given by raw limbs */
void oracle_limbs_tobytes(uint8_t out[32], const int32_t in[40])
{
    ge_p3 h; oracle_init();
    memcpy(h.X, in, 40); memcpy(h.Y, in + 10, 40); memcpy(h.Z, in + 20, 40); memcpy(h.T, in + 30, 40);
    ge_p3_tobytes(out, &h);
}

/* ------------------------------------------------------------------ signatures */
enum { ST_OK = 0, ST_SIG_LENGTH = 1, ST_SIG_NOT_CANONICAL = 2, ST_R_NOT_CANONICAL = 3, ST_R_SMALL_ORDER = 4, ST_PK_NOT_CANONICAL = 5, ST_PK_SMALL_ORDER = 6, ST_MARSHALLING = 7, ST_INVALID_SIGNATURE = 8 };

static void challenge(uint8_t h[32], const uint8_t r[32], const uint8_t a[32], const uint8_t *msg, size_t mlen)
{
    sha512_ctx c; uint8_t d[64];
    sha512_init(&c); sha512_update(&c, r, 32); sha512_update(&c, a, 32); sha512_update(&c, msg, mlen); sha512_final(&c, d);
    oracle_sc_reduce64(h, d);
}
static int check_equation(const ge_p3 *R, const ge_p3 *A, const uint8_t s[32], const uint8_t h[32])
{
    ge_p3 sB, hA, rhs;
    ge_scalarmult_base(&sB, s);
    ge_scalarmult(&hA, h, A);
    pt_add(&rhs, R, &hA);
    return pt_eq(&rhs, &sB) ? ST_OK : ST_INVALID_SIGNATURE;
}
/* sign/eddsa/eddsa_sig.rs:159-212 */
int oracle_eddsa_verify(const uint8_t pk[32], const uint8_t *msg, size_t mlen, const uint8_t *sig, size_t siglen)
{
    ge_p3 R, A; uint8_t h[32];
    oracle_init();
    if (siglen != 64) return ST_SIG_LENGTH;
    if (!oracle_scalar_is_canonical(sig + 32)) return ST_SIG_NOT_CANONICAL;
    if (!oracle_point_is_canonical(sig)) return ST_R_NOT_CANONICAL;
    if (!ge_frombytes(&R, sig)) return ST_MARSHALLING;
    if (has_small_order(&R)) return ST_R_SMALL_ORDER;
    if (!oracle_point_is_canonical(pk)) return ST_PK_NOT_CANONICAL;
    if (!ge_frombytes(&A, pk)) return ST_MARSHALLING;
    if (has_small_order(&A)) return ST_PK_SMALL_ORDER;
    challenge(h, sig, pk, msg, mlen);
    return check_equation(&R, &A, sig + 32, h);
}
/* sign/schnorr/schnorr_sig.rs:53-110 (+ hash :128-141 on re-encoded R, A) */
int oracle_schnorr_verify(const uint8_t pk[32], const uint8_t *msg, size_t mlen, const uint8_t *sig, size_t siglen)
{
    ge_p3 R, A; uint8_t h[32], rb[32], ab[32];
    oracle_init();
    if (siglen != 64) return ST_SIG_LENGTH;
    if (!ge_frombytes(&R, sig)) return ST_MARSHALLING;
    if (!oracle_point_is_canonical(sig)) return ST_R_NOT_CANONICAL;
    if (has_small_order(&R)) return ST_R_SMALL_ORDER;
    if (!oracle_scalar_is_canonical(sig + 32)) return ST_SIG_NOT_CANONICAL;
    if (!ge_frombytes(&A, pk)) return ST_MARSHALLING;
    if (!oracle_point_is_canonical(pk)) return ST_PK_NOT_CANONICAL;
    if (has_small_order(&A)) return ST_PK_SMALL_ORDER;
    ge_p3_tobytes(rb, &R); ge_p3_tobytes(ab, &A);
    challenge(h, rb, ab, msg, mlen);
    return check_equation(&R, &A, sig + 32, h);
}

/* ------------------------------------------------------------------ polynomials (share/poly.rs) */
static void sc_from_u64(uint8_t out[32], uint64_t v) { memset(out, 0, 32); for (int k = 0; k < 8; k++) out[k] = (uint8_t)(v >> (8 * k)); }
/* PubPoly::eval (poly.rs:457-469): t full constant-time scalar mults by xi = 1+i */
static void pubpoly_eval_p3(ge_p3 *v, const ge_p3 *commits, int t, uint32_t idx)
{
    uint8_t xi[32];
    sc_from_u64(xi, 1 + (uint64_t)idx);
    ge_p3_0(v);
    for (int j = t - 1; j >= 0; j--) {
        ge_p3 m;
        ge_scalarmult(&m, xi, v);
        pt_add(v, &m, &commits[j]);
    }
}
/* returns 0 if a commitment fails to decode */
int oracle_pubpoly_eval(uint8_t out[32], const uint8_t *commits32, int t, uint32_t idx)
{
    oracle_init();
    ge_p3 *c = (ge_p3 *)malloc(sizeof(ge_p3) * (size_t)t), v;
    for (int j = 0; j < t; j++) if (!ge_frombytes(&c[j], commits32 + 32 * j)) { free(c); return 0; }
    pubpoly_eval_p3(&v, c, t, idx);
    ge_p3_tobytes(out, &v);
    free(c);
    return 1;
}
/* vss/pedersen/vss.rs:899-912: share*B == eval(idx) on canonical bytes.  1 = verifies, 0 = does not, -1 = bad commitment */
int oracle_vss_verify_deal(const uint8_t *commits32, int t, uint32_t idx, const uint8_t share[32])
{
    oracle_init();
    ge_p3 *c = (ge_p3 *)malloc(sizeof(ge_p3) * (size_t)t), v, fig;
    for (int j = 0; j < t; j++) if (!ge_frombytes(&c[j], commits32 + 32 * j)) { free(c); return -1; }
    ge_scalarmult_base(&fig, share);
    pubpoly_eval_p3(&v, c, t, idx);
    free(c);
    return pt_eq(&fig, &v);
}

/* ------------------------------------------------------------------ scalars used by the interpolation code */
static const uint8_t SC_LM1[32] = {0xec, 0xd3, 0xf5, 0x5c, 0x1a, 0x63, 0x12, 0x58, 0xd6, 0x9c, 0xf7, 0xa2, 0xde, 0xf9, 0xde, 0x14, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x10};   /* L - 1 */
static const uint8_t SC_ZERO[32] = {0};
static void sc_mul(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { oracle_sc_muladd(out, a, b, SC_ZERO); }      /* scalar.rs:132 */
static void sc_sub(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { oracle_sc_muladd(out, SC_LM1, b, a); }        /* scalar.rs:162: a - b */
/* Scalar::inv (scalar.rs:192-214): a^(L-2) by square-and-multiply */
void oracle_sc_invert(uint8_t out[32], const uint8_t a[32])
{
    uint8_t e[32], r[32] = {1};
    memcpy(e, SC_LM1, 32); e[0] -= 1;   /* L - 2 */
    for (int i = 252; i >= 0; i--) {
        sc_mul(r, r, r);
        if ((e[i >> 3] >> (i & 7)) & 1) sc_mul(r, r, a);
    }
    memcpy(out, r, 32);
}

/* ------------------------------------------------------------------ rabin VSS, DSS, interpolation in the exponent */
/* share/vss/rabin/vss.rs:889-900: f*G + g*H == eval(idx).  1 = verifies, 0 = does not, -1 = a point does not decode */
int oracle_rabin_verify_deal(const uint8_t *commits32, int t, uint32_t idx, const uint8_t f[32], const uint8_t g[32], const uint8_t h32[32])
{
    oracle_init();
    ge_p3 *c = (ge_p3 *)malloc(sizeof(ge_p3) * (size_t)t), v, fig, gih, H, ci;
    for (int j = 0; j < t; j++) if (!ge_frombytes(&c[j], commits32 + 32 * j)) { free(c); return -1; }
    if (!ge_frombytes(&H, h32)) { free(c); return -1; }
    ge_scalarmult_base(&fig, f);
    ge_scalarmult(&gih, g, &H);
    pt_add(&ci, &fig, &gih);
    pubpoly_eval_p3(&v, c, t, idx);
    free(c);
    return pt_eq(&ci, &v);
}
/* sign/dss/dss_sig.rs:263-273 with hash_sig :312-326: partial*B == random.eval(idx) + H(R || A || msg) * long.eval(idx) */
int oracle_dss_partial_check(const uint8_t *rand32, const uint8_t *long32, int t, uint32_t idx, const uint8_t *msg, size_t mlen, const uint8_t partial[32], uint8_t hash_out[32])
{
    oracle_init();
    ge_p3 *r = (ge_p3 *)malloc(sizeof(ge_p3) * (size_t)t), *l = (ge_p3 *)malloc(sizeof(ge_p3) * (size_t)t), rs, ls, right, left;
    int ok = 1;
    for (int j = 0; j < t; j++) if (!ge_frombytes(&r[j], rand32 + 32 * j) || !ge_frombytes(&l[j], long32 + 32 * j)) ok = 0;
    if (!ok) { free(r); free(l); return -1; }
    uint8_t rb[32], ab[32], h[32];
    ge_p3_tobytes(rb, &r[0]); ge_p3_tobytes(ab, &l[0]);   /* marshal_to of the two free coefficients */
    challenge(h, rb, ab, msg, mlen);
    if (hash_out) memcpy(hash_out, h, 32);
    pubpoly_eval_p3(&rs, r, t, idx);
    pubpoly_eval_p3(&ls, l, t, idx);
    ge_scalarmult(&right, h, &ls);
    pt_add(&right, &rs, &right);
    ge_scalarmult_base(&left, partial);
    free(r); free(l);
    return pt_eq(&left, &right);
}
/* recover_commit (share/poly.rs:566-603) on the k shares (idx[i], points[i]) already selected by xy_commit; 0 = a point does not decode */
int oracle_recover_commit(uint8_t out[32], int k, const uint32_t *idx, const uint8_t *points32)
{
    oracle_init();
    ge_p3 acc, P, tmp;
    ge_p3_0(&acc);
    for (int i = 0; i < k; i++) {
        uint8_t num[32] = {1}, den[32] = {1}, xi[32], xj[32], d[32], inv[32];
        sc_from_u64(xi, 1 + (uint64_t)idx[i]);
        for (int j = 0; j < k; j++) {
            if (j == i) continue;
            sc_from_u64(xj, 1 + (uint64_t)idx[j]);
            sc_mul(num, num, xj);
            sc_sub(d, xj, xi);
            sc_mul(den, den, d);
        }
        oracle_sc_invert(inv, den);
        sc_mul(num, num, inv);                      /* num.div(num, den) */
        if (!ge_frombytes(&P, points32 + 32 * i)) return 0;
        ge_scalarmult(&tmp, num, &P);
        pt_add(&acc, &acc, &tmp);
    }
    ge_p3_tobytes(out, &acc);
    return 1;
}
/* recover_pub_poly (share/poly.rs:607-635) with lagrange_basis (:640-671): out = k encodings; 0 = a point does not decode */
int oracle_recover_pub_poly(uint8_t *out, int k, const uint32_t *idx, const uint8_t *points32)
{
    oracle_init();
    ge_p3 *acc = (ge_p3 *)malloc(sizeof(ge_p3) * (size_t)k), Y, tmp;
    uint8_t *basis = (uint8_t *)malloc(32 * (size_t)(k + 1)), *next = (uint8_t *)malloc(32 * (size_t)(k + 1));
    int ok = 1;
    for (int j = 0; j < k && ok; j++) {
        /* basis = prod_{m != j} (x - x_m), acc_s = prod 1 / (x_j - x_m) */
        uint8_t xj[32], xm[32], d[32], accs[32] = {1};
        int deg = 0;
        memset(basis, 0, 32 * (size_t)(k + 1)); basis[0] = 1;
        sc_from_u64(xj, 1 + (uint64_t)idx[j]);
        for (int m = 0; m < k; m++) {
            if (m == j) continue;
            sc_from_u64(xm, 1 + (uint64_t)idx[m]);
            uint8_t neg[32];
            sc_sub(neg, SC_ZERO, xm);               /* minus_const: [-x_m, 1] */
            memset(next, 0, 32 * (size_t)(k + 1));
            for (int i = 0; i <= deg; i++) {
                uint8_t p[32];
                sc_mul(p, basis + 32 * i, neg);
                oracle_sc_muladd(next + 32 * i, p, (const uint8_t[32]){1}, next + 32 * i);
                oracle_sc_muladd(next + 32 * (i + 1), basis + 32 * i, (const uint8_t[32]){1}, next + 32 * (i + 1));
            }
            deg++;
            memcpy(basis, next, 32 * (size_t)(k + 1));
            sc_sub(d, xj, xm);
            oracle_sc_invert(d, d);
            sc_mul(accs, accs, d);
        }
        if (!ge_frombytes(&Y, points32 + 32 * j)) { ok = 0; break; }
        for (int i = 0; i < k; i++) {
            uint8_t c[32];
            sc_mul(c, basis + 32 * i, accs);
            ge_scalarmult(&tmp, c, &Y);              /* basis.commit(Some(y_j)) */
            if (j == 0) acc[i] = tmp; else pt_add(&acc[i], &acc[i], &tmp);
        }
    }
    if (ok) for (int i = 0; i < k; i++) ge_p3_tobytes(out + 32 * i, &acc[i]);
    free(acc); free(basis); free(next);
    return ok;
}

/* ------------------------------------------------------------------ threaded batch drivers (CPU baseline: one thread per core over disjoint index ranges) */
typedef struct {
    int kind; size_t lo, hi;
    const uint8_t *a, *b, *c, *h; const uint64_t *off; const uint32_t *idx; uint8_t *out; int t; size_t mlen;
} job_t;

static void *job_run(void *arg)
{
    job_t *j = (job_t *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        switch (j->kind) {
        case 0: oracle_mul_base(j->out + 32 * i, j->a + 32 * i); break;
        case 1: { int ok = oracle_mul(j->out + 32 * i, j->a + 32 * i, j->b + 32 * i); if (!ok) memset(j->out + 32 * i, 0, 32); } break;
        case 2: j->out[i] = (uint8_t)oracle_eddsa_verify(j->a + 32 * i, j->b + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), j->c + 64 * i, 64); break;
        case 3: j->out[i] = (uint8_t)oracle_schnorr_verify(j->a + 32 * i, j->b + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), j->c + 64 * i, 64); break;
        case 4: j->out[i] = (uint8_t)(oracle_vss_verify_deal(j->a, j->t, j->idx[i], j->b + 32 * i) == 1); break;
        case 5: j->out[i] = (uint8_t)(oracle_rabin_verify_deal(j->a, j->t, j->idx[i], j->b + 32 * i, j->c + 32 * i, j->h) == 1); break;
        case 6: j->out[i] = (uint8_t)(oracle_dss_partial_check(j->a, j->h, j->t, j->idx[i], j->c, j->mlen, j->b + 32 * i, NULL) == 1); break;
        }
    }
    return NULL;
}
static void run_jobs(job_t proto, size_t n, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    oracle_init();
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nthreads);
    for (int k = 0; k < nthreads; k++) {
        jobs[k] = proto;
        jobs[k].lo = n * (size_t)k / (size_t)nthreads;
        jobs[k].hi = n * (size_t)(k + 1) / (size_t)nthreads;
        pthread_create(&th[k], NULL, job_run, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th); free(jobs);
}
void oracle_mul_base_batch(size_t n, const uint8_t *scalars, uint8_t *out, int nthreads) { job_t p = {0}; p.kind = 0; p.a = scalars; p.out = out; run_jobs(p, n, nthreads); }
void oracle_mul_batch(size_t n, const uint8_t *scalars, const uint8_t *points, uint8_t *out, int nthreads) { job_t p = {0}; p.kind = 1; p.a = scalars; p.b = points; p.out = out; run_jobs(p, n, nthreads); }
void oracle_eddsa_verify_batch(size_t n, const uint8_t *pk, const uint8_t *msg, const uint64_t *msg_off, const uint8_t *sig, uint8_t *status, int nthreads)
{ job_t p = {0}; p.kind = 2; p.a = pk; p.b = msg; p.off = msg_off; p.c = sig; p.out = status; run_jobs(p, n, nthreads); }
void oracle_schnorr_verify_batch(size_t n, const uint8_t *pk, const uint8_t *msg, const uint64_t *msg_off, const uint8_t *sig, uint8_t *status, int nthreads)
{ job_t p = {0}; p.kind = 3; p.a = pk; p.b = msg; p.off = msg_off; p.c = sig; p.out = status; run_jobs(p, n, nthreads); }
/* one polynomial, m (idx, share) pairs */
void oracle_vss_verify_batch(const uint8_t *commits32, int t, size_t m, const uint32_t *idx, const uint8_t *shares, uint8_t *verdict, int nthreads)
{ job_t p = {0}; p.kind = 4; p.a = commits32; p.t = t; p.idx = idx; p.b = shares; p.out = verdict; run_jobs(p, m, nthreads); }

/* one rabin polynomial, m (idx, f share, g share) triples */
void oracle_rabin_verify_batch(const uint8_t *commits32, int t, const uint8_t h32[32], size_t m, const uint32_t *idx, const uint8_t *f, const uint8_t *g, uint8_t *verdict, int nthreads)
{ job_t p = {0}; p.kind = 5; p.a = commits32; p.t = t; p.idx = idx; p.b = f; p.c = g; p.h = h32; p.out = verdict; run_jobs(p, m, nthreads); }
/* one DSS session, m (idx, partial) pairs */
void oracle_dss_partial_batch(const uint8_t *rand32, const uint8_t *long32, int t, const uint8_t *msg, size_t mlen, size_t m, const uint32_t *idx, const uint8_t *partials, uint8_t *verdict, int nthreads)
{ job_t p = {0}; p.kind = 6; p.a = rand32; p.h = long32; p.t = t; p.idx = idx; p.b = partials; p.c = msg; p.mlen = mlen; p.out = verdict; run_jobs(p, m, nthreads); }

/* Sum_i Point::mul(s_i, P_i) folded with Point::add; returns 0 if a point fails to decode */
int oracle_msm(uint8_t out[32], size_t n, const uint8_t *scalars, const uint8_t *points)
{
    oracle_init();
    ge_p3 acc, A, h;
    ge_p3_0(&acc);
    for (size_t i = 0; i < n; i++) {
        if (!ge_frombytes(&A, points + 32 * i)) return 0;
        ge_scalarmult(&h, scalars + 32 * i, &A);
        pt_add(&acc, &acc, &h);
    }
    ge_p3_tobytes(out, &acc);
    return 1;
}
