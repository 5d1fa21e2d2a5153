// half.cuh — half-size scalars for the signature verifiers.
//
// The reference checks  s*B == R + h*A  on canonical encodings (eddsa_sig.rs:201-210,
// schnorr_sig.rs:96-106) with a 253-bit h: 253 doublings per signature.  Here the SAME verdict is
// obtained from 128 doublings.  Let N = 8L (the order of the whole curve group) and let (u, v) be a
// short vector of the lattice { (u, v) : v = u*h (mod N) } with u ODD.  With D = s*B - h*A - R:
//
//     u*D = (u*s mod L)*B - v*A - u*R            (ord B = L; ord A | N, so u*h*A = v*A exactly —
//                                                 also for keys that carry a small-order component)
//     u*D = 0  <=>  D = 0                          (gcd(u, 8L) = 1: u is odd and 0 < u < L)
//
// so "u*D is the identity" is equivalent to the reference's equation for EVERY input, not only for
// honest ones (the lattice is taken modulo 8L, not L, precisely so that torsion cannot slip through).
// u and |v| are about sqrt(8L) = 2^127.5.  This is the verification trick of Antipa, Brown, Gallant,
// Lambert, Struik and Vanstone ("Accelerated verification of ECDSA signatures", SAC 2005) restated for a
// cofactor-8 curve and a cofactorless verifier.
//
// sc_half runs the (binary long-division form of the) extended Euclidean algorithm on (N, h) and keeps
// the best vector with an odd u it meets; (u, v) = (1, h) is the starting candidate, so a result always
// exists and at worst costs what the full-length path costs.  h is public: nothing here is constant-time.
#pragma once
#include "fe.cuh"
#include "sc.cuh"

// N = 8L
#define KB_8L_WORDS {0xe7ae9f68u, 0xc09318d2u, 0x17bce6b2u, 0xa6f7cef5u, 0u, 0u, 0u, 0x80000000u}

KB_FN int kb_clz32(uint32_t x)
{
#if defined(KB_HOST_EMU)
    return x ? __builtin_clz(x) : 32;
#else
    return __clz((int)x);
#endif
}
// upper word of (hi:lo) << s, s in [0, 32)
KB_FN uint32_t kb_shf_l(uint32_t lo, uint32_t hi, uint32_t s)
{
#if defined(KB_HOST_EMU)
    return s ? ((hi << s) | (lo >> (32 - s))) : hi;
#else
    return __funnelshift_l(lo, hi, s);
#endif
}
// bit length of a 256-bit integer (0 for 0): binary search over the words
KB_FN int kb_bitlen8(const uint32_t* a)
{
    const uint32_t hi4 = a[4] | a[5] | a[6] | a[7];
    const bool h = hi4 != 0;
    const uint32_t b0 = h ? a[4] : a[0], b1 = h ? a[5] : a[1], b2 = h ? a[6] : a[2], b3 = h ? a[7] : a[3];
    const bool m = (b2 | b3) != 0;
    const uint32_t c0 = m ? b2 : b0, c1 = m ? b3 : b1;
    const bool l = c1 != 0;
    const uint32_t w = l ? c1 : c0;
    const int base = (h ? 128 : 0) + (m ? 64 : 0) + (l ? 32 : 0);
    return w ? base + 32 - kb_clz32(w) : 0;
}
// y = a << s truncated to 256 bits, s in [0, 256)
KB_FN void kb_shl8(uint32_t* y, const uint32_t* a, int s)
{
    const uint32_t bs = (uint32_t)s & 31u;
    KB_UNROLL
    for (int i = 7; i >= 1; i--) y[i] = kb_shf_l(a[i - 1], a[i], bs);
    y[0] = a[0] << bs;
    // whole words: almost always none (quotients of a Euclidean run are small)
    KB_NOUNROLL
    for (int k = s >> 5; k > 0; k--) {
        KB_UNROLL
        for (int i = 7; i >= 1; i--) y[i] = y[i - 1];
        y[0] = 0;
    }
}
KB_FN void kb_shr1_8(uint32_t* y)
{
    KB_UNROLL
    for (int i = 0; i < 7; i++) y[i] = (y[i] >> 1) | (y[i + 1] << 31);
    y[7] >>= 1;
}

struct kb_halfsc {
    uint32_t u[8];   // odd, > 0
    uint32_t v[8];   // |v|
    uint32_t vneg;   // v = -|v| when set:   v == u*h (mod 8L)
    int bits;        // max(bitlen(u), bitlen(|v|))
};

// Candidates are only looked at once the remainders are below 2^KB_HALF_TRACK: the short vectors live where
// both coordinates are near 2^128, and the first hundred iterations then carry no bookkeeping at all.
#ifndef KB_HALF_TRACK
#define KB_HALF_TRACK 144
#endif

// Lehmer acceleration: while the remainders are far above 2^128 nothing has to be looked at, so the leading 63
// bits of both rows are reduced on their own (64-bit arithmetic, a few instructions per quotient bit) while a
// 2x2 cofactor matrix with entries below 2^31 is accumulated; the matrix is then applied to the full rows at
// once.  Quotient bits are only taken when they are provably not too large for the FULL rows (interval bounds
// from the cofactors), so the rows stay non-negative and every row remains a lattice vector whatever the
// truncation does; an under-estimated quotient merely leaves work for the next step.
#ifndef KB_HALF_LEHMER_MIN
#define KB_HALF_LEHMER_MIN 152   // batches stop before the rows enter the tracked range
#endif

KB_FN int kb_clz64(uint64_t x) { return (x >> 32) ? kb_clz32((uint32_t)(x >> 32)) : 32 + kb_clz32((uint32_t)x); }
// out = a*x - b*y (mod 2^256); the caller guarantees the true value lies in [0, 2^256)
KB_FN void kb_mulsub8(uint32_t* out, uint32_t a, const uint32_t* x, uint32_t b, const uint32_t* y)
{
    uint64_t cp = 0, cq = 0, br = 0;
    KB_UNROLL
    for (int i = 0; i < 8; i++) {
        const uint64_t p = (uint64_t)a * x[i] + cp;
        const uint64_t q = (uint64_t)b * y[i] + cq;
        cp = p >> 32;
        cq = q >> 32;
        const uint64_t d = (uint64_t)(uint32_t)p - (uint32_t)q - br;
        out[i] = (uint32_t)d;
        br = (d >> 32) & 1u;
    }
}
// out = a*x + b*y (mod 2^256)
KB_FN void kb_muladd8(uint32_t* out, uint32_t a, const uint32_t* x, uint32_t b, const uint32_t* y)
{
    uint64_t c = 0;
    KB_UNROLL
    for (int i = 0; i < 8; i++) {
        const uint64_t p = (uint64_t)a * x[i] + (uint32_t)c;
        const uint64_t q = (uint64_t)b * y[i] + (c >> 32);
        const uint64_t s = (p & 0xffffffffu) + (q & 0xffffffffu);
        out[i] = (uint32_t)s;
        c = (p >> 32) + (q >> 32) + (s >> 32);   // < 2^33: low word and a small high part, re-split above
    }
}

// The run is ONE loop whose body is either a Lehmer batch or a single binary long-division step followed by a
// predicated exchange of the two rows: the lanes of a warp are at different places of their Euclidean
// sequences, and a loop nest would make every lane pay for the slowest one.
// Invariants: r0 >= r1; r0 = -/+ t0 * h, r1 = +/- t1 * h (mod 8L), the t's are magnitudes and the signs
// alternate from row to row.  Every intermediate row is a lattice vector, hence a candidate.
KB_FN void sc_half(kb_halfsc& o, const uint32_t* h)
{
    const uint32_t n8l[8] = KB_8L_WORDS;
    uint32_t r0[8], r1[8], t0[8], t1[8];
    KB_UNROLL
    for (int i = 0; i < 8; i++) {
        r0[i] = n8l[i];
        r1[i] = h[i];
        t0[i] = 0;
        t1[i] = (i == 0) ? 1u : 0u;
        o.u[i] = t1[i];
        o.v[i] = h[i];
    }
    // row 0 starts as (8L, 0): 8L = -0*h;  row 1 as (h, 1): h = +1*h
    uint32_t sg0neg = 1;
    int la = 256, lb = kb_bitlen8(h), lt1 = 1;
    o.vneg = 0;
    o.bits = lb > 1 ? lb : 1;
    KB_NOUNROLL
    while (lb != 0) {
        // every later vector has |u| >= t1: once bitlen(t1) reaches the best cost nothing can improve
        // (lt1 may lag behind while the rows are not tracked, which only delays the exit)
        if (la <= KB_HALF_TRACK && lt1 >= o.bits) break;
        uint32_t y[8], d[8];
        if (la > KB_HALF_LEHMER_MIN && la - lb < 24) {
            // ---- Lehmer batch on the leading 63 bits (x0 < 2^63, so the bounds below cannot overflow)
            kb_shl8(y, r0, 256 - la);
            uint64_t x0 = (((uint64_t)y[7] << 32) | y[6]) >> 1;
            kb_shl8(y, r1, 256 - la);
            uint64_t x1 = (((uint64_t)y[7] << 32) | y[6]) >> 1;
            // current row0 = a*r0 - b*r1, row1 = d*r1 - c*r0 (par = 0) or the negatives of both forms (par = 1);
            // true row / 2^(la-63) lies within (x - (sum of its cofactors), x + (sum of its cofactors))
            uint32_t ca = 1, cb = 0, cc = 0, cd = 1, par = 0;
            int steps = 0;
            const int xstop = KB_HALF_LEHMER_MIN - (la - 63);   // keep row0 above 2^KB_HALF_LEHMER_MIN
            KB_NOUNROLL
            for (;;) {
                const uint64_t e0 = (uint64_t)ca + cb, e1 = (uint64_t)cc + cd;
                if (x0 <= e0) break;
                const uint64_t lo0 = x0 - e0, hi1 = x1 + e1;
                if (hi1 > lo0) break;
                if (64 - kb_clz64(lo0) <= xstop) break;
                int s = kb_clz64(hi1) - kb_clz64(lo0);
                if ((hi1 << s) > lo0) s -= 1;
                if (s > 30) break;
                const uint64_t na = (uint64_t)ca + ((uint64_t)cc << s), nb = (uint64_t)cb + ((uint64_t)cd << s);
                if ((na | nb) >> 31) break;
                x0 -= x1 << s;
                ca = (uint32_t)na;
                cb = (uint32_t)nb;
                steps++;
                if (x0 < x1) {
                    const uint64_t tx = x0; x0 = x1; x1 = tx;
                    uint32_t tc = ca; ca = cc; cc = tc;
                    tc = cb; cb = cd; cd = tc;
                    par ^= 1u;
                }
            }
            if (steps != 0) {
                if (par == 0) {
                    kb_mulsub8(y, ca, r0, cb, r1);
                    kb_mulsub8(d, cd, r1, cc, r0);
                } else {
                    kb_mulsub8(y, cb, r1, ca, r0);
                    kb_mulsub8(d, cc, r0, cd, r1);
                }
                KB_UNROLL
                for (int i = 0; i < 8; i++) { r0[i] = y[i]; r1[i] = d[i]; }
                kb_muladd8(y, ca, t0, cb, t1);
                kb_muladd8(d, cc, t0, cd, t1);
                KB_UNROLL
                for (int i = 0; i < 8; i++) { t0[i] = y[i]; t1[i] = d[i]; }
                sg0neg ^= par;
                la = kb_bitlen8(r0);
                lb = kb_bitlen8(r1);
                bool swap = la < lb;
                if (la == lb) swap = kb_sub8(d, r0, r1) != 0;
                if (swap) {
                    KB_UNROLL
                    for (int i = 0; i < 8; i++) {
                        const uint32_t a = r0[i], b = t0[i];
                        r0[i] = r1[i];
                        r1[i] = a;
                        t0[i] = t1[i];
                        t1[i] = b;
                    }
                    const int l = la;
                    la = lb;
                    lb = l;
                    sg0neg ^= 1u;
                }
                continue;
            }
        }
        // ---- one exact binary long-division step
        int s = la - lb;
        kb_shl8(y, r1, s);
        if (kb_sub8(d, r0, y)) {   // r1 << s overshoots: s >= 1 here because r0 >= r1
            s -= 1;
            kb_shr1_8(y);
            kb_sub8(d, r0, y);
        }
        kb_shl8(y, t1, s);
        kb_add8(t0, t0, y);
        KB_UNROLL
        for (int i = 0; i < 8; i++) r0[i] = d[i];
        la = kb_bitlen8(r0);
        int lt0 = 0;
        if (la <= KB_HALF_TRACK) {
            lt0 = kb_bitlen8(t0);
            const int c = la > lt0 ? la : lt0;
            if ((t0[0] & 1u) && c < o.bits) {
                KB_UNROLL
                for (int i = 0; i < 8; i++) {
                    o.u[i] = t0[i];
                    o.v[i] = r0[i];
                }
                o.vneg = sg0neg;
                o.bits = c;
            }
        }
        bool swap = la < lb;
        if (la == lb) swap = kb_sub8(d, r0, r1) != 0;
        if (swap) {   // the division step is complete: exchange the rows
            KB_UNROLL
            for (int i = 0; i < 8; i++) {
                const uint32_t a = r0[i], b = t0[i];
                r0[i] = r1[i];
                r1[i] = a;
                t0[i] = t1[i];
                t1[i] = b;
            }
            const int l = la;
            la = lb;
            lb = l;
            lt1 = lt0;
            sg0neg ^= 1u;
        }
    }
}
