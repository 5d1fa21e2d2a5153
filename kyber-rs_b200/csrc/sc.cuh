// sc.cuh — scalars mod L = 2^252 + 27742317777372353535851937790883648493, 8 x 32-bit words.
//
// Replaces the pieces of src/group/edwards25519/scalar.rs that sit on the verify path:
//   Scalar::is_canonical (scalar.rs:54-75)       -> sc_is_canonical
//   Scalar::set_bytes on a 64-byte digest          -> sc_reduce512   (scalar.rs:175,
//       integer_field/integer.rs:386-396: little-endian integer mod L)
//   sc_mul_add (scalar.rs:279)                      -> sc_muladd      ((ab+c) mod L)
// The reference's 21-bit-limb ref10 bodies are not reproduced; the results (fully reduced,
// little-endian) are identical.
#pragma once
#include "fe.cuh"

// L and c = L - 2^252 as 32-bit words
#define KB_L_WORDS {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u}

// 1 iff the little-endian integer s[0..8) is < L
KB_FN uint32_t sc_is_canonical(const uint32_t* s)
{
    const uint32_t l[8] = KB_L_WORDS;
    uint32_t t[8];
    return kb_sub8(t, s, l);
}

// r[0..NA+4) = a[0..NA) * c, c = L - 2^252 (125 bits, 4 words)
template <int NA>
KB_FN void sc_mul_c(uint32_t* r, const uint32_t* a)
{
    const uint32_t c[4] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu};
    KB_UNROLL
    for (int i = 0; i < NA + 4; i++) r[i] = 0;
    KB_UNROLL
    for (int i = 0; i < NA; i++) {
        uint64_t carry = 0;
        KB_UNROLL
        for (int j = 0; j < 4; j++) {
            uint64_t t = (uint64_t)a[i] * c[j] + r[i + j] + carry;
            r[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        r[i + 4] = (uint32_t)carry;
    }
}
// lo[0..8) = x mod 2^252, hi[0..NH) = x >> 252, x has NX words (words beyond NX are zero)
template <int NX, int NH>
KB_FN void sc_split252(uint32_t* lo, uint32_t* hi, const uint32_t* x)
{
    KB_UNROLL
    for (int i = 0; i < 8; i++) lo[i] = (i < NX) ? x[i] : 0u;
    lo[7] &= 0x0fffffffu;
    KB_UNROLL
    for (int i = 0; i < NH; i++) {
        uint32_t a = (i + 7 < NX) ? x[i + 7] : 0u;
        uint32_t b = (i + 8 < NX) ? x[i + 8] : 0u;
        hi[i] = (a >> 28) | (b << 4);
    }
}
// t = a - b (mod L) for a, b in [0, L)
KB_FN void sc_sub_mod(uint32_t* t, const uint32_t* a, const uint32_t* b)
{
    const uint32_t l[8] = KB_L_WORDS;
    uint32_t borrow = kb_sub8(t, a, b);
    uint32_t m = 0u - borrow, lm[8];
    KB_UNROLL
    for (int i = 0; i < 8; i++) lm[i] = l[i] & m;
    kb_add8(t, t, lm);
}
// x[0..16) (512 bits, little-endian words) -> r[0..8) = x mod L.
// 2^252 = -c (mod L): three folds bring the high part below 2^131, then
// x = lo1 - (lo2 - (lo3 - w)) with every parenthesis normalised into [0, L).
KB_FN void sc_reduce512(uint32_t* r, const uint32_t* x)
{
    uint32_t lo1[8], hi1[9], y[13], lo2[8], hi2[5], z[9], lo3[8], hi3[1], w5[5], w[8], t[8];
    sc_split252<16, 9>(lo1, hi1, x);   // hi1 < 2^260
    sc_mul_c<9>(y, hi1);               // < 2^385
    sc_split252<13, 5>(lo2, hi2, y);   // hi2 < 2^133
    sc_mul_c<5>(z, hi2);               // < 2^258
    sc_split252<9, 1>(lo3, hi3, z);    // hi3 < 2^6
    sc_mul_c<1>(w5, hi3);              // < 2^131
    KB_UNROLL
    for (int i = 0; i < 8; i++) w[i] = (i < 5) ? w5[i] : 0u;
    sc_sub_mod(t, lo3, w);
    sc_sub_mod(t, lo2, t);
    sc_sub_mod(r, lo1, t);
}
// s = (a*b + c) mod L on raw 256-bit inputs (scalar.rs:279 sc_mul_add)
KB_FN void sc_muladd(uint32_t* s, const uint32_t* a, const uint32_t* b, const uint32_t* c)
{
    uint32_t x[16];
    KB_UNROLL
    for (int i = 0; i < 16; i++) x[i] = (i < 8) ? c[i] : 0u;
    KB_UNROLL
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        KB_UNROLL
        for (int j = 0; j < 8; j++) {
            uint64_t t = (uint64_t)a[i] * b[j] + x[i + j] + carry;
            x[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        // propagate into the untouched upper words
        KB_UNROLL
        for (int k = i + 8; k < 16; k++) {
            carry += x[k];
            x[k] = (uint32_t)carry;
            carry >>= 32;
        }
    }
    sc_reduce512(s, x);
}

// The integer Point::mul multiplies by: a itself when a[31] <= 127; otherwise the
// reference's top radix-16 digit (nibble 63 + carry) may exceed 8, then matches no table
// entry and is dropped, leaving low252 - carry * 2^252 (possibly negative).
KB_FN void sc_effective(uint32_t* mag, uint32_t& neg, const uint32_t* s)
{
    const uint32_t c8[8] = {0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u, 0x08888888u};
    uint32_t low[8], t[8];
    KB_UNROLL
    for (int i = 0; i < 8; i++) low[i] = s[i];
    low[7] &= 0x0fffffffu;
    kb_add8(t, low, c8);
    const uint32_t carry = t[7] >> 28;  // carry out of the 63 low digits
    const uint32_t top = (s[7] >> 28) + carry;
    neg = 0;
    if (top <= 8) {
        KB_UNROLL
        for (int i = 0; i < 8; i++) mag[i] = s[i];
    } else if (carry == 0) {
        KB_UNROLL
        for (int i = 0; i < 8; i++) mag[i] = low[i];
    } else {
        const uint32_t p252[8] = {0, 0, 0, 0, 0, 0, 0, 0x10000000u};
        kb_sub8(mag, p252, low);
        neg = 1;
    }
}
// r = a^(L-2) mod L — Scalar::inv (scalar.rs:192-214): square-and-multiply over the fixed public exponent
// L - 2, so the operation sequence does not depend on a (0 maps to 0).
KB_FN void sc_invert(uint32_t* r, const uint32_t* a)
{
    const uint32_t lm2[8] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};
    const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t res[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    KB_NOUNROLL
    for (int i = 252; i >= 0; i--) {
        uint32_t t[8];
        sc_muladd(t, res, res, zero);
        if ((lm2[i >> 5] >> (i & 31)) & 1u) sc_muladd(res, t, a, zero);
        else {
            KB_UNROLL
            for (int k = 0; k < 8; k++) res[k] = t[k];
        }
    }
    KB_UNROLL
    for (int k = 0; k < 8; k++) r[k] = res[k];
}
