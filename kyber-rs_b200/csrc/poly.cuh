// poly.cuh — PubPoly::eval (share/poly.rs:457-469) as a short-scalar Horner step.
//
// The reference evaluates v = xi*v + commits[j] with a FULL 253-bit constant-time scalar
// multiplication by xi = 1 + i per coefficient.  xi is a small public integer, so the same
// group element is obtained with a double-and-add over the digits of xi (about 10
// doublings for n <= 1024 instead of 252).  Working with the INTEGER xi — not powers of xi
// reduced mod L — keeps the result exact for commitments that carry a small-order component
// (SURVEY §7-H2).  The digits are the non-adjacent form of xi (on average 1/3 of them are
// non-zero instead of 1/2), computed once per evaluation point.
#pragma once
#include "ge.cuh"

struct kb_naf {
    int8_t d[36];  // d[0] least significant; digits in {-1, 0, 1}; d[len-1] = 1
    int len;
};
// x >= 1, x <= 2^33
KB_FN void kb_naf_from(kb_naf& n, uint64_t x)
{
    int len = 0;
    while (x) {
        int z = 0;
        if (x & 1) {
            z = 2 - (int)(x & 3);  // +1 or -1
            x -= (uint64_t)(int64_t)z;
        }
        n.d[len++] = (int8_t)z;
        x >>= 1;
    }
    n.len = len;
}

// v = x * v + c, x given by its NAF.  T is only computed where the next operation reads it.
KB_FN void kb_horner_step(ge_p3& v, const kb_naf& x, const ge_cached& c)
{
    ge_cached vc;
    ge_to_cached(vc, v);
    ge_p3 acc = v;
    KB_NOUNROLL
    for (int i = x.len - 2; i >= 0; i--) {
        const int d = x.d[i];
        ge_dbl_rt(acc, acc, d != 0 || i == 0);
        if (d != 0) ge_addsub_rt(acc, acc, vc, d < 0, i == 0);
    }
    ge_addsub_rt(v, acc, c, false, true);  // same body as the loop's additions: keeps the kernel loop small
}
// convenience form for one-off use
KB_FN void kb_horner_step(ge_p3& v, uint64_t x, const ge_cached& c)
{
    kb_naf n;
    kb_naf_from(n, x);
    kb_horner_step(v, n, c);
}
