// poly.cuh — PubPoly::eval (share/poly.rs:457-469) as a short-scalar Horner step.
//
// The reference evaluates v = xi*v + commits[j] with a FULL 253-bit constant-time scalar
// multiplication by xi = 1 + i per coefficient.  xi is a small public integer, so the same
// group element is obtained with a plain double-and-add over the bits of xi (about 11
// doublings for n <= 1024 instead of 252).  Working with the INTEGER xi — not powers of xi
// reduced mod L — keeps the result exact for commitments that carry a small-order component
// (SURVEY §7-H2).
#pragma once
#include "ge.cuh"

// v = x * v + c,  x >= 1
KB_FN void kb_horner_step(ge_p3& v, uint64_t x, const ge_cached& c)
{
    ge_cached vc;
    ge_to_cached(vc, v);
    int top = 63;
    while (top > 0 && !((x >> top) & 1)) top--;
    ge_p3 acc = v;
    KB_NOUNROLL
    for (int b = top - 1; b >= 0; b--) {
        ge_dbl<true>(acc, acc);
        if ((x >> b) & 1) ge_add<true>(acc, acc, vc);
    }
    ge_add<true>(v, acc, c);
}
