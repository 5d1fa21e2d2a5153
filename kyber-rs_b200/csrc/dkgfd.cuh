// dkgfd.cuh — a whole DKG deal-verification round by FORWARD DIFFERENCES.
//
// The reference checks every share with its own PubPoly::eval (share/poly.rs:457-469): for dealer d and verifier
// i, a Horner run over the t commitments at x = i + 1, i.e. n * t small-scalar steps per dealer (k_poly_eval does
// exactly that, at 96 % of the multiplier pipe).  But the n evaluation points of one dealer are CONSECUTIVE integers,
// and for consecutive points a polynomial of degree t - 1 obeys  Δ^k P(x + 1) = Δ^k P(x) + Δ^(k+1) P(x):  once the t
// forward differences at x = 1 are known, every further evaluation costs t - 1 point ADDITIONS instead of t
// multiplications by x.  All identities used are integer-linear combinations of the commitments, so they hold in any
// abelian group — in particular for commitments that carry a small-order component; the only
// "division" is avoided by going through the Newton form:
//
//   A  Newton coefficients a_k of P on the nodes 1, 2, ..., t  (P(x) = sum_k a_k (x-1)(x-2)...(x-k)) by repeated
//      synthetic division:  step m = 1..t-1:  q[j] += m * q[j+1]  for j = t-2 down to m-1;  then a_k = q[k].
//      t(t-1)/2 cells, each ONE kb_horner_step with a multiplier <= t; cell (m, j) depends on (m, j+1) and (m-1, j),
//      so the cells with m + (t-2-j) = w form wavefront w, w = 1..t-1, and a wavefront is one launch over all dealers.
//      Two arrays hold the rows of even / odd m (Q_m[j] = Q_(m-1)[j] + m * Q_m[j+1]).
//   B  Δ^k P(1) = k! * a_k, with k! taken mod 8L (the group has exponent 8L) — one full scalar multiplication each.
//   C  n steps: record P(x) = Δ^0, then Δ^k += Δ^(k+1) for all k (ping-pong between two arrays).
//   D  verdict(d, i) = [P_d(i+1) == share * B], projectively (as k_poly_eval does), share * B through the comb.
//
// For n = 1024, t = 683 this is 1.6 x 10^9 multiplies per dealer instead of 4.75 x 10^9.
// Arrays are laid out [k][dealer] (dealer fastest): a warp holds 32 dealers and ONE (m, j), so the NAF of the
// multiplier is warp-uniform and its loads are contiguous.
#pragma once
#include "ops.cuh"
#include "poly.cuh"

// ---- per-cell bodies (KB_FN: also compiled by the host emulation of tests/emu) ------------------------------
// A: one cell of the Newton conversion, v = m * v + prev
KB_FN void kb_fd_newton_cell(ge_p3& v, const ge_p3& prev, uint64_t m)
{
    ge_cached c;
    ge_to_cached(c, prev);
    kb_naf xn;
    kb_naf_from(xn, m);
    kb_horner_step(v, xn, c);
}
// B: h = fact * a, fact = 8 words of magnitude (<= 4L) + 1 word of sign; `tbl` is the caller's 8-entry scratch table
KB_FN void kb_fd_scale_cell(ge_p3& h, const ge_p3& a, const uint32_t* fact9, ge_cached* tbl)
{
    int8_t e[64];
    sc_recode16(e, fact9);
    ge_build_table8(tbl, a);
    ge_scalarmult<false>(h, e, tbl);
    if (fact9[8]) {
        fe_neg(h.X, h.X);
        fe_neg(h.T, h.T);
    }
}
// C: p += q
KB_FN void kb_fd_step_cell(ge_p3& p, const ge_p3& q)
{
    ge_cached c;
    ge_to_cached(c, q);
    ge_add<true>(p, p, c);
}

// fact[k] = k! mod 8L in signed form: 8 words of magnitude (<= 4L, inside the domain of the radix-16 recoding) + 1 word
// of sign.  Host integers only (table construction, like the window counts of the MSM plan).
static inline void kb_factorials_mod_8l(size_t t, uint32_t* out)
{
    const uint64_t N[4] = {0xc09318d2e7ae9f68ull, 0xa6f7cef517bce6b2ull, 0ull, 0x8000000000000000ull};
    const uint64_t H[4] = {0x60498c6973d74fb4ull, 0x537be77a8bde7359ull, 0ull, 0x4000000000000000ull};   // N / 2 = 4L
    uint64_t x[4] = {1, 0, 0, 0};
    for (size_t k = 0; k < t; k++) {
        if (k >= 2) {
            uint64_t y[5];
            unsigned __int128 c = 0;
            for (int i = 0; i < 4; i++) {
                c += (unsigned __int128)x[i] * (uint64_t)k;
                y[i] = (uint64_t)c;
                c >>= 64;
            }
            y[4] = (uint64_t)c;
            // N = 2^255 + (a 128-bit number): floor(y / 2^255) is the quotient or one more
            const uint64_t q = (y[4] << 1) | (y[3] >> 63);
            unsigned __int128 mc = 0;
            uint64_t qn[5];
            for (int i = 0; i < 4; i++) {
                mc += (unsigned __int128)N[i] * q;
                qn[i] = (uint64_t)mc;
                mc >>= 64;
            }
            qn[4] = (uint64_t)mc;
            uint64_t borrow = 0;
            for (int i = 0; i < 5; i++) {
                const unsigned __int128 dd = (unsigned __int128)y[i] - qn[i] - borrow;
                y[i] = (uint64_t)dd;
                borrow = (uint64_t)(dd >> 64) & 1u;
            }
            if (borrow) {   // one N too many: add it back
                unsigned __int128 a = 0;
                for (int i = 0; i < 5; i++) {
                    a += (unsigned __int128)y[i] + (i < 4 ? N[i] : 0);
                    y[i] = (uint64_t)a;
                    a >>= 64;
                }
            }
            for (int i = 0; i < 4; i++) x[i] = y[i];
        }
        // signed representative
        bool big = false;
        for (int i = 3; i >= 0; i--) {
            if (x[i] != H[i]) {
                big = x[i] > H[i];
                break;
            }
        }
        uint64_t m[4];
        if (big) {
            uint64_t borrow = 0;
            for (int i = 0; i < 4; i++) {
                const unsigned __int128 dd = (unsigned __int128)N[i] - x[i] - borrow;
                m[i] = (uint64_t)dd;
                borrow = (uint64_t)(dd >> 64) & 1u;
            }
        } else {
            for (int i = 0; i < 4; i++) m[i] = x[i];
        }
        for (int i = 0; i < 4; i++) {
            out[9 * k + 2 * i] = (uint32_t)m[i];
            out[9 * k + 2 * i + 1] = (uint32_t)(m[i] >> 32);
        }
        out[9 * k + 8] = big ? 1u : 0u;
    }
}

#if !defined(KB_HOST_EMU)
#include "kernels.cuh"

#define KB_FD_AT(arr, k, d, nd) ((arr) + (((size_t)(k) * (nd) + (d)) * 32))
#ifndef KB_FD_NEWTON_THREADS
#define KB_FD_NEWTON_THREADS 128
#endif
#ifndef KB_FD_NEWTON_MINBLOCKS
#define KB_FD_NEWTON_MINBLOCKS 3
#endif

__device__ __forceinline__ void kb_fd_load(ge_p3& p, const uint32_t* o)
{
    kb_load_fe(p.X, o);
    kb_load_fe(p.Y, o + 8);
    kb_load_fe(p.Z, o + 16);
    kb_load_fe(p.T, o + 24);
}
__device__ __forceinline__ void kb_fd_store(uint32_t* o, const ge_p3& p)
{
    kb_store_fe(o, p.X);
    kb_store_fe(o + 8, p.Y);
    kb_store_fe(o + 16, p.Z);
    kb_store_fe(o + 24, p.T);
}

// Q0[j][d] = commitment j of dealer d (decoded); the top coefficient also into Q1 (it is never updated, and both
// parities read it); dealer_bad[d] |= 1 if any commitment does not decode
__global__ void __launch_bounds__(KB_THREADS) k_fd_init(size_t nd, size_t t, const uint8_t* commits, uint32_t* q0, uint32_t* q1, uint32_t* dealer_bad)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nd * t) return;
    const size_t j = idx / nd, d = idx % nd;
    uint32_t w[8];
    kb_load32(w, commits, d * t + j);
    ge_p3 p;
    if (!ge_decompress(p, w)) {
        ge_identity(p);
        atomicOr(dealer_bad + d, 1u);
    }
    kb_fd_store(KB_FD_AT(q0, j, d, nd), p);
    if (j == t - 1) kb_fd_store(KB_FD_AT(q1, j, d, nd), p);
}

// wavefront w of the Newton conversion: cells m = 1..w, j = m + t - 2 - w
__global__ void __launch_bounds__(KB_FD_NEWTON_THREADS, KB_FD_NEWTON_MINBLOCKS) k_fd_newton(size_t nd, size_t t, size_t w, uint32_t* q0, uint32_t* q1)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nd * w) return;
    // the cost of a cell grows with log2(m): the blocks with the large multipliers go first, so that the launch does not
    // end on them (each launch waits for its slowest block)
    const size_t m = w - idx / nd, d = idx % nd;
    const size_t j = m + t - 2 - w;
    uint32_t* qm = (m & 1) ? q1 : q0;         // row m
    const uint32_t* qp = (m & 1) ? q0 : q1;   // row m - 1
    ge_p3 v, prev;
    kb_fd_load(v, KB_FD_AT(qm, j + 1, d, nd));
    kb_fd_load(prev, KB_FD_AT(qp, j, d, nd));
    kb_fd_newton_cell(v, prev, (uint64_t)m);   // v = m * v + prev
    kb_fd_store(KB_FD_AT(qm, j, d, nd), v);
}

// D[k][d] = (k! mod 8L) * a_k,  a_k = row (k+1) entry k (k <= t-2), a_(t-1) = the top coefficient.
// fact[k] = 8 words magnitude (<= 4L) + 1 word sign.
__global__ void __launch_bounds__(KB_THREADS) k_fd_scale(size_t nd, size_t t, const uint32_t* q0, const uint32_t* q1, const uint32_t* fact, uint32_t* dout)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < nd * t;   // ge_scalarmult holds block barriers: tail threads redo the last item
    if (!live) idx = nd * t - 1;
    const size_t k = idx / nd, d = idx % nd;
    const uint32_t* src = ((k + 1) & 1) ? q1 : q0;
    ge_p3 a, h;
    kb_fd_load(a, KB_FD_AT(src, k, d, nd));
    uint32_t f9[9];
#pragma unroll
    for (int q = 0; q < 9; q++) f9[q] = fact[9 * k + q];
    ge_cached tbl[8];
    kb_fd_scale_cell(h, a, f9, tbl);
    if (k < 2) h = a;   // 0! = 1! = 1
    if (live) kb_fd_store(KB_FD_AT(dout, k, d, nd), h);
}

// one evaluation point: evals[d][i] = Δ^0 (the value at x = i + 1), then dst[k] = src[k] + src[k+1] for the `live`
// lowest orders (the host drops the orders that can no longer reach a value).  ncu: multiplier pipe 57 % busy,
// long-scoreboard stall 2.5 per issue (a load, one addition, a store per thread).  Two or three elements per thread
// with all loads issued first were SLOWER (251 / 256 ms per round against 239: 168 / 254 registers), and so were
// launch bounds for 5 or 6 blocks per SM (245 ms).
__global__ void __launch_bounds__(KB_THREADS) k_fd_step(size_t nd, size_t t, size_t live, size_t n, size_t i, const uint32_t* src, uint32_t* dst, uint32_t* evals)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nd * live) return;
    const size_t k = idx / nd, d = idx % nd;
    ge_p3 p;
    kb_fd_load(p, KB_FD_AT(src, k, d, nd));
    if (k == 0) kb_store_xyz(evals, d * n + i, p);
    if (k + 1 < t) {
        ge_p3 q;
        kb_fd_load(q, KB_FD_AT(src, k + 1, d, nd));
        kb_fd_step_cell(p, q);
    }
    kb_fd_store(KB_FD_AT(dst, k, d, nd), p);
}

// verdict[d * n + i] = [evals[d][i] == share(d, i) * B] and no undecodable commitment (vss/pedersen/vss.rs:899-912)
__global__ void __launch_bounds__(KB_THREADS) k_fd_check(size_t nd, size_t n, const uint32_t* evals, const uint8_t* shares, const uint32_t* dealer_bad, const ge_precomp* comb, uint8_t* verdict)
{
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nd * n) return;
    const size_t d = slot / n;
    ge_p3 v, h;
    kb_load_fe(v.X, evals + 24 * slot);
    kb_load_fe(v.Y, evals + 24 * slot + 8);
    kb_load_fe(v.Z, evals + 24 * slot + 16);
    uint32_t s[8];
    kb_load32(s, shares, slot);
    ge_scalarmult_base_comb(h, s, comb);
    fe l, r, df;
    fe_mul(l, v.X, h.Z);
    fe_mul(r, h.X, v.Z);
    fe_sub(df, l, r);
    uint32_t same = fe_is_zero(df);
    fe_mul(l, v.Y, h.Z);
    fe_mul(r, h.Y, v.Z);
    fe_sub(df, l, r);
    same &= fe_is_zero(df);
    verdict[slot] = (uint8_t)(same & (dealer_bad[d] ? 0u : 1u));
}
#endif  // !KB_HOST_EMU
