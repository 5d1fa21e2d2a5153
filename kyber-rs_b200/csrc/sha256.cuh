// sha256.cuh — per-thread SHA-256 (FIPS 180-4) over 32-byte records held as little-endian words: what
// session_id (share/vss/pedersen/vss.rs:1069-1090) needs — a hash over 1 + n + t point encodings and a u32.
// The reference calls the third-party `sha2` crate (^0.10.6) through SuiteEd25519::hash (group/edwards25519/suite.rs:93);
// the algorithm is the published standard.
#pragma once
#include <stdint.h>
#include "sha512.cuh"   // KB_CONST, kb_bswap32

KB_CONST uint32_t KB_K256[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

KB_FN uint32_t kb_rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

// one compression; w[16] = the block as big-endian words, consumed as a rolling schedule window
KB_FN void sha256_compress(uint32_t* h, uint32_t* w)
{
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    KB_NOUNROLL
    for (int r = 0; r < 64; r += 16) {
        KB_UNROLL
        for (int i = 0; i < 16; i++) {
            if (r > 0) {
                const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                const uint32_t s0 = kb_rotr32(w15, 7) ^ kb_rotr32(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = kb_rotr32(w2, 17) ^ kb_rotr32(w2, 19) ^ (w2 >> 10);
                w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
            }
            const uint32_t S1 = kb_rotr32(e, 6) ^ kb_rotr32(e, 11) ^ kb_rotr32(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + KB_K256[r + i] + w[i];
            const uint32_t S0 = kb_rotr32(a, 2) ^ kb_rotr32(a, 13) ^ kb_rotr32(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1;
            d = c; c = b; b = a; a = t1 + t2;
        }
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// Streaming state for inputs that arrive as whole 32-byte records (8 little-endian words) and end in up to 7 extra words.
struct kb_sha256 {
    uint32_t h[8];
    uint32_t w[16];
    uint32_t fill;     // words in w
    uint64_t words;    // total words absorbed
};
KB_FN void sha256_init(kb_sha256& s)
{
    const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    KB_UNROLL
    for (int i = 0; i < 8; i++) s.h[i] = iv[i];
    s.fill = 0;
    s.words = 0;
}
// absorb nw little-endian words (byte order of the message = memory order)
KB_FN void sha256_words(kb_sha256& s, const uint32_t* le, int nw)
{
    for (int i = 0; i < nw; i++) {
        s.w[s.fill++] = kb_bswap32(le[i]);
        if (s.fill == 16) {
            sha256_compress(s.h, s.w);
            s.fill = 0;
        }
    }
    s.words += (uint64_t)nw;
}
// absorb one 32-byte record (8 little-endian words); only valid while everything absorbed so far was whole records
// (the buffer is then empty or half full, and every index below is a compile-time constant)
KB_FN void sha256_rec32(kb_sha256& s, const uint32_t* le)
{
    if (s.fill == 0) {
        KB_UNROLL
        for (int i = 0; i < 8; i++) s.w[i] = kb_bswap32(le[i]);
        s.fill = 8;
    } else {
        KB_UNROLL
        for (int i = 0; i < 8; i++) s.w[8 + i] = kb_bswap32(le[i]);
        sha256_compress(s.h, s.w);
        s.fill = 0;
    }
    s.words += 8;
}
// out = digest as 8 little-endian words (digest bytes in memory order)
KB_FN void sha256_final(kb_sha256& s, uint32_t* out)
{
    const uint64_t bits = s.words * 32;
    s.w[s.fill++] = 0x80000000u;
    if (s.fill > 14) {
        while (s.fill < 16) s.w[s.fill++] = 0;
        sha256_compress(s.h, s.w);
        s.fill = 0;
    }
    while (s.fill < 14) s.w[s.fill++] = 0;
    s.w[14] = (uint32_t)(bits >> 32);
    s.w[15] = (uint32_t)bits;
    sha256_compress(s.h, s.w);
    KB_UNROLL
    for (int i = 0; i < 8; i++) out[i] = kb_bswap32(s.h[i]);
}
