// sha512.cuh — per-thread SHA-512 of R || A || M (FIPS 180-4), for the EdDSA/Schnorr challenge
// h = SHA512(R || A || M) (sign/eddsa/eddsa_sig.rs:195-198, sign/schnorr/schnorr_sig.rs:128-141,
// sign/dss/dss_sig.rs:312-326).  The reference calls the third-party `sha2` crate (^0.10.6);
// the algorithm is the published standard.
#pragma once
#include <stdint.h>
#include "fe.cuh"

#if defined(KB_HOST_EMU)
#define KB_CONST static const
#else
#define KB_CONST static __device__ __constant__
#endif

KB_CONST uint64_t KB_K512[80] = {
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

KB_FN uint64_t kb_rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
KB_FN uint32_t kb_bswap32(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24); }

// one compression; w[16] is consumed as a rolling schedule window
KB_FN void sha512_compress(uint64_t* h, uint64_t* w)
{
    uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    KB_NOUNROLL
    for (int r = 0; r < 80; r += 16) {
        KB_UNROLL
        for (int i = 0; i < 16; i++) {
            if (r > 0) {
                uint64_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                uint64_t s0 = kb_rotr64(w15, 1) ^ kb_rotr64(w15, 8) ^ (w15 >> 7);
                uint64_t s1 = kb_rotr64(w2, 19) ^ kb_rotr64(w2, 61) ^ (w2 >> 6);
                w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
            }
            uint64_t S1 = kb_rotr64(e, 14) ^ kb_rotr64(e, 18) ^ kb_rotr64(e, 41);
            uint64_t ch = (e & f) ^ (~e & g);
            uint64_t t1 = hh + S1 + ch + KB_K512[r + i] + w[i];
            uint64_t S0 = kb_rotr64(a, 28) ^ kb_rotr64(a, 34) ^ kb_rotr64(a, 39);
            uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint64_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// byte k of the padded message part: msg[k] for k < mlen, 0x80 at k == mlen, 0 after
KB_FN uint64_t kb_msg_be64(const uint8_t* msg, uint64_t pos, uint64_t mlen)
{
    if (pos + 8 <= mlen && (((uintptr_t)(msg + pos)) & 3u) == 0) {
        const uint32_t* p = (const uint32_t*)(msg + pos);
        return ((uint64_t)kb_bswap32(p[0]) << 32) | kb_bswap32(p[1]);
    }
    uint64_t v = 0;
    for (int k = 0; k < 8; k++) {
        uint64_t q = pos + k;
        uint64_t byte = (q < mlen) ? msg[q] : (q == mlen ? 0x80u : 0u);
        v = (v << 8) | byte;
    }
    return v;
}

// digest[0..16) = SHA-512(head || msg) as little-endian 32-bit words of the 64 output bytes — i.e. directly the
// little-endian integer the reference feeds to Scalar::set_bytes.  head = HEAD_WORDS (8 or 16) little-endian
// words of a 32- or 64-byte prefix held in registers (R || A for the challenge, the key prefix for the EdDSA
// nonce, the seed for key derivation).
template <int HEAD_WORDS>
KB_FN void sha512_prefixed(uint32_t* digest, const uint32_t* head, const uint8_t* msg, uint64_t mlen)
{
    constexpr int HB = 4 * HEAD_WORDS;  // prefix bytes: 32 or 64
    uint64_t h[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                     0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    uint64_t w[16];
    const uint64_t total = HB + mlen;                    // data bytes
    const uint64_t nblocks = (total + 17 + 127) / 128;   // with 0x80 and the 128-bit length
    KB_NOUNROLL
    for (uint64_t blk = 0; blk < nblocks; blk++) {
        if (blk == 0) {
            KB_UNROLL
            for (int j = 0; j < HB / 8; j++) w[j] = ((uint64_t)kb_bswap32(head[2 * j]) << 32) | kb_bswap32(head[2 * j + 1]);
            KB_UNROLL
            for (int j = HB / 8; j < 16; j++) w[j] = kb_msg_be64(msg, (uint64_t)(8 * j - HB), mlen);
        } else {
            const uint64_t base = blk * 128 - HB;
            KB_UNROLL
            for (int j = 0; j < 16; j++) w[j] = kb_msg_be64(msg, base + 8 * j, mlen);
        }
        if (blk == nblocks - 1) {
            w[14] = 0;  // message lengths here are < 2^61 bytes
            w[15] = total * 8;
        }
        sha512_compress(h, w);
    }
    KB_UNROLL
    for (int i = 0; i < 8; i++) {
        digest[2 * i] = kb_bswap32((uint32_t)(h[i] >> 32));
        digest[2 * i + 1] = kb_bswap32((uint32_t)h[i]);
    }
}
// SHA-512(R || A || msg): r_w / a_w are the 32-byte encodings as 8 LE words each
KB_FN void sha512_ram(uint32_t* digest, const uint32_t* r_w, const uint32_t* a_w, const uint8_t* msg, uint64_t mlen)
{
    uint32_t head[16];
    KB_UNROLL
    for (int i = 0; i < 8; i++) {
        head[i] = r_w[i];
        head[8 + i] = a_w[i];
    }
    sha512_prefixed<16>(digest, head, msg, mlen);
}
