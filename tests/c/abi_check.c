/* tests/c/abi_check.c — the drop-in boundary used from plain C (no Python, no torch): what a cgo / Rust -sys /
 * JNI caller sees.  Reads a fixture written by tests/test_gpu_parity.py::test_c_abi_from_plain_c:
 *   u64 n, u64 msg_bytes, pk[32n], sig[64n], msg_off[u64 (n+1)], msg[msg_bytes], want_status[n],
 *   scalars[32n], want_base_mul[32n]
 * and checks kb_eddsa_verify_batch, kb_schnorr_verify_batch (same verdicts for valid items) and kb_point_mul_base_batch
 * against the expected bytes (which the Python side took from the oracle).  Exit code 0 = all equal. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kyber_b200.h"

static void* slurp(FILE* f, size_t bytes)
{
    void* p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
    return p;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint64_t hdr[2];
    if (fread(hdr, 8, 2, f) != 2) return 2;
    const size_t n = (size_t)hdr[0], mb = (size_t)hdr[1];
    uint8_t* pk = slurp(f, 32 * n);
    uint8_t* sig = slurp(f, 64 * n);
    uint64_t* off = slurp(f, 8 * (n + 1));
    uint8_t* msg = slurp(f, mb);
    uint8_t* want = slurp(f, n);
    uint8_t* scalars = slurp(f, 32 * n);
    uint8_t* want_mul = slurp(f, 32 * n);
    fclose(f);

    kb_ctx* ctx = NULL;
    int rc = kb_ctx_create(0, &ctx);
    if (rc != KB_OK) {
        fprintf(stderr, "kb_ctx_create: %d (no GPU: there is no CPU fallback)\n", rc);
        return 3;
    }
    uint8_t* st = malloc(n);
    uint8_t* out = malloc(32 * n);
    int bad = 0;
    rc = kb_eddsa_verify_batch(ctx, n, pk, msg, off, sig, st);
    if (rc != KB_OK) { fprintf(stderr, "verify: %d %s\n", rc, kb_last_error(ctx)); return 4; }
    for (size_t i = 0; i < n; i++) bad += st[i] != want[i];
    rc = kb_schnorr_verify_batch(ctx, n, pk, msg, off, sig, st);
    if (rc != KB_OK) return 4;
    for (size_t i = 0; i < n; i++) bad += (st[i] == KB_SIG_STATUS_OK) != (want[i] == KB_SIG_STATUS_OK);
    rc = kb_point_mul_base_batch(ctx, n, scalars, out, 0);
    if (rc != KB_OK) return 4;
    bad += memcmp(out, want_mul, 32 * n) != 0;
    rc = kb_point_mul_base_batch(ctx, n, scalars, out, KB_FLAG_VARTIME);
    if (rc != KB_OK) return 4;
    bad += memcmp(out, want_mul, 32 * n) != 0;
    printf("abi_check: n=%zu sm=%d launches=%llu mismatches=%d\n", n, kb_device_sm_count(ctx), (unsigned long long)kb_launch_count(ctx), bad);
    kb_ctx_destroy(ctx);
    return bad ? 1 : 0;
}
