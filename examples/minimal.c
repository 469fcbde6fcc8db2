/* Minimal C caller of libmzb200: canonical minimizer positions + k-mer values of the reference's
 * README example (README.md:41-50 of rust-seq/simd-minimizers).
 *   gcc -std=c99 -I include examples/minimal.c -L simd-minimizers_b200 -lmzb200 \
 *       -Wl,-rpath,$PWD/simd-minimizers_b200 -o minimal && ./minimal      (needs a CUDA device) */
#include <stdio.h>
#include <string.h>

#include "mz_b200.h"

int main(void) {
    const char* seq = "ACGTGCTCAGAGACTCAGAGGA";
    const uint64_t n = strlen(seq);
    uint8_t packed[16] = {0};
    for (uint64_t i = 0; i < n; i++) packed[i >> 2] |= (uint8_t)((((uint8_t)seq[i] >> 1) & 3u) << (2 * (i & 3)));

    mz_ctx* ctx = NULL;
    int rc = mz_ctx_create(NULL, 0, &ctx);
    if (rc != MZ_OK) {
        fprintf(stderr, "mz_ctx_create: %s (%s)\n", mz_strerror(rc), mz_last_error());
        return 1;
    }
    mz_params p;
    mz_params_nthash(&p, /*k=*/5, /*w=*/7, MZ_MODE_MINIMIZER, /*canonical=*/1);
    p.value_bits = 64;
    uint32_t pos[32];
    uint64_t val[32];
    mz_out out = {pos, NULL, val, 32, 0};
    rc = mz_run(ctx, &p, packed, 0, n, &out);
    if (rc != MZ_OK) {
        fprintf(stderr, "mz_run: %s (%s)\n", mz_strerror(rc), mz_last_error());
        return 1;
    }
    for (uint64_t i = 0; i < out.count; i++) printf("pos %u value %llu\n", pos[i], (unsigned long long)val[i]);
    /* expected: positions 0 7 9 15, values 721 817 307 817 */
    mz_ctx_destroy(ctx);
    return out.count == 4 && pos[0] == 0 && pos[1] == 7 && pos[2] == 9 && pos[3] == 15 ? 0 : 2;
}
