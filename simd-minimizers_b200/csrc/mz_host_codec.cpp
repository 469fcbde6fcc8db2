// mz_host_codec.cpp -- host half of the transfer codec for minimizer positions / super-k-mer
// starts (the device half is mz_delta_encode_kernel in mz_api.cu): entry i of a chunk crosses
// PCIe as the signed byte v[i] - v[i-1], with an absolute u32 every 256 entries; this file adds
// the deltas up again while it writes the caller's array.  It is a decoder for bytes this
// library's own kernels produced -- it computes nothing about minimizers.
//
// 8 entries per step with AVX2 when the CPU has it (sign-extend, in-register prefix sum,
// streaming stores: the output is written once and not read again here, so it should not pull the
// destination lines into the cache first), scalar otherwise.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

namespace mz {

constexpr uint32_t kDeltaBlockHost = 256;

static void decode_blocks_scalar(const int8_t* delta, const uint32_t* base, uint64_t n, uint64_t b0, uint64_t b1,
                                 uint32_t* out) {
    for (uint64_t b = b0; b < b1; b++) {
        const uint64_t lo = b * kDeltaBlockHost, hi = lo + kDeltaBlockHost < n ? lo + kDeltaBlockHost : n;
        uint32_t v = base[b];
        out[lo] = v;
        for (uint64_t i = lo + 1; i < hi; i++) {
            v += (uint32_t)(int32_t)delta[i];
            out[i] = v;
        }
    }
}

__attribute__((target("avx2"))) static void decode_blocks_avx2(const int8_t* delta, const uint32_t* base, uint64_t n,
                                                               uint64_t b0, uint64_t b1, uint32_t* out) {
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 31u) == 0;  // block starts are multiples of 256 entries
    for (uint64_t b = b0; b < b1; b++) {
        const uint64_t lo = b * kDeltaBlockHost;
        if (lo + kDeltaBlockHost > n) {  // ragged last block
            decode_blocks_scalar(delta, base, n, b, b + 1, out);
            continue;
        }
        __m256i carry = _mm256_set1_epi32((int)base[b]);  // delta[lo] is 0 by construction
        for (uint32_t i = 0; i < kDeltaBlockHost; i += 8) {
            __m256i d = _mm256_cvtepi8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(delta + lo + i)));
            d = _mm256_add_epi32(d, _mm256_slli_si256(d, 4));   // prefix sums inside each 128-bit half
            d = _mm256_add_epi32(d, _mm256_slli_si256(d, 8));
            const __m256i lowtot = _mm256_shuffle_epi32(_mm256_permute2x128_si256(d, d, 0x08), 0xFF);
            d = _mm256_add_epi32(_mm256_add_epi32(d, lowtot), carry);  // + total of the low half, + running value
            if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(out + lo + i), d);
            else _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + lo + i), d);
            carry = _mm256_permutevar8x32_epi32(d, _mm256_set1_epi32(7));
        }
    }
    _mm_sfence();
}

// blocks [b0, b1) of an encoded array of n entries
void delta_decode_blocks(const int8_t* delta, const uint32_t* base, uint64_t n, uint64_t b0, uint64_t b1, uint32_t* out) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) decode_blocks_avx2(delta, base, n, b0, b1, out);
    else decode_blocks_scalar(delta, base, n, b0, b1, out);
}

}  // namespace mz
