// Readme / doc-test vectors of the reference through the C++ mirror (needs a GPU to run).
#include <cassert>
#include <cstdio>
#include <cstring>

#include "simd_minimizers.hpp"
using namespace simd_minimizers;

static std::vector<uint8_t> pack(const char* s) {
    size_t n = strlen(s);
    std::vector<uint8_t> out((n + 3) / 4 + 16, 0);
    for (size_t i = 0; i < n; i++) out[i >> 2] |= (uint8_t)((((uint8_t)s[i] >> 1) & 3) << (2 * (i & 3)));
    return out;
}

int main() {
    {   // src/lib.rs:92-99
        const char* s = "ACGTGCTCAGAGACTCAG";
        auto d = pack(s);
        auto pos = minimizer_positions({d.data(), 0, strlen(s)}, 5, 7);
        assert((pos == std::vector<uint32_t>{4, 5, 8, 13}));
    }
    {   // src/lib.rs:109-129, src/test.rs:402-426
        const char* s = "ACGTGCTCAGAGACTCAGAGGA";
        auto d = pack(s);
        PackedSeq seq{d.data(), 0, strlen(s)};
        assert((canonical_minimizer_positions(seq, 5, 7) == std::vector<uint32_t>{0, 7, 9, 15}));
        std::vector<uint32_t> pos, sk;
        auto out = canonical_minimizers(5, 7).hasher(Hasher::nt(5)).super_kmers(sk).run(seq, pos);
        assert((pos == std::vector<uint32_t>{0, 7, 9, 15}));
        assert((out.values_u64() == std::vector<uint64_t>{0b1011010001, 0b1100110001, 0b0100110011, 0b1100110001}));
        assert(sk.size() == pos.size() && sk[0] == 0);
        bool threw = false;
        try { canonical_minimizer_positions(seq, 4, 3); } catch (const std::invalid_argument&) { threw = true; }
        assert(threw);  // even l, src/canonical.rs:13-16
    }
    {   // src/lib.rs:451-496 / src/test.rs:429-482: no reported k-mer holds an ambiguous base, and
        // a sequence without ambiguous bases gives the plain result
        const char* s = "ACGTGCTCAGAGANTCAGAGGATTACAGATTACANNNNGATTACAGGATCCA";
        const size_t n = strlen(s);
        auto d = pack(s);
        std::vector<uint8_t> amb((n + 7) / 8 + 16, 0), none((n + 7) / 8 + 16, 0);
        for (size_t i = 0; i < n; i++)
            if (s[i] == 'N') amb[i >> 3] |= (uint8_t)(1u << (i & 7));
        PackedNSeq nseq{{d.data(), 0, n}, amb.data(), 0};
        auto pos = canonical_minimizers(5, 3).run_skip_ambiguous_windows_once(nseq);
        assert(!pos.empty());
        for (uint32_t p : pos)
            for (uint32_t j = 0; j < 5; j++) assert(s[p + j] != 'N');
        PackedNSeq clean{{d.data(), 0, n}, none.data(), 0};
        assert(canonical_minimizers(5, 3).run_skip_ambiguous_windows_once(clean) == canonical_minimizer_positions({d.data(), 0, n}, 5, 3));
        bool threw = false;
        try { minimizers(5, 3).run_skip_ambiguous_windows_once(nseq); } catch (const std::logic_error&) { threw = true; }
        assert(threw);
    }
    printf("cpp mirror ok\n");
    return 0;
}
