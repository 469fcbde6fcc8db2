// Host-only invariants of the fast kernel's launch geometry (no device needed): for every window
// length and builder the plan must respect the shared-memory limit, the field widths of the
// queue entries, whole-iteration segments, and the sub-window rules of long windows.
#include <cstdio>
#include <cstdlib>

#include "../simd-minimizers_b200/csrc/mz_fast.cuh"

#define CHECK(c)                                                                         \
    do {                                                                                 \
        if (!(c)) {                                                                      \
            printf("FAILED %s (w=%u mode=%u canon=%u nwin=%llu S=%u)\n", #c, p.w, p.mode, \
                   p.strand_tiebreak, (unsigned long long)nwin, pl.S);                   \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main() {
    const unsigned long long sizes[] = {1, 31, 1000, 123457, 10000000ull, 387500000ull, 3100000000ull, 4294967000ull};
    for (unsigned w = 1; w <= 300; w++) {
        for (unsigned mode = 0; mode < 3; mode++) {
            for (unsigned canon = 0; canon < 2; canon++) {
                mz_params p{};
                p.k = 21, p.w = w, p.mode = mode, p.strand_tiebreak = canon, p.hash_canonical = canon;
                for (unsigned long long nwin : sizes) {
                    mz::FastPlan pl;
                    const bool ok = mz::plan_fast(148, p, nwin, &pl, /*allow_xw=*/true);
                    if (w > mz::FAST_XW_MAX_W) {
                        CHECK(!ok);
                        continue;
                    }
                    CHECK(ok);
                    const unsigned wt = mz::fast_wt(w), sb = mz::fast_sb(wt);
                    CHECK(w <= 32 ? wt == w : (wt >= 17 && wt <= 24));
                    CHECK(sb <= 32 && sb % wt == 0);
                    if (w > 32) {  // consecutive sub-windows overlap or abut and cover the window
                        const unsigned T = (w + wt - 1) / wt - 1;
                        CHECK(T >= 1 && T * wt >= w - wt && (T + 1) * wt >= w);
                        CHECK(w - wt >= 16);  // taps are read a group of four k-mers ahead
                        const unsigned ring = mz::fast_ring_rows(w);
                        CHECK((ring & (ring - 1)) == 0 && ring >= (w - wt) + sb + 4);
                        CHECK(pl.r1_words == (size_t)ring * 32 * (canon ? 2 : 1));
                    } else {
                        CHECK(pl.r1_words == 0);
                    }
                    const unsigned lead = mz::fast_lead(w);
                    CHECK(lead >= w && lead % wt == 0 && lead < w + wt && pl.lead == lead);
                    CHECK(pl.S >= 1 && (unsigned long long)pl.S + lead + sb + 2 < 65535);
                    CHECK(pl.nb == mz::fast_nb(pl.S, w) && pl.nb * sb >= lead + pl.S);
                    // whole iterations whenever the segment length was free to choose
                    CHECK((pl.S + lead) % sb == 0 || getenv("MZ_FAST_S") || pl.S < sb);
                    // queue entries: selected k-mer in 11 bits (u16 entries), window end - k-mer in 5 / 8 bits
                    CHECK(w > 32 || pl.nb * sb + 1 < 2048);
                    CHECK(w <= 32 ? mz::fast_qbytes(w) == 2 : mz::fast_qbytes(w) == 4);
                    CHECK(pl.q_rows == pl.q_trig + (w > 32 ? 4 : sb) && pl.q_trig >= 1);  // guard rows between two overflow checks
                    CHECK((pl.q_bufs == 1 || pl.q_bufs == 2) && mz::fast_smem(w, pl.q_rows, pl.q_bufs) <= mz::FAST_SMEM_MAX);
                    CHECK((unsigned long long)pl.num_tiles * 32 * pl.S >= nwin);
                    CHECK(pl.grid >= 1 && pl.grid <= 148 * mz::FAST_BPS);
                    CHECK(pl.scratch_words_per_block == mz::fast_spill_words(pl.S, w));
                    // the spill area holds every entry a lane can push in a tile
                    CHECK(pl.scratch_words_per_block * 4 >= (size_t)pl.nb * sb * 32 * mz::fast_qbytes(w));
                }
            }
        }
    }
    printf("plan ok\n");
    return 0;
}
