/*
 * mzbaseline_avx2.c -- 8-lane AVX2 CPU baseline of the minimizer path (TEST/BENCH INFRASTRUCTURE).
 *
 * The reference crate is single-threaded 8-lane SIMD (src/lib.rs:9,30): the sequence is cut into
 * 8 chunks that advance in lock-step inside u32x8 registers (src/sliding_min.rs:222-355,
 * src/canonical.rs:42-62, src/collect.rs:128-285).  The crate itself cannot be built here (Rust),
 * so bench.py's cpu_baseline / `--impl reference` arm times THIS restatement of that design
 * ("kind": "port"): same lane split, same packed (key<<16|pos) two-stacks minimum with min/max,
 * same rolling TG count, per-lane dedup, lanes flattened in order.  Only bench.py and tests/ may
 * load it; the product never does.  Limits: minimizer mode, k-mer hash with rot=7 tables,
 * sequences < 2^32 bases; anything else falls back to the scalar oracle.
 */
#include <immintrin.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "mzoracle.h"

#if defined(__AVX2__)

/* Per-lane output region: lane i collects into its own eighth of the caller's slice (the
 * reference keeps one Vec per lane in a thread-local cache, src/collect.rs:124-126, and flattens
 * them in order afterwards, :252-272); nothing is allocated while the lanes run. */
typedef struct {
    uint32_t* p;
    uint64_t n, cap;
    int overflow;
} lanebuf;

static inline void lb_push(lanebuf* b, uint32_t v) {
    if (b->n == b->cap) {
        b->overflow = 1;
        return;
    }
    b->p[b->n++] = v;
}

static inline __m256i rotl7(__m256i x) { return _mm256_or_si256(_mm256_slli_epi32(x, 7), _mm256_srli_epi32(x, 25)); }
static inline __m256i rotr7(__m256i x) { return _mm256_or_si256(_mm256_srli_epi32(x, 7), _mm256_slli_epi32(x, 25)); }
static inline uint32_t rotl32s(uint32_t x, uint32_t r) {
    r &= 31u;
    return r ? (x << r) | (x >> (32u - r)) : x;
}

/* One stream of 2-bit bases for 8 lanes that share the same phase: a gather every 16 bases. */
typedef struct {
    __m256i widx; /* word index per lane */
    __m256i cur;
    int left;
} stream8;

static inline void s8_init(stream8* s, const uint32_t* words, __m256i wlim, const uint64_t bit[8]) {
    uint32_t wi[8];
    for (int i = 0; i < 8; i++) wi[i] = (uint32_t)(bit[i] >> 5);
    uint32_t sh = (uint32_t)(bit[0] & 31u); /* same for all lanes by construction */
    s->widx = _mm256_loadu_si256((const __m256i*)wi);
    __m256i w = _mm256_i32gather_epi32((const int*)words, _mm256_min_epu32(s->widx, wlim), 4);
    s->cur = _mm256_srl_epi32(w, _mm_cvtsi32_si128((int)sh));
    s->left = (int)((32u - sh) >> 1);
}
static inline __m256i s8_next(stream8* s, const uint32_t* words, __m256i wlim) {
    if (s->left == 0) {
        s->widx = _mm256_add_epi32(s->widx, _mm256_set1_epi32(1));
        s->cur = _mm256_i32gather_epi32((const int*)words, _mm256_min_epu32(s->widx, wlim), 4);
        s->left = 16;
    }
    __m256i b = _mm256_and_si256(s->cur, _mm256_set1_epi32(3));
    s->cur = _mm256_srli_epi32(s->cur, 2);
    s->left--;
    return b;
}

/* Windows [wb, we) of the sequence; appends to pos_out/sk_out; returns count or -1. */
static uint64_t run_range_avx2(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p,
                               uint64_t wb, uint64_t we, uint32_t* pos_out, uint32_t* sk_out,
                               uint64_t cap) {
    const uint32_t k = p->k, w = p->w, l = k + w - 1;
    const int lr = (int)p->strand_tiebreak, hc = (int)p->hasher.canonical;
    const uint64_t range = we - wb;
    /* lane length in windows, multiple of 16 so every lane has the same bit phase */
    const uint64_t Lw = ((range + 7) / 8 + 15) / 16 * 16;
    /* words view: align the byte pointer down to 4 bytes */
    const uintptr_t addr = (uintptr_t)packed;
    const uint32_t* words = (const uint32_t*)(addr & ~(uintptr_t)3);
    const uint64_t bit_base = 8 * (addr & 3) + 2 * off;
    const uint64_t nwords_avail = (bit_base + 2 * n + 31) / 32;
    const __m256i wlim = _mm256_set1_epi32((int)(nwords_avail - 1));

    /* Every lane starts at its own first window and emits it unconditionally; lane (and
     * thread) seams are stitched afterwards by dropping a first element that repeats the
     * previous last one -- the reference's flatten rule (src/collect.rs:252-272). */
    uint64_t first[8], lane_end[8], bit0[8];
    uint32_t basepos[8];
    for (int i = 0; i < 8; i++) {
        uint64_t ls = wb + (uint64_t)i * Lw;
        first[i] = ls;
        lane_end[i] = ls + Lw < we ? ls + Lw : we;
        if (ls >= we) lane_end[i] = ls; /* empty lane */
        bit0[i] = bit_base + 2 * first[i];
        basepos[i] = (uint32_t)first[i];
    }
    const __m256i vbase = _mm256_loadu_si256((const __m256i*)basepos);

    /* tables in permutevar layout */
    uint32_t tf[8], tfr[8], tc[8], tcr[8];
    for (int b = 0; b < 8; b++) {
        tf[b] = p->hasher.f[b & 3];
        tfr[b] = rotl32s(p->hasher.f[b & 3], 7u * (k - 1));
        tc[b] = p->hasher.c[b & 3];
        tcr[b] = rotl32s(p->hasher.c[b & 3], 7u * (k - 1));
    }
    const __m256i TF = _mm256_loadu_si256((const __m256i*)tf), TFR = _mm256_loadu_si256((const __m256i*)tfr);
    const __m256i TC = _mm256_loadu_si256((const __m256i*)tc), TCR = _mm256_loadu_si256((const __m256i*)tcr);

    stream8 sin, sout, stg;
    s8_init(&sin, words, wlim, bit0);
    s8_init(&sout, words, wlim, bit0);
    s8_init(&stg, words, wlim, bit0);

    __m256i* ringl = (__m256i*)aligned_alloc(32, sizeof(__m256i) * w * 2);
    __m256i* ringr = ringl + w;
    for (uint32_t i = 0; i < w; i++) ringl[i] = _mm256_set1_epi32(-1), ringr[i] = _mm256_setzero_si256();
    __m256i prel = _mm256_set1_epi32(-1), prer = _mm256_setzero_si256();
    uint32_t ridx = 0;
    __m256i fw = _mm256_setzero_si256(), rc = _mm256_setzero_si256();
    __m256i tg = _mm256_set1_epi32(-(int)l);
    const __m256i vmask = _mm256_set1_epi32((int)0xffff0000u), two = _mm256_set1_epi32(2);
    __m256i prev = _mm256_set1_epi32(-1);
    lanebuf lb[8], lsk[8];
    memset(lb, 0, sizeof lb);
    memset(lsk, 0, sizeof lsk);
    const uint64_t lcap = cap / 8;
    for (int i = 0; i < 8; i++) {
        lb[i].p = pos_out + (uint64_t)i * lcap, lb[i].cap = lcap;
        if (sk_out) lsk[i].p = sk_out + (uint64_t)i * lcap, lsk[i].cap = lcap;
    }

    const uint64_t nsteps = Lw + l - 1; /* bases per lane */
    /* 16-bit positions inside the packed elements are re-based like the reference does
     * (src/sliding_min.rs:117-125); here simply by keeping a 32-bit lane-local counter and
     * restarting the two stacks is avoided: lanes are cut so that Lw + l < 2^16 per pass. */
    uint64_t done = 0;
    (void)done;
    uint32_t e = 0; /* lane-local k-mer index */
    __m256i posoff = vbase;
    uint32_t eoff = 0; /* subtracted from e inside the packed element */
    for (uint64_t t = 0; t < nsteps; t++) {
        const __m256i a = s8_next(&sin, words, wlim);
        fw = _mm256_xor_si256(rotl7(fw), _mm256_permutevar8x32_epi32(TF, a));
        if (hc) rc = _mm256_xor_si256(rotr7(rc), _mm256_permutevar8x32_epi32(TCR, a));
        tg = _mm256_add_epi32(tg, _mm256_and_si256(a, two));
        const __m256i strand = _mm256_cmpgt_epi32(tg, _mm256_setzero_si256());
        if (t + 1 >= l) tg = _mm256_sub_epi32(tg, _mm256_and_si256(s8_next(&stg, words, wlim), two));
        if (t + 1 < k) continue;
        const __m256i h = hc ? _mm256_add_epi32(fw, rc) : fw;
        const __m256i o = s8_next(&sout, words, wlim);
        fw = _mm256_xor_si256(fw, _mm256_permutevar8x32_epi32(TFR, o));
        if (hc) rc = _mm256_xor_si256(rc, _mm256_permutevar8x32_epi32(TC, o));

        if (e - eoff == 0xffffu) { /* re-base positions (src/sliding_min.rs:117-125) */
            const uint32_t delta = 0xfffeu - w;
            const __m256i vd = _mm256_set1_epi32((int)delta);
            eoff += delta;
            prel = _mm256_sub_epi32(prel, vd);
            prer = _mm256_sub_epi32(prer, vd);
            posoff = _mm256_add_epi32(posoff, vd);
            for (uint32_t q = 0; q < w; q++) {
                ringl[q] = _mm256_sub_epi32(ringl[q], vd);
                ringr[q] = _mm256_sub_epi32(ringr[q], vd);
            }
        }
        const __m256i vpos = _mm256_set1_epi32((int)(e - eoff));
        const __m256i el = _mm256_or_si256(_mm256_and_si256(h, vmask), vpos);
        const __m256i er = _mm256_or_si256(_mm256_andnot_si256(h, vmask), vpos);
        e++;
        ringl[ridx] = el;
        prel = _mm256_min_epu32(prel, el);
        if (lr) {
            ringr[ridx] = er;
            prer = _mm256_max_epu32(prer, er);
        }
        if (++ridx == w) {
            ridx = 0;
            __m256i sl = ringl[w - 1], sr = ringr[w - 1];
            for (uint32_t q = w - 1; q-- > 0;) {
                sl = _mm256_min_epu32(sl, ringl[q]);
                ringl[q] = sl;
                if (lr) {
                    sr = _mm256_max_epu32(sr, ringr[q]);
                    ringr[q] = sr;
                }
            }
            prel = el;
            prer = er;
        }
        if (t + 1 < l) continue;
        __m256i sel = _mm256_min_epu32(prel, ringl[ridx]);
        if (lr) sel = _mm256_blendv_epi8(_mm256_max_epu32(prer, ringr[ridx]), sel, strand);
        const __m256i posv = _mm256_add_epi32(_mm256_and_si256(sel, _mm256_set1_epi32(0xffff)), posoff);
        const uint64_t jrel = t + 1 - l; /* window index relative to first[] */
        int m = ~_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(posv, prev))) & 0xff;
        if (jrel == 0) m = 0xff; /* a lane's first window always emits */
        prev = posv;
        if (m) {
            uint32_t pv[8];
            _mm256_storeu_si256((__m256i*)pv, posv);
            while (m) {
                const int lane = __builtin_ctz((unsigned)m);
                m &= m - 1;
                const uint64_t j = first[lane] + jrel;
                if (j < lane_end[lane]) {
                    lb_push(&lb[lane], pv[lane]);
                    if (sk_out) lb_push(&lsk[lane], (uint32_t)j);
                }
            }
        }
    }
    free(ringl);
    /* flatten the lanes in order (in place: lane i never starts left of what is already flat) */
    uint64_t total = 0;
    for (int i = 0; i < 8; i++) {
        if (lb[i].overflow) return (uint64_t)-1;
        uint64_t skip = (total > 0 && lb[i].n > 0 && lb[i].p[0] == pos_out[total - 1]) ? 1 : 0;
        uint64_t cntl = lb[i].n - skip;
        if (cntl) memmove(pos_out + total, lb[i].p + skip, cntl * 4);
        if (sk_out && cntl) memmove(sk_out + total, lsk[i].p + skip, cntl * 4);
        total += cntl;
    }
    return total;
}
#endif /* __AVX2__ */

int mzb_have_avx2(void) {
#if defined(__AVX2__)
    return __builtin_cpu_supports("avx2");
#else
    return 0;
#endif
}

/* k-mer values with word arithmetic (no per-base loop). len <= 32. */
static void values_u64_fast(const uint8_t* packed, uint64_t off, uint32_t len, int canonical,
                            const uint32_t* pos, uint64_t m, uint64_t* out) {
    const uint64_t mask = len < 32 ? ((1ull << (2 * len)) - 1ull) : ~0ull;
    for (uint64_t e = 0; e < m; e++) {
        const uint64_t bit = 2 * (off + pos[e]);
        const uint8_t* p = packed + (bit >> 3);
        uint64_t lo, hi = 0;
        memcpy(&lo, p, 8);
        const uint32_t sh = (uint32_t)(bit & 7u);
        uint64_t v = lo >> sh;
        if (sh) {
            hi = p[8];
            v |= hi << (64 - sh);
        }
        v &= mask;
        if (canonical) {
            uint64_t r = v;
            /* reverse the 2-bit groups */
            r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
            r = ((r >> 4) & 0x0f0f0f0f0f0f0f0full) | ((r & 0x0f0f0f0f0f0f0f0full) << 4);
            r = __builtin_bswap64(r);
            r = (r ^ 0xAAAAAAAAAAAAAAAAull) >> (64 - 2 * len);
            if (r < v) v = r;
        }
        out[e] = v;
    }
}

typedef struct {
    const uint8_t* packed;
    uint64_t off, n;
    const mzo_params* p;
    uint64_t wb, we, cap, count;
    uint32_t *pos, *sk;
    uint64_t* val;
    int use_avx2;
    uint64_t skip, dst, ncopy;
    uint32_t *pos_out, *sk_out;
    uint64_t* val_out;
} bjob;

static void* bcopy_worker(void* arg) {
    bjob* j = (bjob*)arg;
    if (!j->ncopy) return NULL;
    memcpy(j->pos_out + j->dst, j->pos + j->skip, 4 * j->ncopy);
    if (j->sk_out) memcpy(j->sk_out + j->dst, j->sk + j->skip, 4 * j->ncopy);
    if (j->val_out) memcpy(j->val_out + j->dst, j->val + j->skip, 8 * j->ncopy);
    return NULL;
}

static void* bworker(void* arg) {
    bjob* j = (bjob*)arg;
#if defined(__AVX2__)
    if (j->use_avx2)
        j->count = run_range_avx2(j->packed, j->off, j->n, j->p, j->wb, j->we, j->pos, j->sk, j->cap);
    else
#endif
        j->count = mzo_run_range(j->packed, j->off, j->n, j->p, j->wb, j->we, j->pos, j->sk, j->cap);
    if (j->count != (uint64_t)-1 && j->val)
        values_u64_fast(j->packed, j->off, j->p->k, (int)j->p->strand_tiebreak, j->pos, j->count, j->val);
    return NULL;
}

/* Timed CPU arm: like mzb_run_mt, but every thread writes straight into its own slice of the
 * caller's pre-allocated (and pre-faulted) arrays -- slice t = entries [t * cap / threads, ...) --
 * and nothing is concatenated afterwards, which is what the reference's own multi-threaded
 * benchmark does (thread-local output Vec per rayon worker, bench/src/bin/paper.rs:439-461).
 * The logical output is the concatenation of [starts_out[t], starts_out[t] + counts_out[t]) for
 * t = 0 .. threads-1 (a thread's first entry is dropped when it repeats the previous thread's
 * last one: the flatten rule).  No allocation of output memory inside.  Returns the total count. */
uint64_t mzb_run_mt_slices(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p, int threads,
                           uint32_t* pos, uint32_t* sk, uint64_t* val, uint64_t cap, uint64_t* starts_out,
                           uint64_t* counts_out) {
    const uint32_t l = p->k + p->w - 1;
    if (p->k == 0 || p->w == 0 || p->w >= (1u << 15) || n >= (1ull << 32)) return (uint64_t)-1;
    if (p->strand_tiebreak && ((l & 1u) == 0 || !p->hasher.canonical)) return (uint64_t)-1;
    if (threads < 1) threads = 1;
    for (int t = 0; t < threads; t++) starts_out[t] = counts_out[t] = 0;
    if (n < l) return 0;
    const uint64_t nwin = n - l + 1;
    int use_avx2 = mzb_have_avx2() && p->mode == MZO_MINIMIZER && p->hasher.rot == 7;
    if (p->mode != MZO_MINIMIZER || p->k > 32) val = NULL;
    int nt = threads;
    if ((uint64_t)nt * 256 > nwin) nt = (int)(nwin / 256 ? nwin / 256 : 1);
    bjob* jobs = (bjob*)calloc((size_t)nt, sizeof(bjob));
    pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    const uint64_t per = (nwin + (uint64_t)nt - 1) / (uint64_t)nt, slice = cap / (uint64_t)nt;
    for (int t = 0; t < nt; t++) {
        bjob* j = &jobs[t];
        j->packed = packed, j->off = off, j->n = n, j->p = p, j->use_avx2 = use_avx2;
        j->wb = per * (uint64_t)t;
        j->we = j->wb + per < nwin ? j->wb + per : nwin;
        if (j->wb > j->we) j->wb = j->we;
        j->cap = slice;
        j->pos = pos + slice * (uint64_t)t;
        j->sk = sk ? sk + slice * (uint64_t)t : NULL;
        j->val = val ? val + slice * (uint64_t)t : NULL;
    }
    if (nt == 1) bworker(&jobs[0]);
    else {
        for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, bworker, &jobs[t]);
        for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    }
    uint64_t m = 0;
    uint32_t last = 0;
    int have_last = 0, ok = 1;
    for (int t = 0; t < nt; t++) {
        bjob* j = &jobs[t];
        if (j->wb == j->we) continue;
        if (j->count == (uint64_t)-1) {
            ok = 0;
            break;
        }
        const uint64_t skip = (use_avx2 && have_last && j->count > 0 && j->pos[0] == last) ? 1 : 0;
        starts_out[t] = slice * (uint64_t)t + skip;
        counts_out[t] = j->count - skip;
        if (j->count) last = j->pos[j->count - 1], have_last = 1;
        m += counts_out[t];
    }
    free(jobs);
    free(th);
    return ok ? m : (uint64_t)-1;
}

/* Multi-threaded baseline: contiguous window ranges per thread (analogue of the reference
 * benchmark's rayon loop, bench/src/bin/paper.rs:442-459), 8 AVX2 lanes inside each thread.
 * The packed buffer must be readable 16 bytes past its last base. Returns count or -1. */
uint64_t mzb_run_mt(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p, int threads,
                    uint32_t* pos_out, uint32_t* sk_out, uint64_t* val_out, uint64_t cap) {
    const uint32_t l = p->k + p->w - 1;
    if (p->k == 0 || p->w == 0 || p->w >= (1u << 15) || n >= (1ull << 32)) return (uint64_t)-1;
    if (p->strand_tiebreak && ((l & 1u) == 0 || !p->hasher.canonical)) return (uint64_t)-1;
    if (n < l) return 0;
    const uint64_t nwin = n - l + 1;
    int use_avx2 = mzb_have_avx2() && p->mode == MZO_MINIMIZER && p->hasher.rot == 7;
    if (p->mode != MZO_MINIMIZER || p->k > 32) val_out = NULL;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads * 256 > nwin) threads = (int)(nwin / 256 ? nwin / 256 : 1);
    bjob* jobs = (bjob*)calloc((size_t)threads, sizeof(bjob));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    const uint64_t per = (nwin + (uint64_t)threads - 1) / (uint64_t)threads;
    int ok = 1;
    for (int t = 0; t < threads; t++) {
        bjob* j = &jobs[t];
        j->packed = packed, j->off = off, j->n = n, j->p = p, j->use_avx2 = use_avx2;
        j->wb = per * (uint64_t)t;
        j->we = j->wb + per < nwin ? j->wb + per : nwin;
        if (j->wb > j->we) j->wb = j->we;
        const uint64_t range = j->we - j->wb;
        const uint64_t est = range / (p->w + 1) * 3 + 4096;
        /* worst case (every window emits): 8 lanes of ceil16(range / 8) windows each */
        j->cap = (range < (1u << 20) || est > range) ? 8 * (((range + 7) / 8 + 15) / 16 * 16) : est;
        j->pos = (uint32_t*)malloc(4 * (j->cap + 1));
        j->sk = sk_out ? (uint32_t*)malloc(4 * (j->cap + 1)) : NULL;
        j->val = val_out ? (uint64_t*)malloc(8 * (j->cap + 1)) : NULL;
        if (!j->pos || (sk_out && !j->sk) || (val_out && !j->val)) ok = 0;
    }
    if (ok) {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, bworker, &jobs[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    /* ordered concatenation, in parallel: offsets first (seam rule), then one copy per thread */
    uint64_t m = 0;
    uint32_t last = 0;
    int have_last = 0;
    for (int t = 0; t < threads && ok; t++) {
        bjob* j = &jobs[t];
        j->skip = 0, j->dst = m, j->ncopy = 0;
        if (j->wb == j->we) continue;
        if (j->count == (uint64_t)-1) {
            ok = 0;
            break;
        }
        /* thread seam: same flatten rule (only the AVX2 path emits its first window blindly) */
        j->skip = (use_avx2 && have_last && j->count > 0 && j->pos[0] == last) ? 1 : 0;
        j->ncopy = j->count - j->skip;
        if (m + j->ncopy > cap) {
            ok = 0;
            break;
        }
        if (j->count) last = j->pos[j->count - 1], have_last = 1;
        j->pos_out = pos_out, j->sk_out = sk_out, j->val_out = val_out;
        m += j->ncopy;
    }
    if (ok) {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, bcopy_worker, &jobs[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    for (int t = 0; t < threads; t++) free(jobs[t].pos), free(jobs[t].sk), free(jobs[t].val);
    free(jobs);
    free(th);
    return ok ? m : (uint64_t)-1;
}
