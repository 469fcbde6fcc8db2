/*
 * mzoracle.c -- CPU ORACLE (test infrastructure only; see mzoracle.h header comment).
 *
 * Restates, in plain C, the reference's random-minimizer path:
 *   hash        seq-hash 0.2.0 NtHasher/MulHasher (NOT in /root/reference; call sites
 *               src/minimizers.rs:24,44,61,85,143) -- published ntHash32 variant:
 *               per-base table, rotate by 7 bits per base, canonical = fw + rc.
 *   window min  src/sliding_min.rs:86-212 (top 16 bits are the key; leftmost / rightmost)
 *   strand      src/canonical.rs:12-31   (2 * #TG > l)
 *   compose     src/minimizers.rs:22-28 (naive), :38-49, :74-129 (streaming)
 *   dedup / sk  src/collect.rs:15-76
 *   syncmers    src/syncmers.rs:19-48
 *   values      src/lib.rs:598-629
 */
#include "mzoracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t rotl32(uint32_t x, uint32_t r) {
    r &= 31u;
    return r ? (x << r) | (x >> (32u - r)) : x;
}
static inline uint32_t rotr32(uint32_t x, uint32_t r) { return rotl32(x, (32u - (r & 31u)) & 31u); }

/* ------------------------------------------------------------------------------------------
 * Hashers.  seq-hash 0.2.0 is not vendored; these are restatements.
 * NtHasher: low 32 bits of the classic ntHash seeds (same constants appear in the reference at
 * bench/src/nthash.rs:26-29) listed in A,C,G,T order but INDEXED BY PACKED CODE (A,C,T,G), so
 * code 2 (T) carries the "G" seed and code 3 (G) the "T" seed.  Verified against the golden
 * vectors (see header).  Complement of a packed code is code ^ 2.
 * ---------------------------------------------------------------------------------------- */
void mzo_hasher_nt(mzo_hasher* h, int canonical) {
    static const uint32_t F[4] = {0x95c60474u, 0x62a02b4cu, 0x82572324u, 0x4be24456u};
    for (int b = 0; b < 4; b++) {
        h->f[b] = F[b];
        h->c[b] = F[b ^ 2];
    }
    h->rot = 7;
    h->canonical = canonical ? 1 : 0;
}

/* MulHasher: PARITY UNPINNED (recollection of seq-hash): f(b) = b * C, C = low 32 bits of the
 * FxHash constant 0x517cc1b727220a95 (cf. bench/src/rescan_daniel.rs:38), c(b) = f(b ^ 2). */
void mzo_hasher_mul(mzo_hasher* h, int canonical) {
    const uint32_t C = 0x27220a95u;
    for (uint32_t b = 0; b < 4; b++) {
        h->f[b] = b * C;
        h->c[b] = (b ^ 2u) * C;
    }
    h->rot = 7;
    h->canonical = canonical ? 1 : 0;
}

uint64_t mzo_pack_ascii(const char* ascii, uint64_t n, uint8_t* out) {
    uint64_t nbytes = (n + 3) / 4;
    memset(out, 0, nbytes);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t code = ((uint8_t)ascii[i] >> 1) & 3u; /* A=0 C=1 T=2 G=3, src/lib.rs:120-123 */
        out[i >> 2] |= (uint8_t)(code << (2 * (i & 3)));
    }
    return nbytes;
}

void mzo_revcomp(const uint8_t* packed, uint64_t off, uint64_t n, uint8_t* out) {
    memset(out, 0, (n + 3) / 4);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t code = mzo_base(packed, off, n - 1 - i) ^ 2u;
        out[i >> 2] |= (uint8_t)(code << (2 * (i & 3)));
    }
}

uint32_t mzo_hash_kmer(const uint8_t* packed, uint64_t off, uint64_t i, uint32_t k,
                       const mzo_hasher* h) {
    uint32_t fw = 0, rc = 0;
    for (uint32_t j = 0; j < k; j++) {
        uint32_t b = mzo_base(packed, off, i + j);
        fw ^= rotl32(h->f[b], (h->rot * (k - 1 - j)) & 31u);
        rc ^= rotl32(h->c[b], (h->rot * j) & 31u);
    }
    return h->canonical ? fw + rc : fw;
}

static int params_ok(const mzo_params* p) {
    if (p->k == 0 || p->w == 0) return 0;
    if (p->w >= (1u << 15)) return 0;                              /* src/sliding_min.rs:92-95 */
    if (p->strand_tiebreak && ((p->k + p->w - 1) & 1u) == 0) return 0; /* src/canonical.rs:13-16 */
    if (p->strand_tiebreak && !p->hasher.canonical) return 0;      /* src/minimizers.rs:81,139 */
    if (p->mode == MZO_OPEN_SYNCMER && (p->w & 1u) == 0) return 0;  /* src/syncmers.rs:24-29 */
    if (p->mode > 2) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Naive: every window on its own.  one_minimizer (src/minimizers.rs:22-28) takes the position
 * of the minimum of (hash & 0xffff0000); the canonical rule is the scalar pipeline's
 * `if canonical { left } else { right }` (src/minimizers.rs:120-126).
 * ---------------------------------------------------------------------------------------- */
uint64_t mzo_window_positions_naive(const uint8_t* packed, uint64_t off, uint64_t n,
                                    const mzo_params* p, uint32_t* out) {
    const uint32_t k = p->k, w = p->w, l = k + w - 1;
    if (n < l) return 0;
    uint64_t nwin = n - l + 1;
    for (uint64_t j = 0; j < nwin; j++) {
        uint32_t best = 0xffffffffu;
        uint64_t left = 0, right = 0;
        for (uint32_t t = 0; t < w; t++) {
            uint32_t key = mzo_hash_kmer(packed, off, j + t, k, &p->hasher) >> 16;
            if (t == 0 || key < best) {
                best = key;
                left = right = j + t;
            } else if (key == best) {
                right = j + t;
            }
        }
        uint64_t sel = left;
        if (p->strand_tiebreak) {
            uint32_t tg = 0;
            for (uint32_t t = 0; t < l; t++) tg += (mzo_base(packed, off, j + t) >> 1) & 1u;
            sel = (2 * tg > l) ? left : right;
        }
        out[j] = (uint32_t)sel;
    }
    return nwin;
}

/* ------------------------------------------------------------------------------------------
 * Streaming state machine shared by the per-window and fused variants.
 * Two-stacks sliding minimum as in src/sliding_min.rs:145-212, but with 64-bit
 * (key << 32 | pos) elements, which removes the reference's 16-bit position re-basing
 * (:179-189) without changing any result.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* packed;
    uint64_t off;
    uint32_t k, w, l;
    const mzo_hasher* h;
    int lr;
    uint32_t frot[4]; /* f rotated by rot*(k-1): leaving base, fw side */
    uint32_t crot[4]; /* c rotated by rot*(k-1): entering base, rc side */
    uint32_t fw, rc;
    int64_t tg;       /* 2*#TG - l over the current window once warm (src/canonical.rs:19) */
    uint64_t *ringl, *ringr;
    uint32_t ridx;
    uint64_t prel, prer;
} stream_t;

static int stream_init(stream_t* s, const uint8_t* packed, uint64_t off, const mzo_params* p) {
    memset(s, 0, sizeof *s);
    s->packed = packed;
    s->off = off;
    s->k = p->k;
    s->w = p->w;
    s->l = p->k + p->w - 1;
    s->h = &p->hasher;
    s->lr = (int)p->strand_tiebreak;
    uint32_t r = (p->hasher.rot * (p->k - 1)) & 31u;
    for (int b = 0; b < 4; b++) {
        s->frot[b] = rotl32(p->hasher.f[b], r);
        s->crot[b] = rotl32(p->hasher.c[b], r);
    }
    s->tg = -(int64_t)s->l;
    s->ringl = (uint64_t*)malloc(sizeof(uint64_t) * p->w * 2);
    if (!s->ringl) return -1;
    s->ringr = s->ringl + p->w;
    for (uint32_t i = 0; i < p->w; i++) s->ringl[i] = ~0ull, s->ringr[i] = 0;
    s->prel = ~0ull;
    s->prer = 0;
    return 0;
}
static void stream_free(stream_t* s) { free(s->ringl); }

/* Feed base number t (0-based from the stream start `base0`).  Returns 1 and sets *sel when a
 * full window (ending at this base) is available. */
static inline int stream_step(stream_t* s, uint64_t base0, uint64_t t, uint64_t* sel) {
    const uint32_t k = s->k, w = s->w, l = s->l, R = s->h->rot;
    uint32_t a = mzo_base(s->packed, s->off, base0 + t);
    s->fw = rotl32(s->fw, R) ^ s->h->f[a];
    s->rc = rotr32(s->rc, R) ^ s->crot[a];
    s->tg += 2 * (int64_t)((a >> 1) & 1u);
    int strand = s->tg > 0;
    if (t + 1 >= l) s->tg -= 2 * (int64_t)((mzo_base(s->packed, s->off, base0 + t + 1 - l) >> 1) & 1u);
    if (t + 1 < k) return 0;
    uint64_t i = base0 + t + 1 - k; /* k-mer start */
    uint32_t hash = s->h->canonical ? s->fw + s->rc : s->fw;
    uint32_t out = mzo_base(s->packed, s->off, i);
    s->fw ^= s->frot[out];
    s->rc ^= s->h->c[out];

    uint64_t key = hash >> 16;
    uint64_t el = (key << 32) | i;             /* min  -> smallest key, leftmost  */
    uint64_t er = ((0xffffull - key) << 32) | i; /* max -> smallest key, rightmost */
    s->ringl[s->ridx] = el;
    s->ringr[s->ridx] = er;
    if (el < s->prel) s->prel = el;
    if (er > s->prer) s->prer = er;
    if (++s->ridx == w) {
        s->ridx = 0;
        uint64_t sl = s->ringl[w - 1], sr = s->ringr[w - 1];
        for (uint32_t q = w - 1; q-- > 0;) {
            if (s->ringl[q] < sl) sl = s->ringl[q];
            if (s->ringr[q] > sr) sr = s->ringr[q];
            s->ringl[q] = sl;
            s->ringr[q] = sr;
        }
        s->prel = el;
        s->prer = er;
    }
    uint64_t ml = s->prel < s->ringl[s->ridx] ? s->prel : s->ringl[s->ridx];
    uint64_t mr = s->prer > s->ringr[s->ridx] ? s->prer : s->ringr[s->ridx];
    if (t + 1 < l) return 0;
    uint64_t left = ml & 0xffffffffull, right = mr & 0xffffffffull;
    *sel = (s->lr && !strand) ? right : left;
    return 1;
}

uint64_t mzo_window_positions_stream(const uint8_t* packed, uint64_t off, uint64_t n,
                                     const mzo_params* p, uint32_t* out) {
    const uint32_t l = p->k + p->w - 1;
    if (n < l) return 0;
    stream_t s;
    if (stream_init(&s, packed, off, p)) return (uint64_t)-1;
    uint64_t nw = 0, sel;
    for (uint64_t t = 0; t < n; t++)
        if (stream_step(&s, 0, t, &sel)) out[nw++] = (uint32_t)sel;
    stream_free(&s);
    return nw;
}

/* src/collect.rs:15-76: keep a value when it differs from its predecessor; the super-k-mer
 * index is the window at which it first appears (test vector src/test.rs:344-356). */
uint64_t mzo_collect_dedup(const uint32_t* win_pos, uint64_t nwin, uint32_t* pos_out,
                           uint32_t* sk_out) {
    uint64_t m = 0;
    for (uint64_t j = 0; j < nwin; j++) {
        if (j == 0 || win_pos[j] != win_pos[j - 1]) {
            pos_out[m] = win_pos[j];
            if (sk_out) sk_out[m] = (uint32_t)j;
            m++;
        }
    }
    return m;
}

/* src/syncmers.rs:19-48 */
uint64_t mzo_collect_syncmers(const uint32_t* win_pos, uint64_t nwin, uint32_t w, int open,
                              uint32_t* out) {
    uint64_t m = 0;
    for (uint64_t j = 0; j < nwin; j++) {
        uint64_t mp = win_pos[j];
        int is = open ? (mp == j + w / 2) : (mp == j || mp == j + w - 1);
        if (is) out[m++] = (uint32_t)j;
    }
    return m;
}

uint64_t mzo_run(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p, int algo,
                 uint32_t* pos_out, uint32_t* sk_out) {
    if (!params_ok(p)) return (uint64_t)-1;
    if (n >= (1ull << 32)) return (uint64_t)-1; /* src/sliding_min.rs:96-99 */
    const uint32_t l = p->k + p->w - 1;
    if (n < l) return 0;
    uint64_t nwin = n - l + 1;
    uint32_t* wp = (uint32_t*)malloc(sizeof(uint32_t) * nwin);
    if (!wp) return (uint64_t)-1;
    uint64_t got = algo == 0 ? mzo_window_positions_naive(packed, off, n, p, wp)
                             : mzo_window_positions_stream(packed, off, n, p, wp);
    uint64_t m;
    if (got != nwin) {
        m = (uint64_t)-1;
    } else if (p->mode == MZO_MINIMIZER) {
        m = mzo_collect_dedup(wp, nwin, pos_out, sk_out);
    } else {
        m = mzo_collect_syncmers(wp, nwin, p->w, p->mode == MZO_OPEN_SYNCMER, pos_out);
    }
    free(wp);
    return m;
}

/* ---- ambiguous bases (src/lib.rs:451-496, src/minimizers.rs:169-214) ------------------------
 * PackedNSeq = packed 2-bit codes + one ambiguity bit per base.  A window that contains an
 * ambiguous base yields SKIPPED (u32::MAX - 1, src/minimizers.rs:18); the collector then drops
 * every SKIPPED element and every element equal to the element just before it in the stream
 * (src/intrinsics/dedup.rs:147-155: the comparison is against the immediately preceding stream
 * element, skipped or not; test vectors src/test.rs:359-399).
 * PARITY NOTE: the bit layout of packed-seq 5.0.0's BitSeq is not in the reference tree; the
 * mask here is one bit per base, base i -> bit (i & 7) of byte i >> 3 (LSB first). */
#define MZO_SKIPPED 0xfffffffeu

static inline uint32_t amb_bit(const uint8_t* amb, uint64_t off, uint64_t i) {
    uint64_t p = off + i;
    return (amb[p >> 3] >> (p & 7)) & 1u;
}

uint64_t mzo_pack_ascii_n(const char* ascii, uint64_t n, uint8_t* packed_out, uint8_t* amb_out) {
    memset(amb_out, 0, (n + 7) / 8);
    for (uint64_t i = 0; i < n; i++) {
        uint8_t u = (uint8_t)ascii[i] & 0xDFu; /* upper case */
        if (!(u == 'A' || u == 'C' || u == 'G' || u == 'T')) amb_out[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
    return mzo_pack_ascii(ascii, n, packed_out);
}

uint64_t mzo_collect_dedup_skip_max(const uint32_t* win_pos, uint64_t nwin, uint32_t* pos_out) {
    uint64_t m = 0;
    for (uint64_t j = 0; j < nwin; j++) {
        uint32_t x = win_pos[j];
        if (x == MZO_SKIPPED) continue;
        if (j == 0 || x != win_pos[j - 1]) pos_out[m++] = x;
    }
    return m;
}

uint64_t mzo_run_skip_ambiguous(const uint8_t* packed, uint64_t off, uint64_t n, const uint8_t* amb,
                                uint64_t amb_off, const mzo_params* p, int algo, uint32_t* pos_out) {
    if (!params_ok(p)) return (uint64_t)-1;
    if (!p->strand_tiebreak) return (uint64_t)-1; /* Builder<'h, true, ..>: canonical only */
    if (n >= (1ull << 32)) return (uint64_t)-1;
    const uint32_t l = p->k + p->w - 1;
    if (n < l) return 0;
    uint64_t nwin = n - l + 1;
    uint32_t* wp = (uint32_t*)malloc(sizeof(uint32_t) * nwin);
    if (!wp) return (uint64_t)-1;
    uint64_t got = algo == 0 ? mzo_window_positions_naive(packed, off, n, p, wp)
                             : mzo_window_positions_stream(packed, off, n, p, wp);
    uint64_t m = (uint64_t)-1;
    if (got == nwin) {
        if (algo == 0) { /* per window, directly */
            for (uint64_t j = 0; j < nwin; j++)
                for (uint32_t t = 0; t < l; t++)
                    if (amb_bit(amb, amb_off, j + t)) {
                        wp[j] = MZO_SKIPPED;
                        break;
                    }
        } else { /* running count of ambiguous bases inside the window */
            uint32_t cnt = 0;
            for (uint64_t i = 0; i < n; i++) {
                cnt += amb_bit(amb, amb_off, i);
                if (i >= l) cnt -= amb_bit(amb, amb_off, i - l);
                if (i + 1 >= l && cnt) wp[i + 1 - l] = MZO_SKIPPED;
            }
        }
        if (p->mode == MZO_MINIMIZER) m = mzo_collect_dedup_skip_max(wp, nwin, pos_out);
        else m = mzo_collect_syncmers(wp, nwin, p->w, p->mode == MZO_OPEN_SYNCMER, pos_out);
    }
    free(wp);
    return m;
}

uint64_t mzo_run_range(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p,
                       uint64_t win_begin, uint64_t win_end, uint32_t* pos_out, uint32_t* sk_out,
                       uint64_t cap) {
    if (!params_ok(p)) return (uint64_t)-1;
    const uint32_t l = p->k + p->w - 1;
    if (n < l) return 0;
    uint64_t nwin = n - l + 1;
    if (win_end > nwin) win_end = nwin;
    if (win_begin >= win_end) return 0;
    /* one extra window on the left only to evaluate the dedup flag of win_begin */
    uint64_t first = (win_begin > 0 && p->mode == MZO_MINIMIZER) ? win_begin - 1 : win_begin;
    stream_t s;
    if (stream_init(&s, packed, off, p)) return (uint64_t)-1;
    uint64_t m = 0, sel, prev = ~0ull;
    uint64_t nb = (win_end - first) + l - 1;
    for (uint64_t t = 0; t < nb; t++) {
        if (!stream_step(&s, first, t, &sel)) continue;
        uint64_t j = first + t + 1 - l;
        int emit;
        uint32_t val;
        if (p->mode == MZO_MINIMIZER) {
            emit = (j >= win_begin) && (j == 0 || sel != prev);
            prev = sel;
            val = (uint32_t)sel;
        } else if (p->mode == MZO_CLOSED_SYNCMER) {
            emit = (sel == j || sel == j + p->w - 1);
            val = (uint32_t)j;
        } else {
            emit = (sel == j + p->w / 2);
            val = (uint32_t)j;
        }
        if (emit) {
            if (m >= cap) {
                stream_free(&s);
                return (uint64_t)-1;
            }
            pos_out[m] = val;
            if (sk_out) sk_out[m] = (uint32_t)j;
            m++;
        }
    }
    stream_free(&s);
    return m;
}

/* ------------------------------------------------------------------------------------------
 * values (src/lib.rs:598-629): kmer = sum b[p+j] << 2j ; revcomp = sum (b[p+len-1-j]^2) << 2j
 * ---------------------------------------------------------------------------------------- */
void mzo_values_u64(const uint8_t* packed, uint64_t off, uint32_t len, int canonical,
                    const uint32_t* pos, uint64_t m, uint64_t* out) {
    for (uint64_t e = 0; e < m; e++) {
        uint64_t fwd = 0, rev = 0;
        for (uint32_t j = 0; j < len; j++) {
            fwd |= (uint64_t)mzo_base(packed, off, (uint64_t)pos[e] + j) << (2 * j);
            rev |= (uint64_t)(mzo_base(packed, off, (uint64_t)pos[e] + len - 1 - j) ^ 2u) << (2 * j);
        }
        out[e] = (canonical && rev < fwd) ? rev : fwd;
    }
}

void mzo_values_u128(const uint8_t* packed, uint64_t off, uint32_t len, int canonical,
                     const uint32_t* pos, uint64_t m, uint64_t* out) {
    for (uint64_t e = 0; e < m; e++) {
        unsigned __int128 fwd = 0, rev = 0;
        for (uint32_t j = 0; j < len; j++) {
            fwd |= (unsigned __int128)mzo_base(packed, off, (uint64_t)pos[e] + j) << (2 * j);
            rev |= (unsigned __int128)(mzo_base(packed, off, (uint64_t)pos[e] + len - 1 - j) ^ 2u)
                   << (2 * j);
        }
        unsigned __int128 v = (canonical && rev < fwd) ? rev : fwd;
        out[2 * e] = (uint64_t)v;
        out[2 * e + 1] = (uint64_t)(v >> 64);
    }
}

/* ------------------------------------------------------------------------------------------
 * Multi-threaded chunked run (timed CPU baseline; analogue of the reference benchmark's rayon
 * loop bench/src/bin/paper.rs:442-459, applied to contiguous window ranges of one sequence).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* packed;
    uint64_t off, n;
    const mzo_params* p;
    uint64_t wb, we;
    uint32_t *pos, *sk;
    uint64_t* val;
    uint64_t cap, count;
    int want_sk, want_val;
} mt_job;

static void* mt_worker(void* arg) {
    mt_job* j = (mt_job*)arg;
    j->count = mzo_run_range(j->packed, j->off, j->n, j->p, j->wb, j->we, j->pos,
                             j->want_sk ? j->sk : NULL, j->cap);
    if (j->count != (uint64_t)-1 && j->want_val) {
        uint32_t len = j->p->mode == MZO_MINIMIZER ? j->p->k : j->p->k + j->p->w - 1;
        mzo_values_u64(j->packed, j->off, len, (int)j->p->strand_tiebreak, j->pos, j->count, j->val);
    }
    return NULL;
}

uint64_t mzo_run_mt(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p,
                    int threads, uint32_t* pos_out, uint32_t* sk_out, uint64_t* val_out,
                    uint64_t cap) {
    if (!params_ok(p)) return (uint64_t)-1;
    const uint32_t l = p->k + p->w - 1;
    if (n < l) return 0;
    if (threads < 1) threads = 1;
    uint64_t nwin = n - l + 1;
    if ((uint64_t)threads > nwin) threads = (int)nwin;
    mt_job* jobs = (mt_job*)calloc((size_t)threads, sizeof(mt_job));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    uint64_t per = (nwin + (uint64_t)threads - 1) / (uint64_t)threads;
    int ok = 1;
    for (int t = 0; t < threads; t++) {
        mt_job* j = &jobs[t];
        j->packed = packed, j->off = off, j->n = n, j->p = p;
        j->wb = per * (uint64_t)t;
        j->we = j->wb + per < nwin ? j->wb + per : nwin;
        uint64_t range = j->we > j->wb ? j->we - j->wb : 0;
        /* expected density is 2/(w+1); leave 1.5x head-room, fall back to the worst case
         * (every window emits) for small ranges */
        uint64_t est = range / (p->w + 1) * 3 + 4096;
        j->cap = (range < (1u << 20) || est > range) ? range : est;
        j->want_sk = sk_out != NULL, j->want_val = val_out != NULL;
        j->pos = (uint32_t*)malloc(sizeof(uint32_t) * (j->cap + 1));
        j->sk = j->want_sk ? (uint32_t*)malloc(sizeof(uint32_t) * (j->cap + 1)) : NULL;
        j->val = j->want_val ? (uint64_t*)malloc(sizeof(uint64_t) * (j->cap + 1)) : NULL;
        if (!j->pos || (j->want_sk && !j->sk) || (j->want_val && !j->val)) ok = 0;
    }
    if (ok)
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    if (ok)
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    uint64_t m = 0;
    for (int t = 0; t < threads && ok; t++) {
        mt_job* j = &jobs[t];
        if (j->count == (uint64_t)-1 || m + j->count > cap) {
            ok = 0;
            break;
        }
        memcpy(pos_out + m, j->pos, sizeof(uint32_t) * j->count);
        if (sk_out) memcpy(sk_out + m, j->sk, sizeof(uint32_t) * j->count);
        if (val_out) memcpy(val_out + m, j->val, sizeof(uint64_t) * j->count);
        m += j->count;
    }
    for (int t = 0; t < threads; t++) free(jobs[t].pos), free(jobs[t].sk), free(jobs[t].val);
    free(jobs);
    free(th);
    return ok ? m : (uint64_t)-1;
}

/* ------------------------------------------------------------------------------------------
 * Batched short reads: the reference has no batch entry point, callers loop
 *   for s in &seqs { v.clear(); builder.run(s, &mut v) }      (bench/src/bin/paper.rs:98-105,
 * examples/bench.rs:63-89).  This is that loop over reads laid out at a fixed byte stride, every
 * read a stand-alone sequence (positions relative to the read), results as CSR.  One scratch
 * ring per thread, reused for every read (the reference keeps its scratch thread_local,
 * src/lib.rs:217-219); threads take contiguous read ranges (rayon-over-reads analogue) and write
 * straight into disjoint slices of the caller's arrays, closed up afterwards.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* packed;
    uint64_t stride_bytes, r0, r1;
    uint32_t read_len;
    const mzo_params* p;
    uint64_t* offsets; /* offsets[r + 1] = entries of read r (turned into a prefix sum later) */
    uint32_t *pos, *sk;
    uint64_t* val;
    uint64_t cap, count;
    int fail;
} rd_job;

static void stream_reset(stream_t* s, uint64_t off) {
    s->off = off;
    s->fw = s->rc = 0;
    s->tg = -(int64_t)s->l;
    for (uint32_t i = 0; i < s->w; i++) s->ringl[i] = ~0ull, s->ringr[i] = 0;
    s->ridx = 0;
    s->prel = ~0ull;
    s->prer = 0;
}

static void* rd_worker(void* arg) {
    rd_job* j = (rd_job*)arg;
    const mzo_params* p = j->p;
    const uint32_t l = p->k + p->w - 1, n = j->read_len;
    const uint32_t len = p->mode == MZO_MINIMIZER ? p->k : l;
    stream_t s;
    j->count = 0;
    if (stream_init(&s, j->packed, 0, p)) {
        j->fail = 1;
        return NULL;
    }
    uint64_t m = 0;
    for (uint64_t r = j->r0; r < j->r1; r++) {
        const uint64_t m0 = m;
        if (n >= l) {
            stream_reset(&s, r * j->stride_bytes * 4);
            uint64_t sel, prev = ~0ull;
            for (uint64_t t = 0; t < n; t++) {
                if (!stream_step(&s, 0, t, &sel)) continue;
                const uint64_t w0 = t + 1 - l;
                int emit;
                uint32_t v;
                if (p->mode == MZO_MINIMIZER) emit = (w0 == 0 || sel != prev), prev = sel, v = (uint32_t)sel;
                else if (p->mode == MZO_CLOSED_SYNCMER) emit = (sel == w0 || sel == w0 + p->w - 1), v = (uint32_t)w0;
                else emit = (sel == w0 + p->w / 2), v = (uint32_t)w0;
                if (emit) {
                    if (m >= j->cap) {
                        j->fail = 1;
                        stream_free(&s);
                        return NULL;
                    }
                    j->pos[m] = v;
                    if (j->sk) j->sk[m] = (uint32_t)w0;
                    m++;
                }
            }
            if (j->val)
                mzo_values_u64(j->packed, r * j->stride_bytes * 4, len, (int)p->strand_tiebreak, j->pos + m0,
                               m - m0, j->val + m0);
        }
        j->offsets[r + 1] = m - m0;
    }
    stream_free(&s);
    j->count = m;
    return NULL;
}

uint64_t mzo_run_reads(const uint8_t* packed, uint64_t n_reads, uint64_t stride_bytes, uint32_t read_len,
                       const mzo_params* p, int threads, uint64_t* offsets_out, uint32_t* pos_out,
                       uint32_t* sk_out, uint64_t* val_out, uint64_t cap) {
    if (!params_ok(p)) return (uint64_t)-1;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_reads) threads = n_reads ? (int)n_reads : 1;
    rd_job* jobs = (rd_job*)calloc((size_t)threads, sizeof(rd_job));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    const uint64_t per = (n_reads + (uint64_t)threads - 1) / (uint64_t)threads;
    const uint64_t slice = cap / (uint64_t)threads;
    offsets_out[0] = 0;
    for (int t = 0; t < threads; t++) {
        rd_job* j = &jobs[t];
        j->packed = packed, j->stride_bytes = stride_bytes, j->read_len = read_len, j->p = p;
        j->r0 = per * (uint64_t)t < n_reads ? per * (uint64_t)t : n_reads;
        j->r1 = j->r0 + per < n_reads ? j->r0 + per : n_reads;
        j->offsets = offsets_out;
        j->pos = pos_out + slice * (uint64_t)t;
        j->sk = sk_out ? sk_out + slice * (uint64_t)t : NULL;
        j->val = val_out ? val_out + slice * (uint64_t)t : NULL;
        j->cap = slice;
    }
    if (threads == 1) rd_worker(&jobs[0]);
    else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, rd_worker, &jobs[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    uint64_t m = 0;
    int ok = 1;
    for (int t = 0; t < threads; t++) {
        rd_job* j = &jobs[t];
        if (j->fail) {
            ok = 0;
            break;
        }
        if (t && j->count) { /* close the gap between the thread slices */
            memmove(pos_out + m, j->pos, sizeof(uint32_t) * j->count);
            if (sk_out) memmove(sk_out + m, j->sk, sizeof(uint32_t) * j->count);
            if (val_out) memmove(val_out + m, j->val, sizeof(uint64_t) * j->count);
        }
        m += j->count;
    }
    if (ok)
        for (uint64_t r = 0; r < n_reads; r++) offsets_out[r + 1] += offsets_out[r];
    free(jobs);
    free(th);
    return ok ? m : (uint64_t)-1;
}

/* ------------------------------------------------------------------------------------------
 * synthetic input
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

void mzo_synth_packed(uint64_t seed, uint64_t n_bases, uint8_t* out) {
    uint64_t nbytes = (n_bases + 3) / 4;
    uint64_t nwords = (nbytes + 7) / 8;
    for (uint64_t i = 0; i < nwords; i++) {
        uint64_t v = splitmix64(seed + i);
        uint64_t rem = nbytes - i * 8;
        memcpy(out + i * 8, &v, rem < 8 ? rem : 8);
    }
    /* zero the unused high bits of the last byte so buffers compare equal */
    if (n_bases & 3) out[nbytes - 1] &= (uint8_t)((1u << (2 * (n_bases & 3))) - 1);
}
