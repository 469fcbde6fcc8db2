/*
 * mzoracle.h -- CPU ORACLE for the simd-minimizers random-minimizer path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` legs may load it.  The product
 * (simd-minimizers_b200/) never links, imports or falls back to anything here.
 *
 * It is a plain-C restatement of the reference algorithm (rust-seq/simd-minimizers
 * v3.0.0, file:line citations are relative to the reference tree), written from the
 * semantics, not translated from the Rust/AVX2 code.
 *
 * Parity status:
 *   - NtHasher (unseeded): PINNED by the reference's own known-answer vectors
 *       src/lib.rs:92-99   forward  ACGTGCTCAGAGACTCAG k=5 w=7 -> [4,5,8,13]
 *       src/lib.rs:109-129 canonical ACGTGCTCAGAGACTCAGAGGA   -> [0,7,9,15] + 4 values
 *       src/lib.rs:132-135 reverse complement                 -> [2,8,10,17]
 *     (tests/test_oracle_golden.py).  The hash lives in the un-vendored crate
 *     seq-hash 0.2.0 (Cargo.lock:885-891); its published algorithm is restated in
 *     mzo_hasher_nt().
 *   - MulHasher and every *seeded* hasher: PARITY UNPINNED -- the reference holds
 *     only differential tests for them (src/test.rs:81-83,107-109) and seq-hash is
 *     absent; mzo_hasher_mul() is a recollection of seq-hash's mulHash.
 *   - window-min / strand / dedup / super-k-mer / syncmer / value stages: pinned by
 *     src/test.rs:335-356,485-515,578-597 and the rc-symmetry properties
 *     src/test.rs:113-152,642-708.
 */
#ifndef MZORACLE_H
#define MZORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Per-base hash tables, indexed by 2-bit code (A=0,C=1,T=2,G=3).  rot = bits of
 * rotation per base (7 in seq-hash). */
typedef struct {
    uint32_t f[4];
    uint32_t c[4];
    uint32_t rot;
    uint32_t canonical; /* 1: h = fw + rc (wrapping), 0: h = fw */
} mzo_hasher;

enum { MZO_MINIMIZER = 0, MZO_CLOSED_SYNCMER = 1, MZO_OPEN_SYNCMER = 2 };

typedef struct {
    uint32_t k, w;
    uint32_t strand_tiebreak; /* canonical builder: leftmost/rightmost by TG count */
    uint32_t mode;            /* MZO_* */
    mzo_hasher hasher;
} mzo_params;

void mzo_hasher_nt(mzo_hasher* h, int canonical);
void mzo_hasher_mul(mzo_hasher* h, int canonical);

/* 2-bit code of base i of a packed sequence starting `off` bases into `packed`. */
static inline uint32_t mzo_base(const uint8_t* packed, uint64_t off, uint64_t i) {
    uint64_t p = off + i;
    return (packed[p >> 2] >> (2 * (p & 3))) & 3u;
}

/* (ascii >> 1) & 3 packing; returns number of bytes written. */
uint64_t mzo_pack_ascii(const char* ascii, uint64_t n, uint8_t* out);
/* reverse complement of n bases into a fresh packed buffer (offset 0). */
void mzo_revcomp(const uint8_t* packed, uint64_t off, uint64_t n, uint8_t* out);

/* 32-bit k-mer hash of the k-mer starting at base i (direct evaluation). */
uint32_t mzo_hash_kmer(const uint8_t* packed, uint64_t off, uint64_t i, uint32_t k,
                       const mzo_hasher* h);

/* Per-window selected k-mer position P[j], j in [0, n-l], brute force per window.
 * Returns number of windows. */
uint64_t mzo_window_positions_naive(const uint8_t* packed, uint64_t off, uint64_t n,
                                    const mzo_params* p, uint32_t* out);
/* Same, streaming O(n): rolling hash + two-stacks sliding min + rolling TG count. */
uint64_t mzo_window_positions_stream(const uint8_t* packed, uint64_t off, uint64_t n,
                                     const mzo_params* p, uint32_t* out);

/* Collectors over a per-window position stream (hash independent). */
uint64_t mzo_collect_dedup(const uint32_t* win_pos, uint64_t nwin, uint32_t* pos_out,
                           uint32_t* sk_out /* may be NULL */);
uint64_t mzo_collect_syncmers(const uint32_t* win_pos, uint64_t nwin, uint32_t w, int open,
                              uint32_t* out);

/* Whole path, single thread.  algo: 0 = naive, 1 = streaming.  Output arrays must
 * hold n-l+1 entries (worst case).  Returns count, or (uint64_t)-1 on bad params. */
uint64_t mzo_run(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p,
                 int algo, uint32_t* pos_out, uint32_t* sk_out);

/* Ambiguous bases (PackedNSeq; src/lib.rs:451-496): `amb` holds one bit per base (base i ->
 * bit (amb_off+i)&7 of byte (amb_off+i)>>3).  Windows containing an ambiguous base produce
 * nothing; canonical builders only.  mzo_pack_ascii_n: everything except ACGTacgt is ambiguous. */
uint64_t mzo_pack_ascii_n(const char* ascii, uint64_t n, uint8_t* packed_out, uint8_t* amb_out);
uint64_t mzo_collect_dedup_skip_max(const uint32_t* win_pos, uint64_t nwin, uint32_t* pos_out);
uint64_t mzo_run_skip_ambiguous(const uint8_t* packed, uint64_t off, uint64_t n, const uint8_t* amb,
                                uint64_t amb_off, const mzo_params* p, int algo, uint32_t* pos_out);

/* Fused streaming path writing straight into pos/sk (no per-window array);
 * used for the timed CPU baseline.  Handles windows [win_begin, win_end). */
uint64_t mzo_run_range(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p,
                       uint64_t win_begin, uint64_t win_end, uint32_t* pos_out,
                       uint32_t* sk_out, uint64_t cap);

/* Multi-threaded (pthread) chunked run of the streaming path; chunks overlap l-1
 * bases and seams are stitched like the reference's lane seams (src/collect.rs:252-272).
 * Outputs positions (+ optional values) into caller arrays of capacity cap.
 * Returns count ((uint64_t)-1 on error / overflow). */
uint64_t mzo_run_mt(const uint8_t* packed, uint64_t off, uint64_t n, const mzo_params* p,
                    int threads, uint32_t* pos_out, uint32_t* sk_out, uint64_t* val_out,
                    uint64_t cap);

/* Batched short reads = the caller loop `for s in &seqs { v.clear(); builder.run(s, &mut v) }`
 * (bench/src/bin/paper.rs:98-105, examples/bench.rs:63-89) over reads of `read_len` bases laid
 * out every `stride_bytes` bytes; CSR result (offsets_out: n_reads + 1 entries), positions
 * relative to the read.  `threads` workers take contiguous read ranges.  val_out (may be NULL):
 * u64 values, k <= 32 (minimizers) / l <= 32 (syncmers).  Returns the entry count,
 * (uint64_t)-1 on bad parameters or when cap / threads entries do not hold a thread's share. */
uint64_t mzo_run_reads(const uint8_t* packed, uint64_t n_reads, uint64_t stride_bytes, uint32_t read_len,
                       const mzo_params* p, int threads, uint64_t* offsets_out, uint32_t* pos_out,
                       uint32_t* sk_out, uint64_t* val_out, uint64_t cap);

/* k-mer values for positions (src/lib.rs:598-629). len <= 32 for u64, <= 64 for u128
 * (lo,hi pairs). */
void mzo_values_u64(const uint8_t* packed, uint64_t off, uint32_t len, int canonical,
                    const uint32_t* pos, uint64_t m, uint64_t* out);
void mzo_values_u128(const uint8_t* packed, uint64_t off, uint32_t len, int canonical,
                     const uint32_t* pos, uint64_t m, uint64_t* out_lo_hi);

/* splitmix64 counter-based synthetic packed DNA: byte-identical to the generator the
 * bench uses on device. word i (32 bases) = splitmix64(seed + i). */
void mzo_synth_packed(uint64_t seed, uint64_t n_bases, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
