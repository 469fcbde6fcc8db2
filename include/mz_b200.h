/*
 * mz_b200.h -- C ABI of the B200-native random-minimizer path (libmzb200.so).
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no torch / CUDA types.
 * A Rust `-sys` crate binds exactly these symbols (rust/mzb200-sys/src/lib.rs, and
 * INTEGRATION.md shows the call each reference entry point turns into).  Every entry point
 * cites the reference interface (rust-seq/simd-minimizers v3.0.0, paths relative to the
 * reference tree) whose body it replaces.
 *
 * There is NO CPU fallback behind this ABI: without a CUDA device every compute entry point
 * returns MZ_ERR_NO_DEVICE / MZ_ERR_CUDA.
 */
#ifndef MZ_B200_H
#define MZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZ_ABI_VERSION 1u

/* ---- status codes.  The reference panics (assert!) on bad parameters; nothing unwinds
 *      across this ABI -- the host shim turns codes back into the reference's panics. ---- */
enum {
    MZ_OK = 0,
    MZ_ERR_BAD_ARG = 1,       /* null pointer, k == 0, unknown mode ...                          */
    MZ_ERR_W_RANGE = 2,       /* w == 0 or w >= 2^15            (src/sliding_min.rs:91-95)        */
    MZ_ERR_TOO_LONG = 3,      /* n >= 2^32 bases                (src/sliding_min.rs:96-99)        */
    MZ_ERR_EVEN_L = 4,        /* canonical needs odd l=k+w-1    (src/canonical.rs:13-16)          */
    MZ_ERR_OPEN_EVEN_W = 5,   /* open syncmers need odd w       (src/syncmers.rs:24-29)           */
    MZ_ERR_NOT_CANONICAL = 6, /* canonical builder + forward hasher (src/minimizers.rs:81,139)    */
    MZ_ERR_VALUE_WIDTH = 7,   /* values wider than requested bits (k or l > 32 / 64)              */
    MZ_ERR_CAPACITY = 8,      /* output buffers too small; mz_out.count holds the needed size     */
    MZ_ERR_UNSUPPORTED = 9,   /* parameter combination not implemented (very large w)             */
    MZ_ERR_NO_DEVICE = 10,    /* no CUDA device / bad device id                                   */
    MZ_ERR_CUDA = 11,         /* CUDA runtime error; see mz_last_error()                          */
    MZ_ERR_NOMEM = 12
};

/* Builder<.., SYNCMER> const generic (src/lib.rs:221-230). */
enum { MZ_MODE_MINIMIZER = 0, MZ_MODE_CLOSED_SYNCMER = 1, MZ_MODE_OPEN_SYNCMER = 2 };

/* One POD describing a Builder + hasher (src/lib.rs:225-230 fields k, w, hasher, sk_pos).
 * Hashers are passed as their per-base tables, indexed by packed 2-bit code (A=0 C=1 T=2 G=3):
 *   fw(i) = XOR_j rotl(f[b[i+j]], rot*(k-1-j)),  rc(i) = XOR_j rotl(c[b[i+j]], rot*j),
 *   hash  = hash_canonical ? fw + rc : fw        (seq-hash 0.2.0 NtHasher / MulHasher). */
typedef struct mz_params {
    uint32_t k;               /* k-mer length (hasher.k())                                       */
    uint32_t w;               /* k-mers per window                                               */
    uint32_t mode;            /* MZ_MODE_*                                                       */
    uint32_t strand_tiebreak; /* Builder CANONICAL: leftmost/rightmost by TG count               */
    uint32_t hash_canonical;  /* KmerHasher::is_canonical()                                      */
    uint32_t rot;             /* rotation per base (7)                                           */
    uint32_t f[4];            /* forward table                                                   */
    uint32_t c[4];            /* complement table                                                */
    uint32_t want_sk;         /* .super_kmers(): also emit first-window index (minimizer mode)   */
    uint32_t value_bits;      /* 0: positions only; 64: values_u64; 128: values_u128             */
    uint32_t reserved[2];
} mz_params;

/* Caller-owned output buffers (the reference appends into the caller's Vec<u32>,
 * src/lib.rs:80-81; here the caller passes spare capacity).  `val` holds `count` u64 when
 * value_bits == 64, or 2*count u64 (lo, hi) when value_bits == 128. */
typedef struct mz_out {
    uint32_t* pos;     /* minimizer positions, or syncmer window starts                          */
    uint32_t* sk;      /* super-k-mer first-window indices (want_sk), else may be NULL           */
    uint64_t* val;     /* k-mer / l-mer values (value_bits != 0), else may be NULL               */
    uint64_t capacity; /* entries available in each non-NULL array                               */
    uint64_t count;    /* out: entries produced (or required, on MZ_ERR_CAPACITY)                */
} mz_out;

/* Timing breakdown of the last mz_run* call on a context (milliseconds, CUDA events). */
typedef struct mz_timing {
    float h2d_ms, kernel_ms, d2h_ms, total_ms;
    uint32_t kernel_launches; /* launches of this library's own kernels                          */
    uint32_t reserved;
} mz_timing;

typedef struct mz_ctx mz_ctx;

/* ---- library ---- */
uint32_t mz_abi_version(void);
const char* mz_strerror(int code);
/* Thread-local description of the last MZ_ERR_CUDA on the calling thread. */
const char* mz_last_error(void);
int mz_device_count(int* n);

/* ---- hasher helpers (replace `H::new(k)`, src/lib.rs:391; seq-hash NtHasher/MulHasher).
 *      Fill everything in *p except want_sk / value_bits.  `canonical` is the Builder's
 *      CANONICAL flag; the default hasher is NtHasher<CANONICAL> (src/lib.rs:240-321). ---- */
int mz_params_nthash(mz_params* p, uint32_t k, uint32_t w, uint32_t mode, uint32_t canonical);
int mz_params_mulhash(mz_params* p, uint32_t k, uint32_t w, uint32_t mode, uint32_t canonical);
/* Parity status of the built-in tables: the NtHasher tables are pinned by the
 * reference's own k = 5 vectors (src/lib.rs:92-135); the MulHasher tables, f(b) = b * 0x27220a95
 * (the low half of the constant at bench/src/rescan_daniel.rs:38) and c(b) = f(b ^ 2), restate
 * seq-hash 0.2.0, which is not in the reference tree, and are NOT pinned by any reference vector:
 * tests/test_reference_dump.py derives every hasher's tables from a dump of the real crate
 * (tools/dump_reference_vectors.rs) and compares.  A caller that has the real hasher object can
 * always pass its tables with mz_params_set_tables. */
/* Replace only the hasher of *p (Builder::hasher, src/lib.rs:327-337). hash_canonical is the
 * hasher's own RC flag and may differ from strand_tiebreak (forward builder + canonical hasher,
 * src/minimizers.rs:69-71). */
int mz_params_set_nthash(mz_params* p, uint32_t hash_canonical);
int mz_params_set_mulhash(mz_params* p, uint32_t hash_canonical);
/* Any other table hasher (seeded hashers: `H::new_with_seed(k, seed)`, src/lib.rs:157,
 * src/test.rs:287 -- seq-hash derives per-base tables from the seed; the Rust shim reads them out
 * of the hasher object and passes them here). */
int mz_params_set_tables(mz_params* p, const uint32_t f[4], const uint32_t c[4], uint32_t rot,
                         uint32_t hash_canonical);
/* Same checks the reference asserts on (see the error codes). */
int mz_params_validate(const mz_params* p, uint64_t n_bp);

/* ---- context: streams + scratch for a set of devices.  One context per host thread, or
 *      serialise calls yourself (the reference keeps its scratch thread_local,
 *      src/lib.rs:217-219). device_ids == NULL, n == 0 -> current device only. ---- */
int mz_ctx_create(const int* device_ids, int n_devices, mz_ctx** ctx);
void mz_ctx_destroy(mz_ctx* ctx);
int mz_ctx_device_count(const mz_ctx* ctx);

/* ---- pinned host memory for full-speed transfers (optional; any host pointer works) ---- */
int mz_host_alloc(void** p, size_t bytes);
void mz_host_free(void* p);

/*
 * mz_run -- the whole path for one sequence, host buffers in, host buffers out.
 * Replaces Builder::run_impl / run_with_buf (src/lib.rs:386-448, 554-576) followed by
 * Output::values_u64/u128 (src/lib.rs:584-629):
 *   minimizer_positions(seq,k,w)            -> mz_params_nthash(.., MINIMIZER, 0) + mz_run
 *   canonical_minimizer_positions(seq,k,w)  -> mz_params_nthash(.., MINIMIZER, 1) + mz_run
 *   .super_kmers(&mut sk)                   -> want_sk = 1
 *   .values_u64() / .values_u128()          -> value_bits = 64 / 128
 *   closed/open syncmers                    -> mode
 * `packed` is PackedSeq storage: 4 bases per byte, first base in the low bits, A=0 C=1 T=2 G=3;
 * the sequence is bases [bp_offset, bp_offset + n_bp) of it (PackedSeq.offset is 0..3, but any
 * offset is accepted).  When the context holds several devices the windows are split into
 * contiguous shards (k+w-2 base halo + one window for the seam) and outputs are concatenated
 * in order.  Results are written from index 0 of out->pos/sk/val; the *append* quirk of the
 * reference's SIMD collector (src/collect.rs:257,267) is applied by the host shim.
 */
int mz_run(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset,
           uint64_t n_bp, mz_out* out);

/*
 * mz_run_device -- same, but input and outputs are DEVICE pointers on device `dev_index`
 * of the context, and only windows [win_begin, win_end) are produced (win_end == 0 -> all).
 * `d_packed` must hold the whole sequence [bp_offset, bp_offset+n_bp).  out->count is written
 * on the host after the stream is synchronised.  Used for device-resident timing and by
 * callers that keep consuming on the GPU.
 * Device-buffer requirements (checked where possible, MZ_ERR_BAD_ARG otherwise): pos / sk 4-byte
 * aligned, val 8-byte aligned (16-byte for value_bits == 128: (lo, hi) is one 16-byte store);
 * d_packed (and d_ambiguous) may start at any byte, but are read as aligned 32-bit words, so the
 * bytes up to the next 4-byte boundary behind the last base must be readable (any cudaMalloc'ed
 * buffer is; a sub-allocation that ends exactly at the end of its arena needs 3 bytes of padding).
 */
int mz_run_device(mz_ctx* ctx, int dev_index, const mz_params* p, const void* d_packed,
                  uint64_t bp_offset, uint64_t n_bp, uint64_t win_begin, uint64_t win_end,
                  mz_out* d_out);

/*
 * mz_run_batch -- many independent short sequences in one launch (the reference has no batch
 * entry point; callers loop `for s in &seqs { builder.run(s, &mut v) }`,
 * bench/src/bin/paper.rs:98-105, examples/bench.rs:63-89).  Read r is bases
 * [read_start_bp[r], read_start_bp[r] + read_len_bp[r]) of `packed`; with read_start_bp == NULL
 * reads are laid out at a fixed stride of `stride_bytes` bytes and fixed length `fixed_len_bp`.
 * Output is CSR: out_offsets[r]..out_offsets[r+1] index pos/sk/val (n_reads + 1 entries);
 * positions are relative to the read start, exactly what the per-read reference call returns.
 */
int mz_run_batch(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t packed_bytes,
                 uint64_t n_reads, const uint64_t* read_start_bp, const uint32_t* read_len_bp,
                 uint64_t stride_bytes, uint32_t fixed_len_bp, uint64_t* out_offsets, mz_out* out);

/*
 * ASCII ingestion on the device (the step before the path).  `AsciiSeq` /
 * `PackedSeqVec::from_ascii` map a character to its 2-bit code with (c >> 1) & 3
 * (A/a=0 C/c=1 T/t=2 G/g=3; callers bench/src/lib.rs:48-82 pack on the host first).
 * mz_pack_ascii packs n characters into (n+3)/4 bytes of PackedSeq storage (host in, host out);
 * mz_run_ascii packs on the device and runs the path without the packed bytes ever visiting
 * the host.  Results equal mz_run on the packed sequence.
 */
int mz_pack_ascii(mz_ctx* ctx, const char* ascii, uint64_t n, uint8_t* packed_out);
int mz_run_ascii(mz_ctx* ctx, const mz_params* p, const char* ascii, uint64_t n, mz_out* out);

/*
 * Ambiguous bases.  Replaces
 *   Builder<'h, true, H, (), SYNCMERS>::run_skip_ambiguous_windows(nseq: PackedNSeq, &mut Vec<u32>)
 *   (src/lib.rs:451-496; stream src/minimizers.rs:169-214; collector SKIP_MAX
 *   src/collect.rs:213,243 and src/intrinsics/dedup.rs:147-155; syncmers src/syncmers.rs:152).
 * A PackedNSeq is the packed 2-bit sequence plus one ambiguity bit per base: base i of the
 * sequence is bit (amb_bit_offset + i) & 7 of byte (amb_bit_offset + i) >> 3 of `ambiguous`
 * (LSB first; packed-seq 5.0.0's BitSeq is not in the reference tree, the Rust shim passes its
 * storage and offset).  Every window of l = k+w-1 bases that contains an ambiguous base produces
 * nothing; all other windows produce exactly what mz_run produces for them, and the first clean
 * window after an ambiguous stretch always emits (the reference compares it against SKIPPED).
 * As in the reference: canonical builders only (MZ_ERR_NOT_CANONICAL otherwise), no super-k-mer
 * output (want_sk must be 0), minimizers and closed/open syncmers, values as for mz_run.
 * mz_run_device_skip_ambiguous takes device pointers (like mz_run_device).
 * mz_pack_ascii_n is `PackedNSeqVec::from_ascii`: 2-bit codes as mz_pack_ascii plus the mask,
 * (n+7)/8 bytes, bit set for every character outside ACGTacgt; mz_run_ascii_skip_ambiguous
 * builds both on the device and runs the path.
 */
int mz_run_skip_ambiguous(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset,
                          uint64_t n_bp, const uint8_t* ambiguous, uint64_t amb_bit_offset, mz_out* out);
int mz_run_device_skip_ambiguous(mz_ctx* ctx, int dev_index, const mz_params* p, const void* d_packed,
                                 uint64_t bp_offset, uint64_t n_bp, const void* d_ambiguous,
                                 uint64_t amb_bit_offset, uint64_t win_begin, uint64_t win_end,
                                 mz_out* d_out);
int mz_pack_ascii_n(mz_ctx* ctx, const char* ascii, uint64_t n, uint8_t* packed_out, uint8_t* ambiguous_out);
int mz_run_ascii_skip_ambiguous(mz_ctx* ctx, const mz_params* p, const char* ascii, uint64_t n, mz_out* out);

int mz_last_timing(const mz_ctx* ctx, mz_timing* t);

/*
 * mz_values -- Output::values_u64 / values_u128 / pos_and_values_* (src/lib.rs:584-629), the
 * reference's LAZY value iterators: the k-mer (minimizers) or l-mer (syncmers) at each of `n_pos`
 * positions of the sequence, canonical builders return min(kmer, revcomp).  `pos` may be any
 * positions (the reference iterates the caller's whole Vec, including entries of earlier runs);
 * a position whose k-mer runs past the sequence gives MZ_ERR_BAD_ARG (read_kmer asserts).
 * val_out: n_pos u64 (value_bits 64) or 2*n_pos u64 as (lo, hi) (value_bits 128).
 * The shims run positions-only by default and call this when values are asked for; the sequence
 * of the last single-launch mz_run on this context is still resident on the device and is not
 * uploaded again (same `packed`, bp_offset, n_bp -- the caller must not have modified it, which
 * the reference's `Output<'o, 's>` borrow guarantees).  Callers that always want values set
 * mz_params.value_bits in mz_run instead (fused into the same kernel).
 */
int mz_values(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset, uint64_t n_bp,
              const uint32_t* pos, uint64_t n_pos, uint32_t value_bits, uint64_t* val_out);

/*
 * mz_run_bucket_stats -- the path plus a first consumer that stays on the device.
 * A super-k-mer is a maximal run of windows with the same minimizer (bench/src/minimizer.rs:3-36,
 * "Problem C"; the i-th one starts at window sk[i] of a `.super_kmers()` run, src/lib.rs:339-352,
 * and ends where the next one starts).  Tools downstream of a minimizer scan shard super-k-mers by
 * their minimizer; this call does that sharding where the minimizers are, in HBM, and returns only
 * the histograms: for every bucket b < n_buckets (<= 16384)
 *     superkmers_out[b] = super-k-mers whose minimizer falls into b,
 *     windows_out[b]    = windows covered by them (sum over b = n_bp - l + 1),
 * with bucket = floor(mix64(value) * n_buckets / 2^64), value = the minimizer's k-mer as
 * values_u64() reports it (canonical builders: min(kmer, revcomp)), mix64 = the splitmix64 finaliser
 * (x += 0x9E3779B97F4A7C15; x = (x ^ x>>30) * 0xBF58476D1CE4E5B9; x = (x ^ x>>27) * 0x94D049BB133111EB;
 * x ^ x>>31).  Minimizer builders with k <= 32 only.  Nothing but the packed input (host -> device)
 * and 16 * n_buckets bytes (device -> host) crosses the bus; all devices of the context take chunks.
 * Equivalent to running mz_run with want_sk = 1, value_bits = 64 and building the histograms on
 * the host, which moves 16 bytes per minimizer over PCIe instead.
 */
int mz_run_bucket_stats(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset, uint64_t n_bp,
                        uint32_t n_buckets, uint64_t* superkmers_out, uint64_t* windows_out, uint64_t* n_minimizers);

/*
 * mz_pcie_probe -- measured host <-> device copy rate of the context's devices, all of them at
 * the same time (pinned host memory, `reps` copies of `bytes_per_device` per device and
 * direction).  This is the ceiling of every end-to-end number of mz_run / mz_run_batch: one PCIe
 * link per device, host memory and root complex shared.  GB/s, summed over the devices.
 * The reference has no counterpart (it never leaves the host); bench.py reports it as
 * e2e.pcie_peak_gbs.
 */
typedef struct mz_pcie_result {
    double h2d_gbs;       /* host -> device alone                                                */
    double d2h_gbs;       /* device -> host alone                                                */
    double bidir_h2d_gbs; /* both directions at once: host -> device share                       */
    double bidir_d2h_gbs; /*                          device -> host share                       */
    uint32_t n_devices;
    uint32_t reserved;
} mz_pcie_result;
int mz_pcie_probe(mz_ctx* ctx, uint64_t bytes_per_device, uint32_t reps, mz_pcie_result* res);

/*
 * mz_alu_probe -- measured INT32 ALU-pipe peak of device `dev_index` of the context: lane-ops per
 * second of a LOP3 / SHF / VIMNMX / PRMT mix (the instruction mix of the window-minimum loop) with
 * eight independent chains per thread on every SM, no memory traffic.  The minimizer kernels are
 * bound by this pipe, not by HBM; bench.py divides the kernels' integer work by it (alu_roofline).
 */
typedef struct mz_alu_result {
    double lane_ops_per_s; /* 32 lanes x warp instructions retired on the ALU pipe per second      */
    float ms;              /* duration of the probe launch                                         */
    uint32_t sm_count;
} mz_alu_result;
int mz_alu_probe(mz_ctx* ctx, int dev_index, mz_alu_result* res);

#ifdef __cplusplus
}
#endif
#endif /* MZ_B200_H */
