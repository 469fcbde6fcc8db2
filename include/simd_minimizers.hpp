// simd_minimizers.hpp -- header-only C++ mirror of the reference builder API
// (rust-seq/simd-minimizers v3.0.0, src/lib.rs:225-654) over the C ABI in mz_b200.h.
//
//   auto pos  = simd_minimizers::canonical_minimizer_positions(seq, k, w);
//   std::vector<uint32_t> p, sk;
//   auto out  = simd_minimizers::canonical_minimizers(k, w).hasher(h).super_kmers(sk).run(seq, p);
//   auto vals = out.values_u64();
//
// The reference's assert!/panic! become std::invalid_argument with the reference's message.
// Link with -lmzb200.  No CPU fallback: without a CUDA device every run() throws.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mz_b200.h"

namespace simd_minimizers {

// packed_seq::PackedSeq: 4 bases per byte, first base in the low bits, A=0 C=1 T=2 G=3.
struct PackedSeq {
    const uint8_t* data = nullptr;
    uint64_t offset = 0;  // bases into data
    uint64_t len = 0;     // bases
    PackedSeq slice(uint64_t b, uint64_t e) const { return {data, offset + b, e - b}; }
};

// packed_seq::PackedNSeq: PackedSeq + one ambiguity bit per base (bit (amb_offset+i)&7 of byte
// (amb_offset+i)>>3 of `ambiguous`).
struct PackedNSeq {
    PackedSeq seq;
    const uint8_t* ambiguous = nullptr;
    uint64_t amb_offset = 0;
    PackedNSeq slice(uint64_t b, uint64_t e) const { return {seq.slice(b, e), ambiguous, amb_offset + b}; }
};

struct Hasher {  // seq-hash NtHasher<RC> / MulHasher<RC>, or any per-base table hasher (seeded ones)
    enum Kind { Nt, Mul, Tables } kind = Nt;
    uint32_t k = 0;
    bool canonical = true;
    uint32_t f[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, rot = 7;
    static Hasher nt(uint32_t k, bool rc = true) { return {Nt, k, rc}; }
    static Hasher mul(uint32_t k, bool rc = true) { return {Mul, k, rc}; }
    // `H::new_with_seed(k, seed)` (src/lib.rs:157): seq-hash derives per-base tables from the seed;
    // pass those tables (indexed by packed code A=0 C=1 T=2 G=3)
    static Hasher tables(uint32_t k, const uint32_t (&f)[4], const uint32_t (&c)[4], uint32_t rot = 7, bool rc = true) {
        Hasher h{Tables, k, rc};
        for (int b = 0; b < 4; b++) h.f[b] = f[b], h.c[b] = c[b];
        h.rot = rot;
        return h;
    }
    bool is_canonical() const { return canonical; }
};

namespace detail {
inline void check(int rc) {
    if (rc == MZ_OK) return;
    std::string msg = mz_strerror(rc);
    if (rc == MZ_ERR_CUDA) msg += std::string(": ") + mz_last_error();
    if (rc >= MZ_ERR_W_RANGE && rc <= MZ_ERR_VALUE_WIDTH) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
inline mz_ctx* thread_ctx() {  // thread_local scratch, like src/lib.rs:217-219
    struct Holder {
        mz_ctx* c = nullptr;
        ~Holder() { mz_ctx_destroy(c); }
    };
    thread_local Holder h;
    if (!h.c) check(mz_ctx_create(nullptr, 0, &h.c));
    return h.c;
}
}  // namespace detail

// src/lib.rs:232-237, 579-630.  Values are lazy, as in the reference: run() moves positions only,
// values_*() computes the k-mers (l-mers for syncmers) of ALL of min_pos on the device (mz_values;
// the sequence of the run is still resident there).
class Output {
public:
    Output(uint32_t len, mz_params p, PackedSeq seq, const std::vector<uint32_t>* pos) : len_(len), p_(p), seq_(seq), pos_(pos) {}
    uint32_t len() const { return len_; }  // k for minimizers, k+w-1 for syncmers
    std::vector<uint64_t> values_u64() const {
        std::vector<uint64_t> v(pos_->size());
        detail::check(mz_values(detail::thread_ctx(), &p_, seq_.data, seq_.offset, seq_.len, pos_->data(), pos_->size(), 64, v.data()));
        return v;
    }
    // (lo, hi) pairs
    std::vector<std::pair<uint64_t, uint64_t>> values_u128() const {
        std::vector<uint64_t> raw(2 * pos_->size());
        detail::check(mz_values(detail::thread_ctx(), &p_, seq_.data, seq_.offset, seq_.len, pos_->data(), pos_->size(), 128, raw.data()));
        std::vector<std::pair<uint64_t, uint64_t>> v(pos_->size());
        for (size_t i = 0; i < v.size(); i++) v[i] = {raw[2 * i], raw[2 * i + 1]};
        return v;
    }
    std::vector<std::pair<uint32_t, uint64_t>> pos_and_values_u64() const {
        const auto vals = values_u64();
        std::vector<std::pair<uint32_t, uint64_t>> v(vals.size());
        for (size_t i = 0; i < v.size(); i++) v[i] = {(*pos_)[i], vals[i]};
        return v;
    }
    // (pos, (lo, hi)) -- src/lib.rs:622-629
    std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> pos_and_values_u128() const {
        const auto vals = values_u128();
        std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> v(vals.size());
        for (size_t i = 0; i < v.size(); i++) v[i] = {(*pos_)[i], vals[i]};
        return v;
    }
    const std::vector<uint32_t>& positions() const { return *pos_; }

private:
    uint32_t len_;
    mz_params p_;
    PackedSeq seq_;
    const std::vector<uint32_t>* pos_;
};

class Builder {
public:
    Builder(uint32_t k, uint32_t w, bool canonical, uint32_t syncmer) : k_(k), w_(w), canonical_(canonical), syncmer_(syncmer) {}
    Builder hasher(const Hasher& h) const {
        if (sk_) throw std::logic_error("hasher() must be called before super_kmers()");  // src/lib.rs:323-338
        Builder b = *this;
        b.hasher_ = h;
        b.has_hasher_ = true;
        return b;
    }
    Builder super_kmers(std::vector<uint32_t>& sk) const {
        if (syncmer_) throw std::logic_error("super_kmers() is only available for minimizers");  // src/lib.rs:339
        Builder b = *this;
        b.sk_ = &sk;
        return b;
    }
    // Appends to min_pos (and to the super-k-mer vector), src/lib.rs:80-81.
    Output run(const PackedSeq& seq, std::vector<uint32_t>& min_pos) const {
        return run_impl(seq, nullptr, 0, min_pos, /*scalar=*/false);
    }
    // run_scalar (src/lib.rs:372-378, 508-533): same result; the scalar collectors overwrite the
    // vectors from index 0 (src/collect.rs:15-76) instead of appending.
    Output run_scalar(const PackedSeq& seq, std::vector<uint32_t>& min_pos) const {
        return run_impl(seq, nullptr, 0, min_pos, /*scalar=*/true);
    }
    std::vector<uint32_t> run_scalar_once(const PackedSeq& seq) const {
        std::vector<uint32_t> v;
        run_scalar(seq, v);
        return v;
    }
    // src/lib.rs:451-496: windows holding an ambiguous base produce nothing.  Canonical builders
    // without super-k-mers only (a type-state restriction in the reference, checked here).
    Output run_skip_ambiguous_windows(const PackedNSeq& nseq, std::vector<uint32_t>& min_pos) const {
        if (!canonical_ || sk_) throw std::logic_error("run_skip_ambiguous_windows: canonical builder without super_kmers() only");
        if (!nseq.ambiguous) throw std::invalid_argument("PackedNSeq without an ambiguity mask");
        return run_impl(nseq.seq, nseq.ambiguous, nseq.amb_offset, min_pos, false);
    }
    std::vector<uint32_t> run_skip_ambiguous_windows_once(const PackedNSeq& nseq) const {
        std::vector<uint32_t> v;
        run_skip_ambiguous_windows(nseq, v);
        return v;
    }
    std::vector<uint32_t> run_once(const PackedSeq& seq) const {
        std::vector<uint32_t> v;
        run(seq, v);
        return v;
    }
    // Not in the reference crate: super-k-mers (bench/src/minimizer.rs:3-36) sharded by their
    // minimizer on the device; only the two histograms leave the GPU (mz_run_bucket_stats).
    void bucket_stats(const PackedSeq& seq, uint32_t n_buckets, std::vector<uint64_t>& superkmers,
                      std::vector<uint64_t>& windows, uint64_t* n_minimizers = nullptr) const {
        if (syncmer_) throw std::logic_error("bucket_stats() is only available for minimizers");
        const mz_params p = params();
        superkmers.assign(n_buckets, 0);
        windows.assign(n_buckets, 0);
        detail::check(mz_run_bucket_stats(detail::thread_ctx(), &p, seq.data, seq.offset, seq.len, n_buckets,
                                          superkmers.data(), windows.data(), n_minimizers));
    }

private:
    mz_params params() const {
        mz_params p;
        detail::check(mz_params_nthash(&p, k_, w_, syncmer_, canonical_));
        if (has_hasher_) {
            if (hasher_.k != k_) throw std::invalid_argument("hasher.k() must equal k");
            detail::check(hasher_.kind == Hasher::Nt    ? mz_params_set_nthash(&p, hasher_.canonical)
                          : hasher_.kind == Hasher::Mul ? mz_params_set_mulhash(&p, hasher_.canonical)
                                                        : mz_params_set_tables(&p, hasher_.f, hasher_.c, hasher_.rot, hasher_.canonical));
        }
        p.want_sk = sk_ != nullptr;
        p.value_bits = 0;  // positions only; values are lazy (Output)
        return p;
    }
    Output run_impl(const PackedSeq& seq, const uint8_t* amb, uint64_t amb_off, std::vector<uint32_t>& min_pos, bool scalar) const {
        const mz_params p = params();
        const uint32_t len = syncmer_ ? k_ + w_ - 1 : k_;
        detail::check(mz_params_validate(&p, seq.len));
        const uint64_t l = k_ + w_ - 1, nwin = seq.len >= l ? seq.len - l + 1 : 0;
        uint64_t cap = (uint64_t)(nwin * 2.5 / (w_ + 1.0)) + 4096;
        std::vector<uint32_t> pos, sk;
        for (;;) {
            pos.resize(cap);
            if (sk_) sk.resize(cap);
            mz_out out{pos.data(), sk_ ? sk.data() : nullptr, nullptr, cap, 0};
            int rc = amb ? mz_run_skip_ambiguous(detail::thread_ctx(), &p, seq.data, seq.offset, seq.len, amb, amb_off, &out)
                         : mz_run(detail::thread_ctx(), &p, seq.data, seq.offset, seq.len, &out);
            if (rc == MZ_ERR_CAPACITY) {
                cap = out.count;
                continue;
            }
            detail::check(rc);
            pos.resize(out.count);
            if (sk_) sk.resize(out.count);
            break;
        }
        if (scalar) {  // overwrite; an empty window stream leaves the index vector untouched (src/collect.rs:45-48)
            min_pos = pos;
            if (sk_ && nwin) *sk_ = sk;
        } else {
            // SIMD-collector quirk (src/collect.rs:257,267)
            size_t skip = (!syncmer_ && !pos.empty() && !min_pos.empty() && pos[0] == min_pos.back()) ? 1 : 0;
            min_pos.insert(min_pos.end(), pos.begin() + skip, pos.end());
            if (sk_) sk_->insert(sk_->end(), sk.begin() + skip, sk.end());
        }
        mz_params pv = p;
        pv.want_sk = 0;
        return Output(len, pv, seq, &min_pos);
    }

    uint32_t k_, w_;
    bool canonical_;
    uint32_t syncmer_;
    Hasher hasher_{};
    bool has_hasher_ = false;
    std::vector<uint32_t>* sk_ = nullptr;
};

inline Builder minimizers(uint32_t k, uint32_t w) { return {k, w, false, MZ_MODE_MINIMIZER}; }                     // src/lib.rs:240
inline Builder canonical_minimizers(uint32_t k, uint32_t w) { return {k, w, true, MZ_MODE_MINIMIZER}; }            // :250
inline Builder closed_syncmers(uint32_t k, uint32_t w) { return {k, w, false, MZ_MODE_CLOSED_SYNCMER}; }           // :269
inline Builder canonical_closed_syncmers(uint32_t k, uint32_t w) { return {k, w, true, MZ_MODE_CLOSED_SYNCMER}; }  // :282
inline Builder open_syncmers(uint32_t k, uint32_t w) { return {k, w, false, MZ_MODE_OPEN_SYNCMER}; }               // :301
inline Builder canonical_open_syncmers(uint32_t k, uint32_t w) { return {k, w, true, MZ_MODE_OPEN_SYNCMER}; }      // :311
inline Builder canonical_syncmers(uint32_t k, uint32_t w) { return canonical_closed_syncmers(k, w); }              // README.md:65
inline std::vector<uint32_t> minimizer_positions(const PackedSeq& s, uint32_t k, uint32_t w) { return minimizers(k, w).run_once(s); }                      // :639
inline std::vector<uint32_t> canonical_minimizer_positions(const PackedSeq& s, uint32_t k, uint32_t w) { return canonical_minimizers(k, w).run_once(s); }  // :652

}  // namespace simd_minimizers
