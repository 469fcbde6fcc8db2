//! Safe shim with the reference crate's builder surface (simd-minimizers v3.0.0,
//! `src/lib.rs:225-654`): `minimizer_positions`, `canonical_minimizer_positions`,
//! `minimizers/canonical_minimizers/closed_syncmers/.../canonical_open_syncmers(k, w)`,
//! `.hasher(..)`, `.super_kmers(..)`, `.run(..)`, `.run_once(..)`, `.run_scalar(..)`,
//! `.run_scalar_once(..)`, `.run_skip_ambiguous_windows(..)`, and the lazy
//! `Output::{values_u64, values_u128, pos_and_values_u64, pos_and_values_u128}`.
//! Bodies call the C ABI; nothing is computed on the CPU.  Not compiled in this repository's
//! image (no Rust toolchain) -- see INTEGRATION.md; `include/simd_minimizers.hpp` is the same shim
//! in C++ and IS compiled and run by the tests (`tests/cpp_mirror_test.cpp`).
use mzb200_sys as sys;
use seq_hash::packed_seq::{PackedNSeq, PackedSeq, Seq};
use std::cell::RefCell;

/// Hashers the device path understands: anything that is a per-base table with a fixed rotation.
/// Sealed on purpose: an arbitrary `KmerHasher` cannot cross the ABI.
pub trait TableHasher: sealed::Sealed {
    fn k(&self) -> usize;
    fn is_canonical(&self) -> bool;
    /// Fill `f`, `c`, `rot`, `hash_canonical`.
    fn fill(&self, p: &mut sys::mz_params);
}
mod sealed {
    pub trait Sealed {}
    impl<const RC: bool> Sealed for seq_hash::NtHasher<RC> {}
    impl<const RC: bool> Sealed for seq_hash::MulHasher<RC> {}
    impl Sealed for super::Tables {}
}
impl<const RC: bool> TableHasher for seq_hash::NtHasher<RC> {
    fn k(&self) -> usize { seq_hash::KmerHasher::k(self) }
    fn is_canonical(&self) -> bool { RC }
    fn fill(&self, p: &mut sys::mz_params) { unsafe { sys::mz_params_set_nthash(p, RC as u32); } }
}
impl<const RC: bool> TableHasher for seq_hash::MulHasher<RC> {
    fn k(&self) -> usize { seq_hash::KmerHasher::k(self) }
    fn is_canonical(&self) -> bool { RC }
    fn fill(&self, p: &mut sys::mz_params) { unsafe { sys::mz_params_set_mulhash(p, RC as u32); } }
}
/// Any other table hasher, e.g. the seeded ones (`H::new_with_seed(k, seed)`, src/lib.rs:157,
/// src/test.rs:287): seq-hash derives the per-base constants from the seed; the maintainer reads
/// them out of the hasher object (`f[b]` = forward constant of base code b, `c[b]` = constant of its
/// complement, A=0 C=1 T=2 G=3) and passes them here.  `mz_params_set_tables` is the ABI call.
pub struct Tables {
    pub k: usize,
    pub f: [u32; 4],
    pub c: [u32; 4],
    pub rot: u32,
    pub canonical: bool,
}
impl TableHasher for Tables {
    fn k(&self) -> usize { self.k }
    fn is_canonical(&self) -> bool { self.canonical }
    fn fill(&self, p: &mut sys::mz_params) {
        unsafe { sys::mz_params_set_tables(p, self.f.as_ptr(), self.c.as_ptr(), self.rot, self.canonical as u32); }
    }
}

struct Ctx(*mut sys::mz_ctx);
impl Drop for Ctx {
    fn drop(&mut self) { unsafe { sys::mz_ctx_destroy(self.0) } }
}
thread_local! {
    // the reference keeps its scratch thread_local too (src/lib.rs:217-219)
    static CTX: RefCell<Option<Ctx>> = const { RefCell::new(None) };
}
fn with_ctx<R>(f: impl FnOnce(*mut sys::mz_ctx) -> R) -> R {
    CTX.with_borrow_mut(|c| {
        if c.is_none() {
            let mut h = std::ptr::null_mut();
            let rc = unsafe { sys::mz_ctx_create(std::ptr::null(), 0, &mut h) };
            assert!(rc == sys::MZ_OK, "mzb200: {}", err(rc));
            *c = Some(Ctx(h));
        }
        f(c.as_ref().unwrap().0)
    })
}
fn err(rc: i32) -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::mz_strerror(rc)).to_string_lossy().into_owned() }
}

pub struct Builder<'h, const CANONICAL: bool, SkPos, const SYNCMER: u8> {
    k: usize,
    w: usize,
    hasher: Option<&'h dyn TableHasher>,
    sk_pos: SkPos,
}

/// src/lib.rs:232-237, 579-630.  Values are lazy, exactly as in the reference: `run` moves
/// positions only; the iterators compute the k-mers (l-mers for syncmers) of ALL of `min_pos` --
/// entries of earlier runs included, the reference iterates the caller's whole vector
/// (src/lib.rs:598-612) -- with one `mz_values` call (the sequence of the run is still resident on
/// the device and is not uploaded again).
pub struct Output<'o, 's, const CANONICAL: bool> {
    /// k for minimizers, k + w - 1 for syncmers (src/lib.rs:233)
    len: usize,
    params: sys::mz_params,
    seq: PackedSeq<'s>,
    min_pos: &'o Vec<u32>,
}
impl<'o, 's, const CANONICAL: bool> Output<'o, 's, CANONICAL> {
    fn fetch(&self, bits: u32) -> Vec<u64> {
        let words = (bits / 64) as usize;
        let n = self.min_pos.len();
        let mut v = vec![0u64; n * words];
        let (bytes, offset) = self.seq.as_packed_bytes();
        let rc = with_ctx(|c| unsafe {
            sys::mz_values(c, &self.params, bytes.as_ptr(), offset as u64, self.seq.len() as u64,
                           self.min_pos.as_ptr(), n as u64, bits, v.as_mut_ptr())
        });
        // read_kmer asserts on a position whose k-mer runs past the sequence (MZ_ERR_BAD_ARG)
        assert!(rc == sys::MZ_OK, "mzb200: {}", err(rc));
        v
    }
    /// src/lib.rs:584-590, 598-604
    pub fn values_u64(&self) -> impl ExactSizeIterator<Item = u64> {
        assert!(self.len <= 32, "values_u64: k-mer / l-mer longer than 32 bases");
        self.fetch(64).into_iter()
    }
    /// src/lib.rs:591-597, 615-621
    pub fn values_u128(&self) -> impl ExactSizeIterator<Item = u128> {
        assert!(self.len <= 64, "values_u128: k-mer / l-mer longer than 64 bases");
        let raw = self.fetch(128);
        (0..raw.len() / 2).map(move |i| (raw[2 * i] as u128) | ((raw[2 * i + 1] as u128) << 64))
    }
    /// src/lib.rs:605-612
    pub fn pos_and_values_u64(&self) -> impl ExactSizeIterator<Item = (u32, u64)> + '_ {
        self.min_pos.iter().copied().zip(self.values_u64())
    }
    /// src/lib.rs:622-629
    pub fn pos_and_values_u128(&self) -> impl ExactSizeIterator<Item = (u32, u128)> + '_ {
        self.min_pos.iter().copied().zip(self.values_u128())
    }
}

macro_rules! ctor {
    ($name:ident, $canon:literal, $sync:literal) => {
        #[must_use]
        pub const fn $name(k: usize, w: usize) -> Builder<'static, $canon, (), $sync> {
            Builder { k, w, hasher: None, sk_pos: () }
        }
    };
}
ctor!(minimizers, false, 0);
ctor!(canonical_minimizers, true, 0);
ctor!(closed_syncmers, false, 1);
ctor!(canonical_closed_syncmers, true, 1);
ctor!(open_syncmers, false, 2);
ctor!(canonical_open_syncmers, true, 2);
/// README.md:65 name; alias of the closed variant.
pub const fn canonical_syncmers(k: usize, w: usize) -> Builder<'static, true, (), 1> { canonical_closed_syncmers(k, w) }

/// How a run hands its results to the caller's vectors.
#[derive(Clone, Copy, PartialEq)]
enum Collect {
    /// SIMD path: append (src/lib.rs:80-81), dropping the first new position when it repeats the
    /// caller's last one (src/collect.rs:257,267).
    Append,
    /// Scalar path: the scalar collectors overwrite from index 0 (src/collect.rs:15-76); an empty
    /// window stream leaves the super-k-mer vector untouched (src/collect.rs:45-48).
    Overwrite,
}

impl<'h, const CANONICAL: bool, const SYNCMER: u8> Builder<'h, CANONICAL, (), SYNCMER> {
    #[must_use]
    pub fn hasher<'h2>(&self, h: &'h2 dyn TableHasher) -> Builder<'h2, CANONICAL, (), SYNCMER> {
        Builder { k: self.k, w: self.w, hasher: Some(h), sk_pos: () }
    }
    pub fn run<'o, 's>(&self, seq: PackedSeq<'s>, min_pos: &'o mut Vec<u32>) -> Output<'o, 's, CANONICAL> {
        run_impl::<CANONICAL, SYNCMER>(self.k, self.w, self.hasher, seq, None, min_pos, None, Collect::Append)
    }
    pub fn run_once(&self, seq: PackedSeq<'_>) -> Vec<u32> {
        let mut v = vec![];
        self.run(seq, &mut v);
        v
    }
    /// src/lib.rs:358-384, 504-551.  The reference asserts scalar == SIMD (src/test.rs:55-110);
    /// here both are the device path, only the collection semantics differ.
    pub fn run_scalar<'o, 's>(&self, seq: PackedSeq<'s>, min_pos: &'o mut Vec<u32>) -> Output<'o, 's, CANONICAL> {
        run_impl::<CANONICAL, SYNCMER>(self.k, self.w, self.hasher, seq, None, min_pos, None, Collect::Overwrite)
    }
    pub fn run_scalar_once(&self, seq: PackedSeq<'_>) -> Vec<u32> {
        let mut v = vec![];
        self.run_scalar(seq, &mut v);
        v
    }
}
/// src/lib.rs:451-496: canonical builders without super-k-mers only.
impl<'h, const SYNCMER: u8> Builder<'h, true, (), SYNCMER> {
    pub fn run_skip_ambiguous_windows<'o, 's>(&self, nseq: PackedNSeq<'s>, min_pos: &'o mut Vec<u32>) -> Output<'o, 's, true> {
        // BitSeq storage + bit offset of base 0 (packed-seq 5.0.0; accessor names to be confirmed
        // against the crate -- it is not vendored in the reference tree)
        let (amb, amb_off) = nseq.ambiguous.as_bit_bytes();
        run_impl::<true, SYNCMER>(self.k, self.w, self.hasher, nseq.seq, Some((amb, amb_off)), min_pos, None, Collect::Append)
    }
    pub fn run_skip_ambiguous_windows_once(&self, nseq: PackedNSeq<'_>) -> Vec<u32> {
        let mut v = vec![];
        self.run_skip_ambiguous_windows(nseq, &mut v);
        v
    }
}
impl<'h, const CANONICAL: bool> Builder<'h, CANONICAL, (), 0> {
    #[must_use]
    pub fn super_kmers<'o2>(&self, sk_pos: &'o2 mut Vec<u32>) -> Builder<'h, CANONICAL, &'o2 mut Vec<u32>, 0> {
        Builder { k: self.k, w: self.w, hasher: self.hasher, sk_pos }
    }
    /// Not in the reference crate: super-k-mers (bench/src/minimizer.rs:3-36) sharded by their
    /// minimizer on the device; only the two histograms leave the GPU (`mz_run_bucket_stats`).
    /// Returns (super-k-mers per bucket, windows per bucket, number of minimizers).
    pub fn bucket_stats(&self, seq: PackedSeq<'_>, n_buckets: u32) -> (Vec<u64>, Vec<u64>, u64) {
        let p = params::<CANONICAL, 0>(self.k, self.w, self.hasher, true);
        let (bytes, offset) = seq.as_packed_bytes();
        let (mut sk, mut win, mut n) = (vec![0u64; n_buckets as usize], vec![0u64; n_buckets as usize], 0u64);
        let rc = with_ctx(|c| unsafe {
            sys::mz_run_bucket_stats(c, &p, bytes.as_ptr(), offset as u64, seq.len() as u64, n_buckets,
                                     sk.as_mut_ptr(), win.as_mut_ptr(), &mut n)
        });
        assert!(rc == sys::MZ_OK, "mzb200: {}", err(rc));
        (sk, win, n)
    }
}
impl<'h, 'o2, const CANONICAL: bool> Builder<'h, CANONICAL, &'o2 mut Vec<u32>, 0> {
    pub fn run<'o, 's>(self, seq: PackedSeq<'s>, min_pos: &'o mut Vec<u32>) -> Output<'o, 's, CANONICAL> {
        run_impl::<CANONICAL, 0>(self.k, self.w, self.hasher, seq, None, min_pos, Some(self.sk_pos), Collect::Append)
    }
    pub fn run_scalar<'o, 's>(self, seq: PackedSeq<'s>, min_pos: &'o mut Vec<u32>) -> Output<'o, 's, CANONICAL> {
        run_impl::<CANONICAL, 0>(self.k, self.w, self.hasher, seq, None, min_pos, Some(self.sk_pos), Collect::Overwrite)
    }
}

fn params<const CANONICAL: bool, const SYNCMER: u8>(k: usize, w: usize, hasher: Option<&dyn TableHasher>, want_sk: bool) -> sys::mz_params {
    let mut p = sys::mz_params::default();
    unsafe { sys::mz_params_nthash(&mut p, k as u32, w as u32, SYNCMER as u32, CANONICAL as u32) };
    if let Some(h) = hasher {
        assert!(h.k() == k);
        h.fill(&mut p);
    }
    p.want_sk = want_sk as u32;
    p.value_bits = 0; // positions only; values are lazy (Output)
    p
}

/// `Builder::run_impl` / `run_with_buf` (src/lib.rs:386-448, 554-576) -> one `mz_run` call.
fn run_impl<'o, 's, const CANONICAL: bool, const SYNCMER: u8>(
    k: usize, w: usize, hasher: Option<&dyn TableHasher>, seq: PackedSeq<'s>,
    ambiguous: Option<(&[u8], usize)>, min_pos: &'o mut Vec<u32>, mut sk_pos: Option<&mut Vec<u32>>,
    collect: Collect,
) -> Output<'o, 's, CANONICAL> {
    let p = params::<CANONICAL, SYNCMER>(k, w, hasher, sk_pos.is_some());
    let n = seq.len();
    let rc = unsafe { sys::mz_params_validate(&p, n as u64) };
    // the reference panics on these (assert!); keep its messages
    assert!(rc == sys::MZ_OK, "{}", err(rc));
    let nwin = (n + 1).saturating_sub(k + w - 1);
    if collect == Collect::Overwrite {
        min_pos.clear();
        if nwin > 0 {
            if let Some(sk) = sk_pos.as_deref_mut() { sk.clear(); }
        }
    }
    let mut cap = (nwin as f64 * 2.5 / (w as f64 + 1.0)) as usize + 4096;
    let (bytes, offset) = seq.as_packed_bytes(); // (&[u8], bases into the first byte)
    loop {
        let start = min_pos.len();
        min_pos.reserve(cap);
        if let Some(sk) = sk_pos.as_deref_mut() { sk.reserve(cap); }
        let mut out = sys::mz_out {
            pos: min_pos.spare_capacity_mut().as_mut_ptr().cast(),
            sk: sk_pos.as_deref_mut().map_or(std::ptr::null_mut(), |s| s.spare_capacity_mut().as_mut_ptr().cast()),
            val: std::ptr::null_mut(),
            capacity: cap as u64,
            count: 0,
        };
        let rc = with_ctx(|c| unsafe {
            match ambiguous {
                None => sys::mz_run(c, &p, bytes.as_ptr(), offset as u64, n as u64, &mut out),
                Some((amb, amb_off)) => sys::mz_run_skip_ambiguous(
                    c, &p, bytes.as_ptr(), offset as u64, n as u64, amb.as_ptr(), amb_off as u64, &mut out),
            }
        });
        if rc == sys::MZ_ERR_CAPACITY { cap = out.count as usize; continue; }
        assert!(rc == sys::MZ_OK, "mzb200: {}", err(rc));
        let m = out.count as usize;
        // SIMD-collector quirk (src/collect.rs:257,267): drop the first new position if it repeats
        // the caller's last one.
        let skip = (collect == Collect::Append && SYNCMER == 0 && m > 0 && start > 0
            && unsafe { *min_pos.as_ptr().add(start) } == min_pos[start - 1]) as usize;
        unsafe {
            if skip == 1 { std::ptr::copy(min_pos.as_ptr().add(start + 1), min_pos.as_mut_ptr().add(start), m - 1); }
            min_pos.set_len(start + m - skip);
            if let Some(sk) = sk_pos.as_deref_mut() {
                let s0 = sk.len();
                if skip == 1 { std::ptr::copy(sk.as_ptr().add(s0 + 1), sk.as_mut_ptr().add(s0), m - 1); }
                sk.set_len(s0 + m - skip);
            }
        }
        let mut pv = p;
        pv.want_sk = 0;
        let len = if SYNCMER != 0 { k + w - 1 } else { k };
        return Output { len, params: pv, seq, min_pos };
    }
}

pub fn minimizer_positions(seq: PackedSeq<'_>, k: usize, w: usize) -> Vec<u32> { minimizers(k, w).run_once(seq) }
pub fn canonical_minimizer_positions(seq: PackedSeq<'_>, k: usize, w: usize) -> Vec<u32> { canonical_minimizers(k, w).run_once(seq) }
