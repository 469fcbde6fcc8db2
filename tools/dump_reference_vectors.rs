//! Golden-vector generator for the B200 path's parity tests.
//!
//! This file is NOT built in this repository (the image has no Rust toolchain).  Someone with
//! `cargo` drops it into the reference crate and runs it:
//!
//!     cp tools/dump_reference_vectors.rs  <simd-minimizers v3.0.0>/examples/
//!     cd <simd-minimizers v3.0.0>
//!     cargo run --release --example dump_reference_vectors > reference_dump.json
//!     cp reference_dump.json  <this repo>/tests/golden/reference_dump.json
//!
//! tests/test_reference_dump.py then checks the CPU oracle AND the CUDA path against every
//! vector in it, and derives the per-base tables of every hasher from the k = 1 / k = 2 hashes,
//! which pins what the reference's own k = 5 doc vectors cannot: NtHasher at k = 31 (rotation
//! wrap-around), MulHasher, seeded hashers.  It only uses API that the reference's own tests
//! use (src/test.rs, src/lib.rs doc examples) and prints JSON by hand (no extra dependencies).

use simd_minimizers::packed_seq::{PackedNSeqVec, PackedSeqVec, SeqVec};
use simd_minimizers::seq_hash::{AntiLexHasher, KmerHasher, MulHasher, NtHasher};
use simd_minimizers::{
    canonical_closed_syncmers, canonical_minimizers, canonical_open_syncmers, closed_syncmers,
    minimizers, open_syncmers,
};

fn arr<T: std::fmt::Display>(v: &[T]) -> String {
    let items: Vec<String> = v.iter().map(|x| x.to_string()).collect();
    format!("[{}]", items.join(","))
}

/// Deterministic ASCII DNA: 64-bit LCG (Knuth's MMIX constants), base = "ACGT"[(x >> 33) & 3].
/// tests/test_reference_dump.py regenerates the same text and also reads it from the dump.
fn lcg_dna(n: usize, mut x: u64) -> Vec<u8> {
    let mut out = Vec::with_capacity(n);
    for _ in 0..n {
        x = x
            .wrapping_mul(6364136223846793005)
            .wrapping_add(1442695040888963407);
        out.push(b"ACGT"[((x >> 33) & 3) as usize]);
    }
    out
}

/// Raw 32-bit k-mer hashes of the first `count` k-mers of `ascii`.
fn hashes<H: KmerHasher>(h: &H, ascii: &[u8], count: usize) -> Vec<u32> {
    let k = h.k();
    let packed = PackedSeqVec::from_ascii(&ascii[..count + k - 1]);
    h.hash_kmers_scalar(packed.as_slice()).collect()
}

fn dump_hashes<H: KmerHasher>(out: &mut Vec<String>, name: &str, rc: bool, seed: Option<u32>, ascii: &[u8], mk: impl Fn(usize) -> H) {
    // k = 1 gives the per-base table, k = 2 the rotation, the rest pin the wrap-around
    for k in [1usize, 2, 3, 5, 8, 16, 31, 32, 33, 47] {
        let h = mk(k);
        // "ACTG" first: codes 0, 1, 2, 3 in order
        let mut text = b"ACTGACTGGTCA".to_vec();
        text.extend_from_slice(ascii);
        let hs = hashes(&h, &text, 64);
        out.push(format!(
            "{{\"hasher\":\"{}\",\"hash_canonical\":{},\"seed\":{},\"k\":{},\"text\":\"{}\",\"hashes\":{}}}",
            name,
            rc,
            seed.map(|s| s.to_string()).unwrap_or("null".to_string()),
            k,
            String::from_utf8_lossy(&text[..64 + k - 1]),
            arr(&hs)
        ));
    }
}

#[allow(clippy::too_many_arguments)]
fn case_json(name: &str, hasher: &str, hash_rc: bool, seed: Option<u32>, builder_rc: bool, mode: u32, k: usize, w: usize,
             seq_name: &str, pos: &[u32], sk: Option<&[u32]>, v64: Option<&[u64]>, v128: Option<&[u128]>) -> String {
    format!(
        "{{\"name\":\"{}\",\"hasher\":\"{}\",\"hash_canonical\":{},\"seed\":{},\"builder_canonical\":{},\"mode\":{},\"k\":{},\"w\":{},\"seq\":\"{}\",\"pos\":{},\"sk\":{},\"values_u64\":{},\"values_u128\":{}}}",
        name, hasher, hash_rc, seed.map(|s| s.to_string()).unwrap_or("null".to_string()), builder_rc, mode, k, w, seq_name,
        arr(pos),
        sk.map(arr).unwrap_or("null".to_string()),
        v64.map(arr).unwrap_or("null".to_string()),
        // u128 as decimal strings (JSON numbers that large are not portable)
        v128.map(|v| format!("[{}]", v.iter().map(|x| format!("\"{}\"", x)).collect::<Vec<_>>().join(","))).unwrap_or("null".to_string()),
    )
}

/// minimizers (+ super-k-mer starts, + values) for one hasher, forward and canonical builder
fn dump_minimizers<HF: KmerHasher, HC: KmerHasher>(out: &mut Vec<String>, hname: &str, seed: Option<u32>, k: usize, w: usize,
                                                    ascii: &[u8], fwd: &HF, can: &HC) {
    let packed = PackedSeqVec::from_ascii(ascii);
    let seq = packed.as_slice();
    // forward builder, forward hasher
    {
        let (mut pos, mut sk) = (vec![], vec![]);
        let vals: Vec<u64> = minimizers(k, w).hasher(fwd).super_kmers(&mut sk).run(seq, &mut pos).values_u64().collect();
        out.push(case_json(&format!("{hname}_fwd_k{k}_w{w}"), hname, false, seed, false, 0, k, w, "seq", &pos, Some(&sk), Some(&vals), None));
    }
    // forward builder, canonical hasher (src/minimizers.rs:69-71)
    {
        let mut pos = vec![];
        minimizers(k, w).hasher(can).run(seq, &mut pos);
        out.push(case_json(&format!("{hname}_fwdbuilder_canhash_k{k}_w{w}"), hname, true, seed, false, 0, k, w, "seq", &pos, None, None, None));
    }
    // canonical builder
    if (k + w - 1) % 2 == 1 {
        let (mut pos, mut sk) = (vec![], vec![]);
        let vals: Vec<u64> = canonical_minimizers(k, w).hasher(can).super_kmers(&mut sk).run(seq, &mut pos).values_u64().collect();
        out.push(case_json(&format!("{hname}_can_k{k}_w{w}"), hname, true, seed, true, 0, k, w, "seq", &pos, Some(&sk), Some(&vals), None));
        // scalar path must agree (the reference asserts this in its own tests)
        let mut pos2 = vec![];
        canonical_minimizers(k, w).hasher(can).run_scalar(seq, &mut pos2);
        assert_eq!(pos, pos2);
    }
}

fn main() {
    let ascii = lcg_dna(4096, 42);
    // the same text with runs of N (and single Ns) for run_skip_ambiguous_windows
    let mut nascii = ascii.clone();
    for (start, len) in [(100usize, 1usize), (300, 3), (700, 60), (1500, 1), (1501, 1), (2000, 250), (3900, 20), (4090, 6)] {
        for c in &mut nascii[start..start + len] {
            *c = b'N';
        }
    }

    let mut hs: Vec<String> = vec![];
    dump_hashes(&mut hs, "nt", false, None, &ascii, |k| NtHasher::<false>::new(k));
    dump_hashes(&mut hs, "nt", true, None, &ascii, |k| NtHasher::<true>::new(k));
    dump_hashes(&mut hs, "mul", false, None, &ascii, |k| MulHasher::<false>::new(k));
    dump_hashes(&mut hs, "mul", true, None, &ascii, |k| MulHasher::<true>::new(k));
    dump_hashes(&mut hs, "antilex", false, None, &ascii, |k| AntiLexHasher::<false>::new(k));
    dump_hashes(&mut hs, "antilex", true, None, &ascii, |k| AntiLexHasher::<true>::new(k));
    dump_hashes(&mut hs, "nt", false, Some(1234), &ascii, |k| NtHasher::<false>::new_with_seed(k, 1234));
    dump_hashes(&mut hs, "nt", true, Some(1234), &ascii, |k| NtHasher::<true>::new_with_seed(k, 1234));
    dump_hashes(&mut hs, "mul", false, Some(1234), &ascii, |k| MulHasher::<false>::new_with_seed(k, 1234));
    dump_hashes(&mut hs, "mul", true, Some(1234), &ascii, |k| MulHasher::<true>::new_with_seed(k, 1234));

    let mut cs: Vec<String> = vec![];
    for (k, w) in [(31usize, 19usize), (21, 11), (5, 7), (32, 2), (13, 33), (8, 101)] {
        dump_minimizers(&mut cs, "nt", None, k, w, &ascii, &NtHasher::<false>::new(k), &NtHasher::<true>::new(k));
        dump_minimizers(&mut cs, "mul", None, k, w, &ascii, &MulHasher::<false>::new(k), &MulHasher::<true>::new(k));
    }
    dump_minimizers(&mut cs, "antilex", None, 31, 19, &ascii, &AntiLexHasher::<false>::new(31), &AntiLexHasher::<true>::new(31));
    dump_minimizers(&mut cs, "nt", Some(1234), 31, 19, &ascii, &NtHasher::<false>::new_with_seed(31, 1234), &NtHasher::<true>::new_with_seed(31, 1234));
    dump_minimizers(&mut cs, "mul", Some(1234), 31, 19, &ascii, &MulHasher::<false>::new_with_seed(31, 1234), &MulHasher::<true>::new_with_seed(31, 1234));

    // syncmers: k = 31, w = 11 (l = 41: u128 values) and a short one with u64 values
    let packed = PackedSeqVec::from_ascii(&ascii);
    let seq = packed.as_slice();
    for (k, w) in [(31usize, 11usize), (9, 5)] {
        let l = k + w - 1;
        let mut pos = vec![];
        let out = canonical_closed_syncmers(k, w).run(seq, &mut pos);
        let (v64, v128): (Option<Vec<u64>>, Option<Vec<u128>>) = if l <= 32 { (Some(out.values_u64().collect()), None) } else { (None, Some(out.values_u128().collect())) };
        cs.push(case_json(&format!("nt_can_closed_k{k}_w{w}"), "nt", true, None, true, 1, k, w, "seq", &pos, None, v64.as_deref(), v128.as_deref()));
        let mut pos = vec![];
        let out = canonical_open_syncmers(k, w).run(seq, &mut pos);
        let (v64, v128): (Option<Vec<u64>>, Option<Vec<u128>>) = if l <= 32 { (Some(out.values_u64().collect()), None) } else { (None, Some(out.values_u128().collect())) };
        cs.push(case_json(&format!("nt_can_open_k{k}_w{w}"), "nt", true, None, true, 2, k, w, "seq", &pos, None, v64.as_deref(), v128.as_deref()));
        let mut pos = vec![];
        closed_syncmers(k, w).run(seq, &mut pos);
        cs.push(case_json(&format!("nt_fwd_closed_k{k}_w{w}"), "nt", false, None, false, 1, k, w, "seq", &pos, None, None, None));
        let mut pos = vec![];
        open_syncmers(k, w).run(seq, &mut pos);
        cs.push(case_json(&format!("nt_fwd_open_k{k}_w{w}"), "nt", false, None, false, 2, k, w, "seq", &pos, None, None, None));
    }

    // run_skip_ambiguous_windows on a PackedNSeq (src/lib.rs:451-496)
    let nseq = PackedNSeqVec::from_ascii(&nascii);
    for (k, w) in [(31usize, 19usize), (21, 11), (5, 7)] {
        let pos = canonical_minimizers(k, w).run_skip_ambiguous_windows_once(nseq.as_slice());
        cs.push(case_json(&format!("nt_can_skipamb_k{k}_w{w}"), "nt", true, None, true, 0, k, w, "nseq", &pos, None, None, None));
        let pos = canonical_closed_syncmers(k, w).run_skip_ambiguous_windows_once(nseq.as_slice());
        cs.push(case_json(&format!("nt_can_closed_skipamb_k{k}_w{w}"), "nt", true, None, true, 1, k, w, "nseq", &pos, None, None, None));
    }

    println!("{{");
    println!("\"generator\":\"tools/dump_reference_vectors.rs\",");
    println!("\"crate\":\"simd-minimizers {}\",", env!("CARGO_PKG_VERSION"));
    println!("\"seq\":\"{}\",", String::from_utf8_lossy(&ascii));
    println!("\"nseq\":\"{}\",", String::from_utf8_lossy(&nascii));
    println!("\"hashes\":[\n{}\n],", hs.join(",\n"));
    println!("\"cases\":[\n{}\n]", cs.join(",\n"));
    println!("}}");
}
