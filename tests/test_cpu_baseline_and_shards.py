"""CPU tests: (1) the AVX2 8-lane CPU baseline used by bench.py equals the scalar oracle;
(2) the N>1 path -- contiguous window shards with a (k+w-2)-base halo and one seam window,
ordered concatenation -- exercised with world_size=2 on the gloo backend, using the oracle's
run_range as the per-rank compute stand-in (the GPU kernel obeys the same shard rule, which
tests/test_gpu_parity.py::test_device_resident_and_window_ranges checks on the device)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_avx2_baseline_matches_oracle(oracle):
    rng = np.random.default_rng(12)
    for _ in range(60):
        n = int(rng.integers(0, 40000))
        k, w, off = int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(0, 4))
        seq = oracle.synth_packed(int(rng.integers(0, 1 << 40)), n + off + 64)
        for canon in (False, True):
            if canon and (k + w - 1) % 2 == 0:
                continue
            for kind in ("nt", "mul"):
                pr = oracle.make_params(k, w, canonical=canon, hasher=oracle.make_hasher(kind, canon))
                epos, esk = oracle.run(seq, off, n, pr, want_sk=True)
                pos, sk, val = oracle.baseline_run_mt(seq, off, n, pr, int(rng.integers(1, 5)),
                                                      want_sk=True, want_val=k <= 32)
                assert np.array_equal(pos, epos) and np.array_equal(sk, esk), (n, k, w, off, canon, kind)
                if k <= 32:
                    assert np.array_equal(val, oracle.values_u64(seq, off, k, canon, epos))


def test_avx2_baseline_position_rebase(oracle):
    """> 65535 k-mers per lane exercises the 16-bit position re-basing (src/sliding_min.rs:117-125)."""
    n = 1_200_000
    seq = oracle.synth_packed(3, n + 64)
    pr = oracle.make_params(31, 19, canonical=True)
    epos, _ = oracle.run(seq, 0, n, pr)
    pos, _, _ = oracle.baseline_run_mt(seq, 0, n, pr, 1)
    assert np.array_equal(pos, epos)


def _shard_worker(rank, world, port, n, k, w, tmp):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, ROOT)
    import bench
    import mzoracle as o

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    l = k + w - 1
    nwin = n - l + 1
    per = (nwin + world - 1) // world
    wb, we = per * rank, min(per * (rank + 1), nwin)
    base_lo, base_hi = max(wb - 1, 0), we + l - 1              # halo + one seam window (bench.py rule)
    packed, off = bench.synth_packed_range(bench.SEED, base_lo, base_hi)
    pr = o.make_params(k, w, canonical=True)
    n_local = base_hi - base_lo
    pos, _ = o.run_range(np.ascontiguousarray(packed), off, n_local, pr, wb - base_lo, we - base_lo)
    pos = pos.astype(np.int64) + base_lo                         # back to global coordinates
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(pos)], dtype=torch.int64))
    mx = max(int(c) for c in counts)
    buf = torch.zeros(mx, dtype=torch.int64)
    buf[:len(pos)] = torch.from_numpy(pos)
    gathered = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, buf)
    if rank == 0:
        full = np.concatenate([g[:int(c)].numpy() for g, c in zip(gathered, counts)])
        np.save(os.path.join(tmp, "sharded.npy"), full)
    dist.destroy_process_group()


def test_two_rank_gloo_shards_concatenate(oracle, tmp_path):
    import torch.multiprocessing as mp

    import bench

    n, k, w = 300_000, 31, 19
    port = 29500 + os.getpid() % 2000
    mp.spawn(_shard_worker, args=(2, port, n, k, w, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npy")
    packed, off = bench.synth_packed_range(bench.SEED, 0, n)
    epos, _ = oracle.run(np.ascontiguousarray(packed), off, n, oracle.make_params(k, w, canonical=True))
    assert np.array_equal(got, epos.astype(np.int64))
