#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 minimizer path.

Metric (BASELINE.json): Gbp/s of canonical minimizer positions + u64 values, k=31 w=19, on a
synthetic uniform-random 2-bit packed 3.1 Gbp sequence (configs[1]).  One "step" = one pass of
the hot path over the whole sequence.  With N GPUs the windows are split into N contiguous
shards (strong scaling: the total stays 3.1 Gbp), one process per GPU, no data-path collective.

  value      device-resident: input and outputs in HBM, CUDA-event time on the library's stream
  e2e        host pinned buffers in -> host pinned buffers out through mz_run (H2D + D2H inside)
  roofline   algorithmic bytes of the kernel / its mean launch duration vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the oracle port of the reference's CPU algorithm (the crate itself is Rust and
             cannot be built in this image) on the box's host cores, on a bounded sample

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 42
CONFIGS = {
    # name: (k, w, canonical, mode, hasher, want_sk, value_bits)
    "c1": dict(k=21, w=11, canonical=False, mode=0, hasher="nt", want_sk=0, value_bits=0,
               n=10_000_000, desc="forward minimizer_positions k=21 w=11 NtHasher"),
    "c2": dict(k=31, w=19, canonical=True, mode=0, hasher="nt", want_sk=0, value_bits=64,
               n=3_100_000_000, desc="canonical_minimizer_positions + values_u64 k=31 w=19 NtHasher"),
    "c3": dict(k=31, w=19, canonical=True, mode=0, hasher="mul", want_sk=1, value_bits=64,
               n=3_100_000_000, desc="canonical_minimizers k=31 w=19 .super_kmers() + values_u64, MulHasher"),
    "c4": dict(k=31, w=11, canonical=True, mode=1, hasher="nt", want_sk=0, value_bits=128,
               n=3_100_000_000, desc="canonical closed syncmers k=31 w=11 positions + values_u128"),
}


def synth_words(seed: int, first_word: int, nwords: int) -> np.ndarray:
    """splitmix64(seed + i) for i in [first_word, first_word + nwords): 32 bases per u64.
    Byte-identical to oracle/mzoracle.c:mzo_synth_packed (tests/test_bench_utils.py)."""
    with np.errstate(over="ignore"):
        x = np.arange(first_word, first_word + nwords, dtype=np.uint64) + np.uint64(seed)
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def synth_packed_range(seed: int, base_lo: int, base_hi: int):
    """Packed bytes covering bases [base_lo, base_hi) of the synthetic sequence.
    Returns (uint8 array, bp_offset of base_lo inside it)."""
    w0 = base_lo // 32
    w1 = (base_hi + 31) // 32
    words = np.empty(w1 - w0 + 2, dtype=np.uint64)  # + padding
    chunk = 1 << 24
    for s in range(0, w1 - w0, chunk):
        e = min(s + chunk, w1 - w0)
        words[s:e] = synth_words(seed, w0 + s, e - s)
    words[w1 - w0:] = 0
    return words.view(np.uint8), base_lo - w0 * 32


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run (20 ms period).  Samples
    carry nvidia-smi's own timestamp; only those inside a marked busy window are used."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines, self.windows = gpu_index, None, [], []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self, t0: float, t1: float):
        """A timed (busy) window in time.time() seconds."""
        self.windows.append((t0, t1))

    def stop(self) -> dict:
        import datetime

        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), f[6:10]))
            except ValueError:
                continue
        busy = [r for r in rows if any(a - 0.02 <= r[0] <= b + 0.02 for a, b in self.windows)]
        use = busy or rows
        reasons = set()
        for r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in use])) if use else None,
                "sm_max_mhz": max(r[2] for r in use) if use else None, "reasons": sorted(reasons),
                "samples": len(use), "samples_total": len(rows)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes(config: str, n: int, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed ncu capture of this exact command (profiles/traffic.json); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(f"{config}:{n}:{world}")
        return None if e is None else float(e["dram_bytes_read"] + e["dram_bytes_write"])
    except Exception:
        return None


def oracle_params(o, cfg):
    hasher = o.make_hasher(cfg["hasher"], cfg["canonical"])
    return o.make_params(cfg["k"], cfg["w"], canonical=cfg["canonical"], mode=cfg["mode"], hasher=hasher)


def cpu_port_run(cfg, sample_bases: int, threads: int):
    """Time the oracle port (reference CPU algorithm) on `sample_bases` of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mzoracle as o

    o.build()
    key = (sample_bases,)
    if _CPU_CACHE.get("key") != key:  # synthetic sample is generated once, outside the timing
        _CPU_CACHE["key"] = key
        _CPU_CACHE["data"] = synth_packed_range(SEED, 0, sample_bases)
    packed, off = _CPU_CACHE["data"]
    pr = oracle_params(o, cfg)
    cap = int(sample_bases * (2.4 / (cfg["w"] + 1) if cfg["mode"] == 0 else 2.4 / cfg["w"])) + 65536
    t0 = time.perf_counter()
    # 8-lane AVX2 x pthreads restatement of the reference design (oracle/mzbaseline_avx2.c);
    # syncmer modes / wide values fall back to the scalar port inside
    pos, sk, val = o.baseline_run_mt(packed, off, sample_bases, pr, threads,
                                     want_sk=bool(cfg["want_sk"]),
                                     want_val=cfg["value_bits"] == 64 and cfg["mode"] == 0, cap=cap)
    dt = time.perf_counter() - t0
    return dt, len(pos)


_CPU_CACHE: dict = {}


def limit_host_threads(world: int) -> None:
    """One rank per GPU shares the host: split its cores between the ranks' copy / decode threads."""
    if world > 1 and "MZ_HOST_THREADS" not in os.environ:
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        os.environ["MZ_HOST_THREADS"] = str(max(2, min(16, cores // world)))


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; kind='port' because the
    Rust crate cannot be compiled here) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else cores
    n = cfg["n"] if args.n_bases is None else args.n_bases
    # calibrate: ~4 s of wall per step
    dt, _ = cpu_port_run(cfg, min(n, 4_000_000 * threads), threads)
    rate = min(n, 4_000_000 * threads) / dt
    sample = int(min(n, max(8_000_000, rate * 4.0)))
    times = []
    for i in range(args.warmup + args.steps):
        dt, cnt = cpu_port_run(cfg, sample, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    gbps = sample / (ms * 1e-3) / 1e9
    sample_desc = f"first {sample} bases of the {n}-base workload, {threads} threads, all outputs"
    line = {
        "impl": "reference", "metric": "Gbp/s canonical minimizer pos+vals (k=31,w=19)" if args.config == "c2" else f"Gbp/s {cfg['desc']}",
        "value": gbps, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "n_bases": n, "k": cfg["k"], "w": cfg["w"], "sample_bases": sample},
        "cpu_baseline": {"value": gbps, "unit": "Gbp/s", "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--n-bases", type=int, default=None, help="override sequence length (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    n = cfg["n"] if args.n_bases is None else args.n_bases
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    limit_host_threads(world)  # before libmzb200 reads MZ_HOST_THREADS
    sm = importlib.import_module("simd-minimizers_b200")
    ffi = importlib.import_module("simd-minimizers_b200._ffi")
    L = ffi.lib()

    k, w = cfg["k"], cfg["w"]
    l = k + w - 1
    nwin = n - l + 1
    per = (nwin + world - 1) // world
    wb, we = per * rank, min(per * (rank + 1), nwin)
    base_lo, base_hi = max(wb - 1, 0), we + l - 1

    # ---- synthetic shard on the host (pinned), then resident in HBM --------------------------
    host_np, off = synth_packed_range(SEED, base_lo, base_hi)
    host_pin = torch.empty(host_np.size, dtype=torch.uint8).pin_memory()
    host_pin.numpy()[:] = host_np
    del host_np
    d_in = host_pin.cuda(non_blocking=False)

    p = ffi.MzParams()
    (L.mz_params_mulhash if cfg["hasher"] == "mul" else L.mz_params_nthash)(
        C.byref(p), k, w, cfg["mode"], int(cfg["canonical"]))
    p.want_sk, p.value_bits = cfg["want_sk"], cfg["value_bits"]
    vw = cfg["value_bits"] // 64
    dens = 2.0 / (w + 1) if cfg["mode"] == 0 else (2.0 / w if cfg["mode"] == 1 else 1.0 / w)
    cap = int((we - wb) * dens * 1.15) + 65536
    d_pos = torch.empty(cap, dtype=torch.int32, device="cuda")
    d_sk = torch.empty(cap if cfg["want_sk"] else 1, dtype=torch.int32, device="cuda")
    d_val = torch.empty(max(cap * vw, 1), dtype=torch.int64, device="cuda")
    ctx = sm.Context([local_rank])
    # The shard is addressed as a stand-alone sequence of (base_hi - base_lo) bases: its windows
    # [wb - base_lo, we - base_lo) are produced, which keeps the left seam window (dedup rule).
    n_local = base_hi - base_lo
    lw0, lw1 = wb - base_lo, we - base_lo

    def step_device():
        out = ffi.MzOut(d_pos.data_ptr(), d_sk.data_ptr() if cfg["want_sk"] else None,
                        d_val.data_ptr() if vw else None, cap, 0)
        rc = L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n_local, lw0, lw1, C.byref(out))
        ffi.check(rc)
        t = ctx.last_timing()
        return out.count, t["kernel_ms"], t["kernel_launches"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs ~0.2 s to start: launch it before the warm-up steps and keep it sampling
    # (every 20 ms) through the device-resident and the end-to-end timed regions
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        count, _, _ = step_device()
    barrier()
    dev_ms, launches = [], 0
    t0, w0 = time.perf_counter(), time.time()
    for _ in range(args.steps):
        count, ms, nl = step_device()
        dev_ms.append(ms)
        launches += nl
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.mark(w0, time.time())
    ms_dev = float(np.mean(dev_ms))

    # ---- end to end: pinned host in -> pinned host out through mz_run ------------------------
    e2e_ms, h2d_bytes, d2h_bytes = None, 0, 0
    if not args.no_e2e:
        h_pos = torch.empty(cap, dtype=torch.int32).pin_memory()
        h_sk = torch.empty(cap if cfg["want_sk"] else 1, dtype=torch.int32).pin_memory()
        h_val = torch.empty(max(cap * vw, 1), dtype=torch.int64).pin_memory()

        def step_e2e():
            out = ffi.MzOut(h_pos.data_ptr(), h_sk.data_ptr() if cfg["want_sk"] else None,
                            h_val.data_ptr() if vw else None, cap, 0)
            # one process per GPU: the shard (with its halo) is run as a stand-alone sequence
            rc = L.mz_run(ctx.handle, C.byref(p), host_pin.data_ptr(), off, n_local, C.byref(out))
            ffi.check(rc)
            return out.count

        for _ in range(2):
            c2 = step_e2e()
        barrier()
        t0, w0 = time.perf_counter(), time.time()
        for _ in range(args.steps):
            c2 = step_e2e()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        sampler.mark(w0, time.time())
        h2d_bytes = (n_local * 2 + 7) // 8
        d2h_bytes = int(c2) * (4 + 4 * cfg["want_sk"] + 8 * vw)

    clocks = sampler.stop()

    # ---- reduce over ranks: max time, sum counts ----------------------------------------------
    stats = torch.tensor([ms_dev, wall_ms, e2e_ms or 0.0, float(count), float(launches),
                          float(h2d_bytes), float(d2h_bytes)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm_ = stats.clone()
        dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
        ms_dev, wall_ms, e2e_ms_max = mx[0].item(), mx[1].item(), mx[2].item()
        tot_count, tot_launch = sm_[3].item(), sm_[4].item()
        h2d_bytes, d2h_bytes = sm_[5].item(), sm_[6].item()
    else:
        e2e_ms_max, tot_count, tot_launch = (e2e_ms or 0.0), float(count), float(launches)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # algorithmic bytes of THIS rank's launch: 2-bit input read once + outputs written once
        alg_bytes = n_local * 0.25 + count * (4 + 4 * cfg["want_sk"] + 8 * vw)
        achieved = alg_bytes / (float(np.mean(dev_ms)) * 1e-3) / 1e9
        line = {
            "metric": "Gbp/s canonical minimizer pos+vals (k=31,w=19)" if args.config == "c2" else f"Gbp/s {cfg['desc']}",
            "value": n / (ms_dev * 1e-3) / 1e9, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "wall_ms_per_step": wall_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": cfg["desc"], "n_bases": n, "k": k, "w": w,
                       "outputs": int(tot_count), "parallelism": f"{world} contiguous window shards, halo k+w-2 (+1 seam window)",
                       "l2": "input shard %.0f MB > 126 MB L2; outputs rewritten every step" % (n_local / 4e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(args.config, n, world),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "kernel is integer-ALU bound, not HBM bound (ncu: ALU pipe ~78% "
                                 "of peak, DRAM ~12%); alu_roofline gives the bound that applies"},
            "clocks": clocks, "gpu_launches": int(tot_launch),
        }
        # north_star: "the roofline is the slower of bytes at HBM bandwidth and integer ops at
        # ALU peak".  ALU-pipe lane-ops per base are the ones this kernel executes, counted by ncu
        # (profiles/r1_fast_kernel_ncu_full.csv: sm__inst_executed_pipe_alu, 800 Mbp capture of the
        # C2 launch; LOP3/SHF/IMNMX/PRMT/ISETP share the pipe, 64 lanes/clk/SM, B300_MICROARCH
        # 'rt_SMSP=2'); the fraction is live: ops/bp x measured Gbp/s over SMs x 64 x sampled clock.
        # SURVEY 8(d)'s a-priori model was 35 ops/bp; the kernel needs fewer (3-input min, two bases
        # per table step), so that model would read > 1.
        ops_bp = {"c2": 24.8, "c3": 24.8}.get(args.config)
        if ops_bp is not None:
            sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
            alu_peak = 148 * 64 * sm_clock
            gbp_rank = (n_local / (float(np.mean(dev_ms)) * 1e-3))
            line["alu_roofline"] = {"alu_lane_ops_per_bp": ops_bp, "peak_lane_ops_per_s": alu_peak,
                                    "achieved_lane_ops_per_s": gbp_rank * ops_bp,
                                    "frac": gbp_rank * ops_bp / alu_peak,
                                    "source": "ops/bp from the committed ncu capture; peak = 148 SMs x 64 "
                                              "ALU lanes/clk x sampled SM clock"}
        if not args.no_e2e:
            line["e2e"] = {"value": n / (e2e_ms_max * 1e-3) / 1e9, "unit": "Gbp/s",
                           "ms_per_step": e2e_ms_max, "h2d_bytes_per_step": int(h2d_bytes),
                           "d2h_bytes_per_step": int(d2h_bytes)}
            # d2h_bytes_per_step = bytes delivered into the caller's host arrays.  Positions and
            # super-k-mer starts of minimizer runs (w <= 127) cross PCIe delta-coded (1 byte per
            # entry + a u32 per 256 entries, decoded by mz_run while it fills the arrays), so
            # fewer bytes are on the wire; MZ_NO_POS_DELTA=1 disables the codec.
            if cfg["mode"] == 0 and cfg["w"] <= 127 and not os.environ.get("MZ_NO_POS_DELTA"):
                narr = 1 + int(cfg["want_sk"])
                wire = tot_count * (8 * vw) + narr * (tot_count + 4 * ((tot_count + 255) // 256))
                line["e2e"]["d2h_wire_bytes_per_step"] = int(wire)
                line["e2e"]["d2h_codec"] = "pos/sk as int8 deltas + u32 base per 256 entries"
        if world == 1 and not args.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            sample = min(n, 32_000_000 * max(1, threads // 2))
            cpu_port_run(cfg, sample, threads)  # warm-up (page faults, thread start)
            dt, _ = cpu_port_run(cfg, sample, threads)
            s1 = min(n, 32_000_000)
            dt1, _ = cpu_port_run(cfg, s1, 1)
            line["cpu_baseline"] = {
                "value": sample / dt / 1e9, "unit": "Gbp/s", "cores": threads, "kind": "port",
                "sample": f"first {sample} bases, {threads} threads, 8-lane AVX2 restatement of the "
                          f"reference design (the Rust crate cannot be built here); single thread: "
                          f"{s1 / dt1 / 1e9:.4f} Gbp/s on {s1} bases "
                          f"(reference publishes ~0.46 Gbp/s/thread without values, BASELINE.md)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
