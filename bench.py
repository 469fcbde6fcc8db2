#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 minimizer path.

Metric (BASELINE.json): Gbp/s of canonical minimizer positions + u64 values, k=31 w=19, on a
synthetic uniform-random 2-bit packed 3.1 Gbp sequence (configs[1], `--config c2`, the default).
One "step" = one pass of the hot path over the whole workload.

  value      device-resident: input and outputs in HBM, CUDA-event time on the library's stream.
             With N GPUs (one process per GPU) the windows are split into N contiguous shards
             (halo k+w-2 bases + one seam window, no data-path collective); max over ranks.
  e2e        host pinned buffers in -> ONE ordered, globally indexed host output, through ONE
             context over all N devices (mz_ctx_create(ids, N) + mz_run: the product API), H2D
             and D2H inside the timed region.  Rank 0 drives it; the other ranks idle on a CPU
             barrier.  Next to it the measured copy ceiling of the same devices (mz_pcie_probe).
  roofline   algorithmic bytes of the kernel / its mean launch duration vs MEASURED_PEAKS.json,
             plus the integer-ALU roofline against a measured INT32 peak (mz_alu_probe).
  cpu_baseline / --impl reference
             the oracle port of the reference's CPU algorithm (the crate itself is Rust and
             cannot be built in this image) on the box's host cores.

`--config c5` is BASELINE configs[4]: 25 M x 150 bp reads per GPU (200 M on 8), canonical k=21
w=11 per-read minimizers through mz_run_batch, reads dealt over the context's devices.

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
    "c5": dict(k=21, w=11, canonical=True, mode=0, hasher="nt", want_sk=0, value_bits=0,
               reads_per_gpu=25_000_000, read_len=150, stride=38,
               desc="batched short reads: 150 bp reads at a 38-byte stride, canonical k=21 w=11 per-read "
                    "minimizer positions (CSR), 25 M reads per GPU (200 M on 8)"),
}
CHECK_PREFIX = 1 << 20  # entries of the output whose checksum both arms print in `config`


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


def synth_fill(dst_u64: np.ndarray, seed: int, first_word: int) -> None:
    """Fill a uint64 view with the synthetic stream starting at word `first_word`."""
    chunk = 1 << 24
    for s in range(0, dst_u64.size, chunk):
        e = min(s + chunk, dst_u64.size)
        dst_u64[s:e] = synth_words(seed, first_word + s, e - s)


def synth_packed_range(seed: int, base_lo: int, base_hi: int):
    """Packed bytes covering bases [base_lo, base_hi) of the synthetic sequence.
    Returns (uint8 array, bp_offset of base_lo inside it)."""
    w0 = base_lo // 32
    w1 = (base_hi + 31) // 32
    words = np.empty(w1 - w0 + 2, dtype=np.uint64)  # + padding
    synth_fill(words[:w1 - w0], seed, w0)
    words[w1 - w0:] = 0
    return words.view(np.uint8), base_lo - w0 * 32


def checksum_entries(pos: np.ndarray, val: np.ndarray | None) -> str:
    """64-bit checksums of an output stream: wrapping sums of the positions and of the values
    (u128 values: of their 64-bit halves)."""
    with np.errstate(over="ignore"):
        ps = int(pos.astype(np.uint64).sum(dtype=np.uint64))
        vs = int(val.view(np.uint64).sum(dtype=np.uint64)) if val is not None else 0
    return f"pos:{ps:016x} val:{vs:016x}"


WORLD = 1  # --gpus N (both arms): the per-GPU share decides whether the L2 has to be flushed


def needs_l2_flush(n_bases: int) -> bool:
    """Timing rule: inputs must be larger than the 126 MB L2 or the L2 is flushed between timed steps."""
    return n_bases / 4 <= 126e6 * 1.5


def workload_config(name: str, cfg: dict, n: int, extra: dict | None = None) -> dict:
    """The `config` object: identical in the b200 and the reference arm (the driver compares them)."""
    if name == "c5":
        c = {"workload": cfg["desc"], "reads_per_gpu": cfg["reads_per_gpu"], "read_len": cfg["read_len"],
             "stride_bytes": cfg["stride"], "k": cfg["k"], "w": cfg["w"],
             "parallelism": "reads are independent units: dealt chunk by chunk over the GPUs of one context "
                            "(b200 arm) / over host threads (reference arm); no halo, no collective",
             "l2": "input %.0f MB per GPU > 126 MB L2; outputs rewritten every step" % (cfg["reads_per_gpu"] * cfg["stride"] / 1e6)}
    else:
        c = {"workload": cfg["desc"], "n_bases": n, "k": cfg["k"], "w": cfg["w"],
             "parallelism": "contiguous window shards, halo k+w-2 bases (+1 seam window): one per GPU for the "
                            "device-resident value, chunks dealt over the GPUs of one context for e2e (b200 arm) / "
                            "one per host thread (reference arm); no collective on the data path",
             "l2": ("input %.0f MB per GPU > 126 MB L2; outputs rewritten every step" % (n / 4e6 / WORLD))
                   if not needs_l2_flush(n // WORLD)
                   else ("input %.1f MB per GPU is not larger than 1.5x the 126 MB L2: a 256 MB buffer is written between timed steps "
                         "(L2 flush)" % (n / 4e6 / WORLD))}
    if extra:
        c.update(extra)
    return c


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


def profile_entry(config: str, n: int, world: int):
    """Numbers taken from the committed ncu captures of this exact command (profiles/traffic.json):
    dram bytes per launch of the dominant kernel and its executed ALU-pipe lane-ops per base."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(f"{config}:{n}:{world}")
    except Exception:
        return None


def oracle_mod():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mzoracle as o

    o.build()
    return o


def oracle_params(o, cfg):
    hasher = o.make_hasher(cfg["hasher"], cfg["canonical"])
    return o.make_params(cfg["k"], cfg["w"], canonical=cfg["canonical"], mode=cfg["mode"], hasher=hasher)


# ---- CPU arm ---------------------------------------------------------------------------------
class CpuArm:
    """The oracle port of the reference's CPU algorithm on `sample_bases` of the workload.  Input and
    output arrays are allocated and touched ONCE, outside every timed region; each thread writes
    straight into its own slice of the output (no allocation, no concatenation inside the timing --
    the reference's own multi-threaded benchmark keeps a thread-local Vec per worker,
    bench/src/bin/paper.rs:439-461)."""

    def __init__(self, cfg, sample_bases: int, max_threads: int):
        self.o = oracle_mod()
        self.cfg, self.n = cfg, sample_bases
        self.packed, self.off = synth_packed_range(SEED, 0, sample_bases)
        self.pr = oracle_params(self.o, cfg)
        dens = 2.0 / (cfg["w"] + 1) if cfg["mode"] == 0 else 2.0 / cfg["w"]
        self.cap = int(sample_bases * dens * 1.3) + 65536 * max_threads
        self.fast = cfg["mode"] == 0  # AVX2 lanes; syncmers / wide values: scalar port
        self.pos = np.zeros(self.cap, dtype=np.uint32)
        self.sk = np.zeros(self.cap, dtype=np.uint32) if cfg["want_sk"] else None
        self.val = np.zeros(self.cap, dtype=np.uint64) if (cfg["value_bits"] == 64 and self.fast) else None

    def run(self, threads: int):
        """-> (seconds, count, (starts, counts))"""
        t0 = time.perf_counter()
        if self.fast:
            m, st, ct = self.o.baseline_run_mt_slices(self.packed, self.off, self.n, self.pr, threads,
                                                      self.pos, self.sk, self.val)
        else:
            p, _, _ = self.o.run_mt(self.packed, self.off, self.n, self.pr, threads, cap=self.cap)
            m, st, ct = len(p), np.array([0], dtype=np.uint64), np.array([len(p)], dtype=np.uint64)
            self.pos[:m] = p
        return time.perf_counter() - t0, m, (st, ct)

    def prefix_checksum(self, slices) -> str | None:
        st, ct = slices
        if int(ct[0]) < CHECK_PREFIX:
            return None
        a = int(st[0])
        return checksum_entries(self.pos[a:a + CHECK_PREFIX], None if self.val is None else self.val[a:a + CHECK_PREFIX])

    def full_checksum(self, slices) -> str:
        st, ct = slices
        with np.errstate(over="ignore"):
            ps = sum(int(self.pos[int(a):int(a + c)].astype(np.uint64).sum(dtype=np.uint64)) for a, c in zip(st, ct)) & (2**64 - 1)
            vs = 0 if self.val is None else sum(int(self.val[int(a):int(a + c)].sum(dtype=np.uint64)) for a, c in zip(st, ct)) & (2**64 - 1)
        return f"pos:{ps:016x} val:{vs:016x}"


def host_threads_available() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference(args, name, cfg, rank):
    """--impl reference: the reference's CPU algorithm (oracle port; kind='port' because the
    Rust crate cannot be compiled here) on all host threads."""
    if rank != 0:
        return
    threads = host_threads_available()
    if name == "c5":
        return run_reference_c5(args, cfg, threads)
    n = cfg["n"] if args.n_bases is None else args.n_bases
    # the whole workload when a step then takes no more than ~6 s, else a bounded prefix of it
    probe = CpuArm(cfg, min(n, 16_000_000 * threads), threads)
    probe.run(threads)
    dt, _, _ = probe.run(threads)
    rate = probe.n / dt
    sample = n if n / rate <= 6.0 else int(max(64_000_000, rate * 4.0))
    del probe
    arm = CpuArm(cfg, sample, threads)
    times, m, sl = [], 0, None
    for i in range(args.warmup + args.steps):
        dt, m, sl = arm.run(threads)
        if i >= args.warmup:
            times.append(dt)
    dt1, _, _ = arm.run(1) if sample <= 400_000_000 else CpuArm(cfg, 100_000_000, 1).run(1)
    n1 = sample if sample <= 400_000_000 else 100_000_000
    ms = 1e3 * float(np.mean(times))
    gbps = sample / (ms * 1e-3) / 1e9
    what = "the whole workload" if sample == n else f"first {sample} bases of the {n}-base workload"
    line = {
        "impl": "reference", "metric": metric_name(name, cfg),
        "value": gbps, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": workload_config(name, cfg, n, {"checksum_first_%d_entries" % CHECK_PREFIX: arm.prefix_checksum(sl)}),
        "result": {"sample_bases": sample, "outputs": int(m), "checksum": arm.full_checksum(sl),
                   "single_thread_gbps": n1 / dt1 / 1e9},
        "cpu_baseline": {"value": gbps, "unit": "Gbp/s", "cores": threads, "kind": "port",
                         "sample": f"{what}, {threads} threads, all outputs, 8-lane AVX2 restatement of the reference "
                                   f"design; single thread: {n1 / dt1 / 1e9:.4f} Gbp/s on {n1} bases"},
        "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def c5_reads(cfg, args, world: int) -> int:
    per = cfg["reads_per_gpu"] if args.reads_per_gpu is None else args.reads_per_gpu
    return per * world


def run_reference_c5(args, cfg, threads):
    """The per-read caller loop (bench/src/bin/paper.rs:98-105, examples/bench.rs:63-89) over
    host threads, each read a stand-alone sequence."""
    o = oracle_mod()
    n_reads_total = c5_reads(cfg, args, max(1, args.gpus))
    stride, rl = cfg["stride"], cfg["read_len"]
    pr = oracle_params(o, cfg)
    sample = min(n_reads_total, 2_000_000 * threads)
    words = np.empty((sample * stride + 7) // 8 + 8, dtype=np.uint64)
    synth_fill(words, SEED, 0)
    packed = words.view(np.uint8)
    nwin = rl - (cfg["k"] + cfg["w"] - 1) + 1
    cap = int(sample * nwin * 2.0 / (cfg["w"] + 1) * 1.5) + 65536 * threads
    bufs = (np.zeros(sample + 1, dtype=np.uint64), np.zeros(cap, dtype=np.uint32), None, None)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        offs, pos, _, _ = o.run_reads(packed, sample, stride, rl, pr, threads, bufs=bufs)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    s1 = min(sample, 1_000_000)
    o.run_reads(packed, s1, stride, rl, pr, 1, bufs=bufs)
    dt1 = time.perf_counter() - t0
    ms = 1e3 * float(np.mean(times))
    gbps = sample * rl / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": metric_name("c5", cfg), "value": gbps, "unit": "Gbp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config("c5", cfg, 0, {"checksum_first_%d_entries" % CHECK_PREFIX: checksum_entries(pos[:CHECK_PREFIX], None)}),
        "result": {"sample_reads": sample, "outputs": int(len(pos)), "single_thread_gbps": s1 * rl / dt1 / 1e9},
        "cpu_baseline": {"value": gbps, "unit": "Gbp/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample} of {n_reads_total} reads, per-read scalar loop over {threads} threads "
                                   f"(one scratch ring per thread); single thread {s1 * rl / dt1 / 1e9:.4f} Gbp/s"},
        "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(name, cfg):
    return "Gbp/s canonical minimizer pos+vals (k=31,w=19)" if name == "c2" else f"Gbp/s {cfg['desc']}"


def limit_host_threads(world: int) -> None:
    """Rank 0 drives the whole end-to-end run (all devices, one context): it keeps the library's
    default of 16 copy / decode threads; nothing to split between ranks any more."""
    return None


class HostBuf:
    """Pinned host memory from the library's own allocator (mz_host_alloc: spread over the NUMA
    nodes of the host so that the devices of both sockets reach it at the same rate)."""

    def __init__(self, L, ffi, nbytes: int, dtype):
        self.L, self.p = L, C.c_void_p()
        self.nbytes = max(int(nbytes), 8)
        ffi.check(L.mz_host_alloc(C.byref(self.p), self.nbytes))
        self.np = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.p.value)).view(dtype)

    @property
    def ptr(self):
        return self.p.value

    def free(self):
        if self.p:
            self.np = None
            self.L.mz_host_free(self.p)
            self.p = C.c_void_p()


def pcie_bound_ms(r, h2d_bytes: float, d2h_bytes: float) -> float:
    """Lower bound on the copy time given the measured link rates: each direction alone, and both
    directions sharing the measured bidirectional rates."""
    alone = max(h2d_bytes / (r.h2d_gbs * 1e9), d2h_bytes / (r.d2h_gbs * 1e9))
    both = max(h2d_bytes / (r.bidir_h2d_gbs * 1e9), d2h_bytes / (r.bidir_d2h_gbs * 1e9))
    # the shorter stream overlaps the longer one at the bidirectional rate, the rest runs alone
    t_h, t_d = h2d_bytes / (r.bidir_h2d_gbs * 1e9), d2h_bytes / (r.bidir_d2h_gbs * 1e9)
    if t_h < t_d:
        mixed = t_h + (d2h_bytes - t_h * r.bidir_d2h_gbs * 1e9) / (r.d2h_gbs * 1e9)
    else:
        mixed = t_d + (h2d_bytes - t_d * r.bidir_h2d_gbs * 1e9) / (r.h2d_gbs * 1e9)
    return 1e3 * max(alone, min(both, mixed))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--n-bases", type=int, default=None, help="override sequence length (testing)")
    ap.add_argument("--reads-per-gpu", type=int, default=None, help="c5: override reads per GPU (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    global WORLD
    WORLD = max(1, args.gpus)
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, args.config, cfg, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # ranks that wait for rank 0's end-to-end run must not spin on their GPU (rank 0's context
        # uses every device): they wait on a CPU (gloo) barrier
        cpu_group = dist.new_group(backend="gloo")
    sm = importlib.import_module("simd-minimizers_b200")
    ffi = importlib.import_module("simd-minimizers_b200._ffi")
    L = ffi.lib()
    env = dict(torch=torch, dist=dist, sm=sm, ffi=ffi, L=L, rank=rank, world=world, local_rank=local_rank,
               cpu_group=cpu_group, args=args)
    if args.config == "c5":
        bench_reads(env, cfg)
    else:
        bench_sequence(env, args.config, cfg)
    if world > 1:
        dist.destroy_process_group()


def alu_probe(L, ffi, ctx):
    r = ffi.MzAluResult()
    ffi.check(L.mz_alu_probe(ctx.handle, 0, C.byref(r)))
    return r


def bench_sequence(env, name, cfg):
    torch, dist, sm, ffi, L = env["torch"], env["dist"], env["sm"], env["ffi"], env["L"]
    rank, world, local_rank, args = env["rank"], env["world"], env["local_rank"], env["args"]
    n = cfg["n"] if args.n_bases is None else args.n_bases
    k, w = cfg["k"], cfg["w"]
    l = k + w - 1
    nwin = n - l + 1
    per = (nwin + world - 1) // world
    wb, we = min(per * rank, nwin), min(per * (rank + 1), nwin)
    base_lo, base_hi = max(wb - 1, 0), we + l - 1

    # ---- this rank's shard, resident in HBM ----------------------------------------------------
    host_np, off = synth_packed_range(SEED, base_lo, base_hi)
    d_in = torch.from_numpy(host_np).cuda()
    del host_np

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

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if needs_l2_flush(n_local) else None

    def step_device():
        if flush is not None:  # small inputs: evict them (and the outputs) from L2 before every step
            flush.fill_(1)
            torch.cuda.synchronize()  # (the library launches on its own stream)
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
    alu = alu_probe(L, ffi, ctx) if rank == 0 else None
    # checksums of this rank's device-resident shard output (wrapping sums add up over ranks)
    # (the shard is addressed as a stand-alone sequence: its positions are relative to base_lo)
    shard_ps = (int(d_pos[:count].to(torch.int64).bitwise_and(0xffffffff).sum().item()) + int(count) * base_lo) & (2**64 - 1)
    shard_vs = int(d_val[:count * vw].sum().item()) & (2**64 - 1) if vw else 0

    # ---- reduce over ranks: max time, sum counts / checksums ------------------------------------
    if world > 1:
        mx = torch.tensor([ms_dev, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ms_dev_max, wall_ms = mx[0].item(), mx[1].item()
        # exact integer sums: split the 64-bit checksums into 32-bit halves
        sums = torch.tensor([count, launches, shard_ps & 0xffffffff, shard_ps >> 32, shard_vs & 0xffffffff, shard_vs >> 32],
                            dtype=torch.int64, device="cuda")
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        s = [int(x) for x in sums.tolist()]
        tot_count, tot_launch = s[0], s[1]
        tot_ps = (s[2] + (s[3] << 32)) & (2**64 - 1)
        tot_vs = (s[4] + (s[5] << 32)) & (2**64 - 1)
    else:
        ms_dev_max, tot_count, tot_launch, tot_ps, tot_vs = ms_dev, int(count), int(launches), shard_ps, shard_vs
    dev_checksum = f"pos:{tot_ps:016x} val:{tot_vs:016x}"

    # ---- end to end: ONE context over all devices, pinned host in -> ONE ordered host output ----
    e2e = None
    if not args.no_e2e:
        del d_pos, d_sk, d_val, d_in
        torch.cuda.empty_cache()
        if rank == 0:
            e2e = e2e_sequence(env, name, cfg, p, n, sampler, tot_count)
            # ONE ordered host output of the N-device context == the N ranks' device-resident
            # shards (count and 64-bit checksums of positions and values)
            assert e2e["checksum"] == dev_checksum, (e2e["checksum"], dev_checksum)
        if world > 1:
            dist.barrier(group=env["cpu_group"])

    clocks = sampler.stop()
    if rank != 0:
        return
    peak, peak_src = measured_peak_gbs()
    # algorithmic bytes of THIS rank's launch: 2-bit input read once + outputs written once
    alg_bytes = n_local * 0.25 + count * (4 + 4 * cfg["want_sk"] + 8 * vw)
    achieved = alg_bytes / (ms_dev * 1e-3) / 1e9
    prof = profile_entry(name, n, world) or {}
    extra = {}
    if e2e and e2e.get("prefix_checksum"):
        extra["checksum_first_%d_entries" % CHECK_PREFIX] = e2e["prefix_checksum"]
    line = {
        "metric": metric_name(name, cfg),
        "value": n / (ms_dev_max * 1e-3) / 1e9, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev_max, "wall_ms_per_step": wall_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": workload_config(name, cfg, n, extra),
        "result": {"outputs": int(tot_count), "checksum_device_shards": dev_checksum,
                   "checksum": e2e["checksum"] if e2e else None,
                   "checksum_note": "wrapping 64-bit sums of all positions / values: equal for every N, and equal between "
                                    "the ranks' device-resident shards and the one host output of the N-device context"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (prof["dram_bytes_read"] + prof["dram_bytes_write"]) if "dram_bytes_read" in prof else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "the kernel is integer-ALU bound, not HBM bound (ncu: ALU pipe busy, DRAM ~12%); "
                             "alu_roofline gives the bound that applies"},
        "clocks": clocks, "gpu_launches": int(tot_launch),
    }
    # north_star: "the roofline is the slower of bytes at HBM bandwidth and integer ops at ALU peak".
    # Peak = INT32 lane-ops/s of a LOP3/SHF/VIMNMX/PRMT mix measured by mz_alu_probe on
    # this GPU right now; executed ops/bp come from the committed ncu capture of this config
    # (profiles/traffic.json: sm__inst_executed_pipe_alu x 32 lanes / bases); useful ops/bp is the
    # written-down minimum of the algorithm, so `useful_frac` cannot be inflated by
    # wasteful code.
    if alu is not None:
        gbp_rank = n_local / (ms_dev * 1e-3)
        ar = {"alu_peak_measured_lane_ops_per_s": alu.lane_ops_per_s, "probe_ms": alu.ms,
              "probe_sm_mhz_implied": alu.lane_ops_per_s / (148 * 64) / 1e6,
              "source": "peak: mz_alu_probe (INT32 LOP3/SHF/VIMNMX/PRMT mix, 8 independent chains per thread, all SMs)"}
        if "alu_lane_ops_per_bp" in prof:
            ar["alu_lane_ops_per_bp_executed"] = prof["alu_lane_ops_per_bp"]
            ar["frac"] = gbp_rank * prof["alu_lane_ops_per_bp"] / alu.lane_ops_per_s
        useful = USEFUL_OPS_PER_BP.get(name)
        if useful:
            ar["useful_lane_ops_per_bp"] = useful
            ar["useful_frac"] = gbp_rank * useful / alu.lane_ops_per_s
        line["alu_roofline"] = ar
    if e2e:
        line["e2e"] = e2e["line"]
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads_available()
        sample = min(n, 32_000_000 * max(1, threads // 2))
        arm = CpuArm(cfg, sample, threads)
        arm.run(threads)  # warm-up (page faults, thread start)
        dt, _, _ = arm.run(threads)
        s1 = min(n, 32_000_000)
        arm1 = CpuArm(cfg, s1, 1)
        arm1.run(1)
        dt1, _, _ = arm1.run(1)
        line["cpu_baseline"] = {
            "value": sample / dt / 1e9, "unit": "Gbp/s", "cores": threads, "kind": "port",
            "sample": f"first {sample} bases, {threads} threads, 8-lane AVX2 restatement of the "
                      f"reference design (the Rust crate cannot be built here), outputs pre-allocated; single thread: "
                      f"{s1 / dt1 / 1e9:.4f} Gbp/s on {s1} bases "
                      f"(reference publishes ~0.46 Gbp/s/thread without values, BASELINE.md)"}
    print(json.dumps(line), flush=True)


# Minimal INT32 lane-ops per base of the algorithm itself (DESIGN.md section 6 derives them): canonical
# minimizers with values -- 2 strands x (rotate, xor) + add, key|pos pack, complement key,
# 2 x (prefix min, suffix min, window min), duplicate flag, tie check; + values at density 0.1.
USEFUL_OPS_PER_BP = {"c2": 17.0, "c3": 17.0}


def e2e_sequence(env, name, cfg, p, n, sampler, expect_count):
    """Rank 0: mz_run over the whole sequence through one context that spans all `world` devices."""
    torch, sm, ffi, L, world, args = env["torch"], env["sm"], env["ffi"], env["L"], env["world"], env["args"]
    k, w = cfg["k"], cfg["w"]
    vw = cfg["value_bits"] // 64
    dens = 2.0 / (w + 1) if cfg["mode"] == 0 else (2.0 / w if cfg["mode"] == 1 else 1.0 / w)
    cap = int(n * dens * 1.08) + 65536
    nwords = (n + 31) // 32
    h_in = HostBuf(L, ffi, (nwords + 2) * 8, np.uint64)
    synth_fill(h_in.np[:nwords], SEED, 0)
    h_in.np[nwords:] = 0
    h_pos = HostBuf(L, ffi, cap * 4, np.uint32)
    h_sk = HostBuf(L, ffi, cap * 4 if cfg["want_sk"] else 8, np.uint32)
    h_val = HostBuf(L, ffi, cap * 8 * vw if vw else 8, np.uint64)
    ctx = sm.Context(list(range(world)))

    def step():
        out = ffi.MzOut(h_pos.ptr, h_sk.ptr if cfg["want_sk"] else None, h_val.ptr if vw else None, cap, 0)
        ffi.check(L.mz_run(ctx.handle, C.byref(p), h_in.ptr, 0, n, C.byref(out)))
        return int(out.count)

    for _ in range(2):
        cnt = step()
    times = []
    w0 = time.time()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cnt = step()
        times.append((time.perf_counter() - t0) * 1e3)
    sampler.mark(w0, time.time())
    tm = ctx.last_timing()
    assert cnt == expect_count, f"end-to-end count {cnt} != sum of the ranks' device-resident shards {expect_count}"
    e2e_ms = float(np.mean(times))
    h2d_bytes = (2 * n + 7) // 8
    d2h_bytes = cnt * (4 + 4 * cfg["want_sk"] + 8 * vw)
    wire = d2h_bytes
    codec = None
    if cfg["mode"] == 0 and w <= 127 and not os.environ.get("MZ_NO_POS_DELTA") and \
            world <= int(os.environ.get("MZ_DELTA_MAX_DEVICES", "2")):
        narr = 1 + int(cfg["want_sk"])
        wire = cnt * (8 * vw) + narr * (cnt + 4 * ((cnt + 255) // 256))
        codec = "pos/sk as int8 deltas + u32 base per 256 entries"
    r = ffi.MzPcieResult()
    ffi.check(L.mz_pcie_probe(ctx.handle, 256 << 20, 4, C.byref(r)))
    bound = pcie_bound_ms(r, h2d_bytes, wire)
    pos = h_pos.np[:cnt]
    val = h_val.np[:cnt * vw] if vw else None
    res = {
        "checksum": checksum_entries(pos, val),
        "prefix_checksum": checksum_entries(pos[:CHECK_PREFIX], None if val is None else val[:CHECK_PREFIX * vw])
        if cnt >= CHECK_PREFIX else None,
        "line": {"value": n / (e2e_ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": e2e_ms,
                 "ms_min": float(np.min(times)),
                 "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                 "d2h_wire_bytes_per_step": int(wire), "d2h_codec": codec,
                 "api": f"mz_ctx_create({world} devices) + mz_run, host pinned (mz_host_alloc) in/out, one ordered output",
                 "pcie_peak_gbs": {"h2d": r.h2d_gbs, "d2h": r.d2h_gbs, "both_h2d": r.bidir_h2d_gbs,
                                   "both_d2h": r.bidir_d2h_gbs, "devices": r.n_devices,
                                   "how": "mz_pcie_probe: 4 x 256 MiB pinned copies per device and direction, all devices at once"},
                 "pcie_bound_ms": bound, "frac": bound / e2e_ms,
                 "phases_busiest_device_ms": {"h2d": tm["h2d_ms"], "kernels": tm["kernel_ms"], "d2h": tm["d2h_ms"]},
                 "kernel_launches_per_step": tm["kernel_launches"]},
    }
    ctx.close()
    for b in (h_in, h_pos, h_sk, h_val):
        b.free()
    return res


def bench_reads(env, cfg):
    """BASELINE configs[4]: batched short reads through mz_run_batch on one context over all devices."""
    torch, dist, sm, ffi, L = env["torch"], env["dist"], env["sm"], env["ffi"], env["L"]
    rank, world, local_rank, args = env["rank"], env["world"], env["local_rank"], env["args"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        n_reads = c5_reads(cfg, args, world)
        stride, rl, k, w = cfg["stride"], cfg["read_len"], cfg["k"], cfg["w"]
        nbytes = n_reads * stride
        h_in = HostBuf(L, ffi, (nbytes + 7) // 8 * 8 + 64, np.uint64)
        synth_fill(h_in.np[:(nbytes + 7) // 8], SEED, 0)
        nwin = rl - (k + w - 1) + 1
        cap = int(n_reads * (nwin * 2.0 / (w + 1) * 1.12 + 1)) + 65536
        h_pos = HostBuf(L, ffi, cap * 4, np.uint32)
        h_off = HostBuf(L, ffi, (n_reads + 1) * 8, np.uint64)
        p = ffi.MzParams()
        L.mz_params_nthash(C.byref(p), k, w, cfg["mode"], int(cfg["canonical"]))
        p.want_sk, p.value_bits = 0, 0
        ctx = sm.Context(list(range(world)))

        def step():
            out = ffi.MzOut(h_pos.ptr, None, None, cap, 0)
            ffi.check(L.mz_run_batch(ctx.handle, C.byref(p), h_in.ptr, nbytes, n_reads, None, None, stride, rl,
                                     h_off.ptr, C.byref(out)))
            return int(out.count), ctx.last_timing()

        for _ in range(max(2, min(args.warmup, 3))):
            cnt, tm = step()
        times, kms, launches = [], [], 0
        w0 = time.time()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            cnt, tm = step()
            times.append((time.perf_counter() - t0) * 1e3)
            kms.append(tm["kernel_ms"])
            launches += tm["kernel_launches"]
        sampler.mark(w0, time.time())
        clocks = sampler.stop()
        bases = n_reads * rl
        e2e_ms, ker_ms = float(np.mean(times)), float(np.mean(kms))
        h2d, d2h = nbytes, cnt * 4 + (n_reads + 1) * 8
        r = ffi.MzPcieResult()
        ffi.check(L.mz_pcie_probe(ctx.handle, 256 << 20, 4, C.byref(r)))
        bound = pcie_bound_ms(r, h2d, d2h)
        peak, peak_src = measured_peak_gbs()
        offs = h_off.np[:n_reads + 1]
        assert int(offs[-1]) == cnt and (np.diff(offs[:100000].astype(np.int64)) >= 0).all()
        # kernel_ms = the busiest device's sum of its chunk launches (CUDA events): the device-side rate
        alg_bytes = (h2d + d2h) / world
        line = {
            "metric": metric_name("c5", cfg), "value": bases / (ker_ms * 1e-3) / 1e9, "unit": "Gbp/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(2, min(args.warmup, 3)), "ms_per_step": ker_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config("c5", cfg, 0, {"checksum_first_%d_entries" % CHECK_PREFIX: checksum_entries(h_pos.np[:CHECK_PREFIX], None)}),
            "result": {"reads": n_reads, "outputs": cnt, "checksum": checksum_entries(h_pos.np[:cnt], None),
                       "value_is": "all reads x 150 bp / (busiest device's kernel time per step: CUDA events around "
                                   "every chunk launch, minus the time a launch spent queued behind the device's "
                                   "previous chunk); the launches run while other chunks' copies are in flight"},
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ker_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes / (ker_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "per device and step: packed reads in + positions + CSR offsets out (SURVEY 8d: 0.863 B/bp)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": bases / (e2e_ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": f"mz_ctx_create({world} devices) + mz_run_batch, host pinned in/out, CSR output",
                    "pcie_peak_gbs": {"h2d": r.h2d_gbs, "d2h": r.d2h_gbs, "both_h2d": r.bidir_h2d_gbs,
                                      "both_d2h": r.bidir_d2h_gbs, "devices": r.n_devices},
                    "pcie_bound_ms": bound, "frac": bound / e2e_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            o = oracle_mod()
            threads = host_threads_available()
            sample = min(n_reads, 1_000_000 * threads)
            pr = oracle_params(o, cfg)
            ccap = int(sample * nwin * 2.0 / (w + 1) * 1.5) + 65536 * threads
            bufs = (np.zeros(sample + 1, dtype=np.uint64), np.zeros(ccap, dtype=np.uint32), None, None)
            packed = h_in.np.view(np.uint8)
            o.run_reads(packed, sample, stride, rl, pr, threads, bufs=bufs)
            t0 = time.perf_counter()
            coffs, cpos, _, _ = o.run_reads(packed, sample, stride, rl, pr, threads, bufs=bufs)
            dt = time.perf_counter() - t0
            assert np.array_equal(coffs, offs[:sample + 1]) and np.array_equal(cpos, h_pos.np[:int(offs[sample])]), \
                "GPU batch output differs from the per-read CPU loop"
            line["cpu_baseline"] = {"value": sample * rl / dt / 1e9, "unit": "Gbp/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sample} reads, per-read scalar loop of the oracle over {threads} threads "
                                              f"(bench/src/bin/paper.rs:98-105); output identical to the GPU's for these reads"}
        print(json.dumps(line), flush=True)
        ctx.close()
    if world > 1:
        dist.barrier(group=env["cpu_group"])


if __name__ == "__main__":
    main()
