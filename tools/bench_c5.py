"""Config 5 shape: 150 bp reads at a 38-byte stride, canonical k=21 w=11, one GPU's share
(200M reads / 8 GPUs = 25M reads).  Prints kernel-only and end-to-end throughput."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200")
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
stride, read_len, k, w = 38, 150, 21, 11
words = bench.synth_words(bench.SEED, 0, (n_reads * stride + 7) // 8 + 8)
packed = words.view(np.uint8)
b = sm.canonical_minimizers(k, w)
ctx = sm.default_context()
for it in range(3):
    t0 = time.perf_counter()
    offs, pos, _, vals = b.run_batch(packed, stride_bytes=stride, read_len=read_len, n_reads=n_reads, value_bits=0)
    dt = time.perf_counter() - t0
    t = ctx.last_timing()
    bp = n_reads * read_len
    print(f"iter {it}: reads={n_reads} minimizers={len(pos)} ({len(pos)/n_reads:.2f}/read) kernel {t['kernel_ms']:.2f} ms = "
          f"{bp/t['kernel_ms']/1e6:.1f} Gbp/s ({n_reads/t['kernel_ms']/1e3:.1f} Mreads/s); gpu h2d+kernel+d2h {t['total_ms']:.1f} ms; wall {dt*1e3:.0f} ms")
