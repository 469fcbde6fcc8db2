"""Does the D2H rate depend on how the copies are issued?  24 chunks of (13 MB + 103 MB) device -> pinned host:
one stream back to back / four streams round-robin / four streams with a small kernel in front of every chunk."""
import time, torch
n_chunks, a, b = 24, 13 << 20, 103 << 20
dev_a = [torch.empty(a, dtype=torch.uint8, device="cuda") for _ in range(4)]
dev_b = [torch.empty(b, dtype=torch.uint8, device="cuda") for _ in range(4)]
host_a = torch.empty(4 * a, dtype=torch.uint8).pin_memory()
host_b = torch.empty(n_chunks * b, dtype=torch.uint8).pin_memory()
streams = [torch.cuda.Stream() for _ in range(4)]
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(nstreams, kernel):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in range(n_chunks):
        s = streams[c % nstreams]
        with torch.cuda.stream(s):
            if kernel:
                dev_a[c % 4].add_(1)          # something that has to run before the copies of this chunk
            host_a[(c % 4) * a:(c % 4 + 1) * a].copy_(dev_a[c % 4], non_blocking=True)
            host_b[c * b:(c + 1) * b].copy_(dev_b[c % 4], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n_chunks * (a + b) / dt / 1e9

for name, ns, k in (("1 stream", 1, False), ("4 streams", 4, False), ("4 streams + kernel per chunk", 4, True), ("1 stream + kernel per chunk", 1, True)):
    r = [run(ns, k) for _ in range(4)]
    print(f"{name:32s}: {max(r):.1f} GB/s best, {sorted(r)[len(r)//2]:.1f} median")
# the same while a long kernel keeps the SMs and HBM busy
def busy():
    for _ in range(40):
        big.add_(1)
for name, ns in (("1 stream, GPU busy", 1), ("4 streams, GPU busy", 4)):
    r = []
    for _ in range(3):
        busy()
        r.append(run(ns, False))
    print(f"{name:32s}: {max(r):.1f} GB/s best")
