"""PCIe ceiling of the box: pinned D2H / H2D copies of one solution row (100.7 MB), alone and while a kernel streams HBM."""
import torch, time
n = 12587008
dev = torch.device("cuda:0")
src = torch.randn(n, dtype=torch.float64, device=dev)
big = torch.randn(4 * n, dtype=torch.float64, device=dev)
host = torch.zeros((8, n), dtype=torch.float64, pin_memory=True)
cs = torch.cuda.Stream()
def run(tag, reps=40, busy=False, chunks=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for i in range(reps):
            if chunks == 1:
                host[i % 8].copy_(src, non_blocking=True)
            else:
                m = n // chunks
                for c in range(chunks):
                    host[i % 8][c * m:(c + 1) * m].copy_(src[c * m:(c + 1) * m], non_blocking=True)
        e1.record(cs)
    if busy:
        while not e1.query():
            big.mul_(1.0000001)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag}: {ms:.3f} ms per row, {8 * n / ms / 1e6:.1f} GB/s", flush=True)
run("warm", 8)
run("D2H alone")
run("D2H while HBM busy", busy=True)
run("D2H 4 chunks", chunks=4)
# two streams, two halves
cs2 = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(40):
    with torch.cuda.stream(cs):
        host[i % 8][: n // 2].copy_(src[: n // 2], non_blocking=True)
    with torch.cuda.stream(cs2):
        host[i % 8][n // 2:].copy_(src[n // 2:], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 40
print(f"D2H two streams: {dt*1e3:.3f} ms per row, {8 * n / dt / 1e9:.1f} GB/s")
# H2D
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(40):
    src.copy_(host[i % 8], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 40
print(f"H2D: {dt*1e3:.3f} ms per row, {8 * n / dt / 1e9:.1f} GB/s")
# pageable-free check: time to allocate + zero 6 GB pinned
t0 = time.perf_counter()
h2 = torch.zeros((60, n), dtype=torch.float64, pin_memory=True)
print(f"pinned zeros 60 rows: {time.perf_counter() - t0:.2f} s")
