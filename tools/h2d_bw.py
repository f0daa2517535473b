"""Pinned host -> device copy bandwidth on this box (ceiling for the e2e metric): copy size and stream count."""
import torch, time
def run(mb, nstreams, reps=8):
    n = mb * (1 << 20)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    per = n // nstreams
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * per:(i + 1) * per].copy_(h[i * per:(i + 1) * per], non_blocking=True)
    go(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"H2D {mb} MiB over {nstreams} stream(s): {n / dt / 1e9:.1f} GB/s ({dt*1e3:.2f} ms)")
for mb in (16, 151, 604):
    for ns in (1, 2, 4):
        run(mb, ns)
