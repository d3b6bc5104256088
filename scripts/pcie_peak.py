"""Developer tool: pinned host -> device copy bandwidth in the e2e pipeline's pattern (chunks of a
batch's size on several streams), to put the e2e number next to what the link can do."""
import torch, time
dev = torch.device("cuda:0")
for mb, nstreams in ((7.2, 1), (7.2, 4), (64, 1), (2.6, 4)):
    n = int(mb * 1e6)
    hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(nstreams)]
    ds = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(nstreams)]
    ss = [torch.cuda.Stream() for _ in range(nstreams)]
    reps = 200
    for warm in (True, False):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(reps):
            with torch.cuda.stream(ss[i % nstreams]):
                ds[i % nstreams].copy_(hs[i % nstreams], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"H2D {mb} MB x {nstreams} streams: {reps * n / dt / 1e9:.1f} GB/s")
