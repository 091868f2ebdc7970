"""PCIe probe: H2D bandwidth from pinned memory with 1 / 2 / 4 concurrent streams, with and without a concurrent D2H."""
import os, torch, time
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))   # under torchrun: one probe per GPU, all at once
n = 1 << 28
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
ho = torch.empty(n // 8, dtype=torch.float32).pin_memory()
do = torch.empty(n // 8, dtype=torch.float32, device="cuda")
def run(k, with_d2h):
    ss = [torch.cuda.Stream() for _ in range(k)]
    so = torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = n // k
    for i, s in enumerate(ss):
        with torch.cuda.stream(s):
            d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
    if with_d2h:
        with torch.cuda.stream(so):
            ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0
for k in ((1,) if 'LOCAL_RANK' in os.environ else (1, 2, 4)):
    for w in (False, True):
        run(k, w)
        t = min(run(k, w) for _ in range(3))
        print(f"[gpu {torch.cuda.current_device()}] {k} H2D stream(s), D2H concurrently={w}: {t*1e3:.2f} ms  H2D {n*4/t/1e9:.1f} GB/s", flush=True)
