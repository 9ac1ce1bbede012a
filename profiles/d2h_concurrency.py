import time, torch
dev = torch.device("cuda:0")
n = 4096 * 961
src = torch.randn(n, device=dev)
dst = torch.empty(n).pin_memory()
s1, s2, s3, s4 = (torch.cuda.Stream() for _ in range(4))
def one():
    dst.copy_(src, non_blocking=True)
def split(k):
    streams = [s1, s2, s3, s4][:k]
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    step = n // k
    for i, s in enumerate(streams):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)
    for s in streams:
        cur.wait_stream(s)
for name, fn in (("1 copy", one), ("2 concurrent", lambda: split(2)), ("4 concurrent", lambda: split(4)), ("1 copy", one)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fn(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200
    print(f"{name}: {dt*1e6:.1f} us per 15.7 MB = {n*4/dt/1e9:.1f} GB/s")
