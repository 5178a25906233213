"""PCIe copy rates of the box (pinned H2D / D2H / both at once): context for bench.py's e2e number."""
import torch, time
n = 64 << 20
h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    t = run(fn)
    print("%s: %.2f ms per 64 MiB -> %.1f GB/s%s" % (name, t * 1e3, n / t / 1e9, " per direction" if name == "both" else ""))
# many small copies (the bench copies ~100 tensors per step)
hs = [torch.empty(1 << 20, dtype=torch.uint8).pin_memory() for _ in range(64)]
ds = [torch.empty(1 << 20, dtype=torch.uint8, device="cuda") for _ in range(64)]
def small():
    with torch.cuda.stream(s1):
        for a, b in zip(ds, hs): a.copy_(b, non_blocking=True)
t = run(small)
print("64 x 1 MiB h2d: %.2f ms -> %.1f GB/s" % (t * 1e3, 64 * (1 << 20) / t / 1e9))
