"""PCIe probe: pinned H2D, D2H and simultaneous bidirectional bandwidth (sets the ceiling of the e2e number)."""
import time, torch
n = 2 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(2 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {n/a/1e9:.1f} GB/s  D2H {2*n/b/1e9:.1f} GB/s  both: {c*1e3:.1f} ms for {n/1e9:.1f}+{2*n/1e9:.1f} GB -> H2D+D2H aggregate {(3*n)/c/1e9:.1f} GB/s (D2H-only time would be {b*1e3:.1f} ms)")
