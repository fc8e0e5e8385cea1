"""PCIe probe (GPU box): pinned-host H2D alone, D2H alone, and both at once on two streams (252 MB each way — the
per-step transfer of the 1/12° Float64 host-buffer entry)."""
import torch, time
n = 252 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D alone {a:.2f} ms ({n/a/1e6:.1f} GB/s)  D2H alone {b:.2f} ms ({n/b/1e6:.1f} GB/s)  both {c:.2f} ms ({2*n/c/1e6:.1f} GB/s total)")
