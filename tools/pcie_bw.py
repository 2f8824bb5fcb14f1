"""GPU diagnostic: pinned D2H / H2D bandwidth of this box (the e2e number moves 643 MB of encoder output per step)."""
import torch
dev = torch.device("cuda", 0)
n = 256 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {8 * n / e0.elapsed_time(e1) * 1e-6:.1f} GB/s (8 x 256 MiB, pinned)")
