"""GPU diagnostic: the TF32 tcgen05 GEMM at the unary-Linear and KPConv-contraction shapes of the KFE encoder for a
super-batch of 8 KITTI-shaped pairs (level sizes 257k / 104k / 42k / 16k).  usage: gemm_bench.py [opt=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import _native, ops

for kv in sys.argv[1:]:
    k, v = kv.split("=")
    _native.check(_native.lib().aprb_set_option(k.encode(), int(v)), "aprb_set_option")
dev = torch.device("cuda", 0)
NL = [257031, 103717, 41645, 15924]
shapes = []
for l, n in enumerate(NL):
    d = 128 << l                      # block input width at this level (after the first block)
    shapes += [(f"L{l} unary1 {d}->{d//2}", n, d, d // 2), (f"L{l} unary2 {d//2}->{2*d}", n, d // 2, 2 * d),
               (f"L{l} shortcut {d}->{2*d}", n, d, 2 * d), (f"L{l} unary1' {2*d}->{d//2}", n, 2 * d, d // 2),
               (f"L{l} kpconv 15x{d//2}->{d//2}", n, 15 * (d // 2), d // 2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0.0
for name, n, cin, cout in shapes:
    x = torch.randn(n, cin, device=dev); w = torch.randn(cout, cin, device=dev)
    ops.linear_tf32(x, w)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear_tf32(x, w); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = float(np.median(ts)); tot += t
    print(f"{name:28s} M {n:7d} K {cin:5d} N {cout:5d}  {t:7.1f} us  {2*n*cin*cout/t*1e-6:6.1f} TFLOP/s  {4*(n*cin+n*cout)/t*1e-3:6.0f} GB/s")
print(f"total {tot:.0f} us")
