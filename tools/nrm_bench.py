"""A/B per encoder block shape: stored path (Linear -> fp32 product -> segmented InstanceNorm (+ shortcut) -> LeakyReLU)
vs recompute path (statistics pass -> per-segment mean/rstd -> contraction again with the normalisation in the epilogue).
  python tools/nrm_bench.py [pairs]      (CUDA events, 20 repetitions per shape, super-batch of `pairs` KITTI-sized pairs)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apr_b200 import _native  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ONLY = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None     # 1-based block numbers
REPS, WARM = int(os.environ.get("NRM_REPS", "20")), int(os.environ.get("NRM_WARM", "3"))
N = [int(x * P / 8) for x in (254972, 103738, 41805, 15854)]
# (name, rows, mid, cout, in_dim of the shortcut Linear or 0 = plain residual, fp16 output)
SHAPES = [("b1  L0 dual", N[0], 64, 256, 128, 1), ("b2  L1 res ", N[1], 64, 256, 0, 1), ("b3  L1 dual", N[1], 128, 512, 256, 1),
          ("b4  L1 res ", N[1], 128, 512, 0, 1), ("b5  L2 res ", N[2], 128, 512, 0, 1), ("b6  L2 dual", N[2], 256, 1024, 512, 1),
          ("b7  L2 res ", N[2], 256, 1024, 0, 1), ("b8  L3 res ", N[3], 256, 1024, 0, 1), ("b9  L3 dual", N[3], 512, 2048, 1024, 1),
          ("b10 L3 res ", N[3], 512, 2048, 0, 0)]
L = _native.lib()
for kv in os.environ.get("NRM_OPTS", "").split(","):                 # e.g. NRM_OPTS=nrm_park=0
    if "=" in kv:
        _native.check(L.aprb_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1])), "aprb_set_option")
ptr = _native.ptr; sp = _native.stream_ptr; C = _native.C
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
tot = [0.0, 0.0]
for bi, (name, n, mid, cout, cin_sc, out16) in enumerate(SHAPES):
    if ONLY and bi + 1 not in ONLY:
        continue
    S = P
    seg = torch.tensor([round(i * n / S) for i in range(S + 1)], dtype=torch.int32, device=dev)
    x = torch.randn(n, mid, generator=g).half().to(dev); w = (torch.randn(cout, mid, generator=g) / np.sqrt(mid)).half().to(dev)
    xs = ws = None
    if cin_sc:
        xs = torch.randn(n, cin_sc, generator=g).half().to(dev); ws = (torch.randn(cout, cin_sc, generator=g) / np.sqrt(cin_sc)).half().to(dev)
    res = None if cin_sc else torch.randn(n, cout, generator=g).half().to(dev)
    gb = int(L.aprb_group_stats_bytes(n, cout)) // 4
    y = torch.empty(n, cout, device=dev); y2 = torch.empty(n, cout, device=dev) if cin_sc else None
    g1 = torch.empty(gb, device=dev); g2 = torch.empty(gb, device=dev) if cin_sc else None
    wsn = torch.empty(int(L.aprb_instnorm_seg_ws_bytes(n, cout, S)), dtype=torch.uint8, device=dev)
    st = torch.empty(S * 2 * 2 * cout, device=dev)
    out = torch.empty(n, cout, dtype=torch.float16 if out16 else torch.float32, device=dev)
    wr = C.c_int(0)

    def old():
        _native.check(L.aprb_linear_f16_stats(ptr(x), ptr(w), n, mid, cout, ptr(y), ptr(g1), C.byref(wr), sp()))
        if cin_sc:
            _native.check(L.aprb_linear_f16_stats(ptr(xs), ptr(ws), n, cin_sc, cout, ptr(y2), ptr(g2), C.byref(wr), sp()))
        _native.check(L.aprb_instnorm_lrelu_seg_f16(ptr(y), n, cout, ptr(seg), S, 1e-5, 0.1, ptr(y2 if cin_sc else res), 0 if cin_sc else 1,
                                                    1 if cin_sc else 0, 1, ptr(out), out16, ptr(g1), ptr(g2), ptr(wsn), wsn.numel(), sp()))

    def new():
        _native.check(L.aprb_linear_f16_stats_ragged(ptr(x), ptr(w), n, mid, cout, ptr(y), ptr(g1), ptr(seg), S, sp()))
        if cin_sc:
            _native.check(L.aprb_linear_f16_stats_ragged(ptr(xs), ptr(ws), n, cin_sc, cout, ptr(y2), ptr(g2), ptr(seg), S, sp()))
        _native.check(L.aprb_instnorm_seg_stats(ptr(y), ptr(y2), n, cout, ptr(seg), S, 1e-5, ptr(g1), ptr(g2), ptr(st), sp()))
        _native.check(L.aprb_linear_f16_norm_apply(ptr(x), ptr(w), n, mid, cout, ptr(xs), ptr(ws), cin_sc, ptr(res), ptr(seg), S, ptr(st),
                                                   0.1, ptr(out), out16, sp()))

    ms = []
    for fn in (old, new):
        for _ in range(WARM):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(REPS):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / REPS * 1e3)
    tot[0] += ms[0]; tot[1] += ms[1]
    fl = 2.0 * n * cout * (mid + cin_sc) / 1e9
    print(f"{name} rows {n:7d} K {mid:4d}+{cin_sc:4d} -> {cout:4d}  {fl:6.1f} GFLOP   stored {ms[0]:7.1f} us   recompute {ms[1]:7.1f} us   {ms[0] / ms[1]:.2f}x")
print(f"sum: stored {tot[0]:.0f} us, recompute {tot[1]:.0f} us")
