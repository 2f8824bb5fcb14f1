"""GPU diagnostic: KPConv at the ten shapes of the KFE encoder for a super-batch of P KITTI-shaped pairs.
Times stage A+B alone (aprb_kpconv_weighted) and the whole operator (aprb_kpconv_forward, tensor path) with CUDA
events on torch's current stream, L2 flushed between repetitions.  usage: kpw_bench.py [P] [opt=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import _native, dataloader, ops, synth
from apr_b200.config import kitti_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    _native.check(_native.lib().aprb_set_option(k.encode(), int(v)), "aprb_set_option")
dev = torch.device("cuda", 0)
cfg = kitti_config()
ps, ls = [], []
for sd in range(P):
    a, b = synth.pair_raw(sd)
    raw = torch.from_numpy(np.concatenate([a, b])).to(dev); lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
    p0, l0 = ops.grid_subsample(raw, lens, 0.3)
    ps.append(p0); ls.append(l0)
p0, l0 = torch.cat(ps).contiguous(), torch.cat(ls).contiguous()
pyr = dataloader.build_pyramid_device(p0, l0, cfg, [57, 53, 54, 55])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device="cpu").manual_seed(0)
shapes = []   # (name, q level, s level, index matrix, Cin, Cout)
for l, c in enumerate((64, 128, 256, 512)):
    shapes.append((f"L{l} conv   C{c}", l, l, pyr['neighbors'][l], c, c))
    if l < 3:
        shapes.append((f"L{l} stride C{c}", l + 1, l, pyr['pools'][l], c, c))


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


tot_w = tot_f = 0.0
for name, ql, sl, inds, cin, cout in shapes:
    q, s = pyr['points'][ql], pyr['points'][sl]
    x = torch.randn(len(s), cin, generator=gen).to(dev)
    kp = (torch.randn(15, 3, generator=gen) * 0.3).to(dev)
    r = cfg.first_subsampling_dl * cfg.conv_radius * 2 ** sl
    kp = kp / kp.norm(dim=1, keepdim=True).clamp_min(1e-6) * 0.66 * r; kp[0] = 0
    ext = r * cfg.KP_extent / cfg.conv_radius
    w = (torch.randn(15, cin, cout, generator=gen) / np.sqrt(15 * cin)).to(dev)
    prep = ops.kpconv_prepare_weights(w)
    ops.kpconv_weighted(q, s, inds, x, kp, ext, round_tf32=True); ops.kpconv(q, s, inds, x, kp, w, ext, wprep=prep, mode=2)
    tw = timeit(lambda: ops.kpconv_weighted(q, s, inds, x, kp, ext, round_tf32=True))
    tf = timeit(lambda: ops.kpconv(q, s, inds, x, kp, w, ext, wprep=prep, mode=2))
    valid = (inds < len(s)).sum().item() / len(q)
    tot_w += tw; tot_f += tf
    print(f"{name}: Nq {len(q):7d} Ns {len(s):7d} H {inds.shape[1]} valid/row {valid:5.1f}  weighted {tw:7.1f} us  "
          f"kpconv {tf:7.1f} us  ({2*len(q)*15*cin*cout/tf*1e-6:6.1f} TFLOP/s, wf {len(q)*15*cin*4/tw*1e-3:6.0f} GB/s)")
print(f"total weighted {tot_w:.0f} us, kpconv {tot_f:.0f} us  (resnetb blocks use L0 conv x1, others x2: "
      f"see bench.py for the whole encoder)")
