"""GPU diagnostic: KPConv in the native pipeline's mode (mode 4: fp16 features in, fp16 weighted tile, fp16 contraction)
at the ten shapes of the KFE encoder for a super-batch of P KITTI-shaped pairs; per-kernel CUDA-event times from the
library's own timers (aprb_prof_*), L2 flushed between repetitions, for each value of the given A/B option.
usage: kpconv_mode4_bench.py [P] [option] [values...]      e.g.  kpconv_mode4_bench.py 8 kpconv_tc 0 1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import _native, dataloader, ops, synth
from apr_b200.config import kitti_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
opt = sys.argv[2] if len(sys.argv) > 2 else "kpconv_tc"
vals = [int(v) for v in sys.argv[3:]] or [0, 1]
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
shapes = []
for l, c in enumerate((64, 128, 256, 512)):
    shapes.append((f"L{l} conv   C{c}", l, l, pyr['neighbors'][l], c, c))
    if l < 3:
        shapes.append((f"L{l} stride C{c}", l + 1, l, pyr['pools'][l], c, c))
REPS = 5
tot = {v: 0.0 for v in vals}
outs = {}
for name, ql, sl, inds, cin, cout in shapes:
    q, s = pyr['points'][ql], pyr['points'][sl]
    x = torch.nn.functional.leaky_relu(torch.randn(len(s), cin, generator=gen), 0.1).half().to(dev)
    kp = (torch.randn(15, 3, generator=gen) * 0.3).to(dev)
    r = cfg.first_subsampling_dl * cfg.conv_radius * 2 ** sl
    kp = kp / kp.norm(dim=1, keepdim=True).clamp_min(1e-6) * 0.66 * r; kp[0] = 0
    ext = r * cfg.KP_extent / cfg.conv_radius
    w = (torch.randn(15, cin, cout, generator=gen) / np.sqrt(15 * cin)).to(dev)
    prep = ops.kpconv_prepare_weights_f16(w)
    line = f"{name}: Nq {len(q):7d} Ns {len(s):7d}"
    for v in vals:
        _native.check(_native.lib().aprb_set_option(opt.encode(), v), "aprb_set_option")
        y = ops.kpconv(q, s, inds, x, kp, w, ext, wprep=prep, mode=4)
        outs.setdefault(name, []).append(y.clone())
        _native.prof_enable(True); _native.prof_report()
        for _ in range(REPS):
            flush.zero_()
            ops.kpconv(q, s, inds, x, kp, w, ext, wprep=prep, mode=4)
        prof = _native.prof_report(); _native.prof_enable(False)
        us = {k: ms / REPS * 1e3 for k, (cnt, ms) in prof.items()}
        t = sum(us.values()); tot[v] += t
        line += f" | {opt}={v}: " + " ".join(f"{k.replace('_kernel','')} {u:6.1f}" for k, u in sorted(us.items(), key=lambda kv: -kv[1])) + f" = {t:7.1f} us"
    if len(vals) > 1:
        a, b = outs[name][0].double(), outs[name][-1].double()
        line += f" | rel diff {((a - b).norm() / a.norm()).item():.1e}"
    print(line)
print("total per option (one launch of each of the 7 shapes; the encoder runs L0 conv once and the other conv shapes twice): " +
      ", ".join(f"{opt}={v}: {t:.0f} us" for v, t in tot.items()))
