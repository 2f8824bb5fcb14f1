"""Runs the tcgen05 weighting kernel once per shape (L0 C64, L1 C128) on a super-batch of 8 KITTI-shaped pairs (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import _native, dataloader, ops, synth
from apr_b200.config import kitti_config
dev = torch.device("cuda", 0)
cfg = kitti_config(); P = 8
gen = torch.Generator().manual_seed(0)
ps, ls = [], []
for sd in range(P):
    a, b = synth.pair_raw(sd)
    raw = torch.from_numpy(np.concatenate([a, b])).to(dev); lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
    p0, l0 = ops.grid_subsample(raw, lens, 0.3); ps.append(p0); ls.append(l0)
p0, l0 = torch.cat(ps).contiguous(), torch.cat(ls).contiguous()
pyr = dataloader.build_pyramid_device(p0, l0, cfg, [57, 53, 54, 55])
for l, cin in ((0, 64), (1, 128)):
    q = s = pyr['points'][l]; inds = pyr['neighbors'][l]
    x = torch.randn(len(s), cin, generator=gen).half().to(dev)
    r = 0.3 * 4.25 * 2 ** l
    kp = (torch.randn(15, 3, generator=gen)); kp = (kp / kp.norm(dim=1, keepdim=True) * 0.66 * r).to(dev); kp[0] = 0
    for _ in range(2):
        ops.kpconv_weighted_f16(q, s, inds, x, kp, r * 2.0 / 4.25, layout_ck=True)
torch.cuda.synchronize()
print("ok")
