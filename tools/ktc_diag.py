"""GPU diagnostic for kpconv_tc_kernel: (1) wf vs the fp32 definition per 64-channel block, (2) time split by ktc_dbg bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import _native, dataloader, ops, synth
from apr_b200.config import kitti_config
dev = torch.device("cuda", 0)
setopt = lambda k, v: _native.check(_native.lib().aprb_set_option(k.encode(), v), "opt")
gen = torch.Generator().manual_seed(0)
for cin, h in ((64, 57), (128, 33), (256, 40))[:int(os.environ.get('KTC_CHECK', '0'))]:
    ns, nq = 2000, 1500
    s = torch.rand(ns, 3, generator=gen) * 2; q = torch.rand(nq, 3, generator=gen) * 2
    inds = torch.randint(0, ns + 1, (nq, h), generator=gen)
    x16 = torch.randn(ns, cin, generator=gen).half()
    kp = torch.randn(15, 3, generator=gen) * 0.4
    sp = torch.cat((s, torch.zeros(1, 3) + 1e6)); nb = sp[inds] - q.unsqueeze(1)
    w = torch.clamp(1 - torch.sqrt(((nb.unsqueeze(2) - kp) ** 2).sum(3)) / 0.7, min=0).transpose(1, 2)
    xz = torch.cat((x16.float(), torch.zeros(1, cin)))
    want = torch.matmul(w, xz[inds])                                  # [nq, 15, cin]
    for tc in (0, 1):
        setopt("kpconv_tc", tc)
        wf, inv = ops.kpconv_weighted_f16(q.to(dev), s.to(dev), inds.to(dev).int(), x16.to(dev), kp.to(dev), 0.7, layout_ck=bool(tc))
        got = wf.float().cpu().view(nq, cin, 16).permute(0, 2, 1)[:, :15] if tc else wf.float().cpu().view(nq, 15, cin)
        errs = [((got[:, :, c:c + 64] - want[:, :, c:c + 64]).norm() / want[:, :, c:c + 64].norm()).item() for c in range(0, cin, 64)]
        ek = [((got[:, k] - want[:, k]).norm() / want[:, k].norm().clamp_min(1e-9)).item() for k in (0, 7, 14)]
        print(f"Cin {cin} H {h} tc={tc}: rel err per 64-channel block {['%.1e' % e for e in errs]}  per k(0,7,14) {['%.1e' % e for e in ek]}")
if len(sys.argv) > 1 and sys.argv[1] == "time":
    cfg = kitti_config(); P = 8
    ps, ls = [], []
    for sd in range(P):
        a, b = synth.pair_raw(sd)
        raw = torch.from_numpy(np.concatenate([a, b])).to(dev); lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
        p0, l0 = ops.grid_subsample(raw, lens, 0.3); ps.append(p0); ls.append(l0)
    p0, l0 = torch.cat(ps).contiguous(), torch.cat(ls).contiguous()
    pyr = dataloader.build_pyramid_device(p0, l0, cfg, [57, 53, 54, 55])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    setopt("kpconv_tc", 1)
    for l, cin in ((0, 64), (1, 128), (2, 256)):
        q = s = pyr['points'][l]; inds = pyr['neighbors'][l]
        x = torch.randn(len(s), cin, generator=gen).half().to(dev)
        r = 0.3 * 4.25 * 2 ** l
        kp = (torch.randn(15, 3, generator=gen)); kp = (kp / kp.norm(dim=1, keepdim=True) * 0.66 * r).to(dev); kp[0] = 0
        for dbg in (0, 1, 2, 4, 32):
            setopt("ktc_dbg", dbg)
            ops.kpconv_weighted_f16(q, s, inds, x, kp, r * 2.0 / 4.25, layout_ck=True)
            _native.prof_enable(True); _native.prof_report()
            for _ in range(3):
                flush.zero_(); ops.kpconv_weighted_f16(q, s, inds, x, kp, r * 2.0 / 4.25, layout_ck=True)
            prof = _native.prof_report(); _native.prof_enable(False)
            print(f"L{l} C{cin} Nq {len(q)} dbg={dbg:2d}: " + " ".join(f"{k} {ms / 3 * 1e3:7.1f} us" for k, (c, ms) in prof.items()))
        setopt("ktc_dbg", 0)
