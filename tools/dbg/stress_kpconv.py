"""GPU stress: random small KPConv problems, workspace poisoned with NaN before every call, fp32 path vs an fp64 torch
evaluation of the same formula on the device. Prints every mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from apr_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)

def ref(q, s, inds, x, kp, w, ext):
    q, s, x, kp, w = q.double(), s.double(), x.double(), kp.double(), w.double()
    sp = torch.cat((s, torch.full_like(s[:1], 1e6)), 0)
    nb = sp[inds] - q.unsqueeze(1)
    d = ((nb.unsqueeze(2) - kp) ** 2).sum(3).sqrt()
    wt = (1 - d / ext).clamp_min(0).transpose(1, 2)
    xz = torch.cat((x, torch.zeros_like(x[:1])), 0)
    nx = xz[inds]
    wf = wt @ nx
    out = torch.einsum('nkc,kco->no', wf, w)
    nn_ = (x.float().sum(1) > 0)
    nn_ = torch.cat((nn_, torch.zeros(1, dtype=torch.bool, device=x.device)))[inds].sum(1).clamp_min(1)
    return out / nn_.unsqueeze(1)

bad = 0
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for it in range(iters):
    g = torch.Generator(device="cpu").manual_seed(it)
    ns = int(torch.randint(50, 3000, (1,), generator=g)); nq = int(torch.randint(1, 3000, (1,), generator=g))
    h = int(torch.randint(1, 70, (1,), generator=g)); cin = [1, 3, 8, 24, 32, 64, 128, 256][it % 8]; cout = [16, 40, 64, 128][it % 4]
    s = torch.rand(ns, 3, generator=g) * 3; q = torch.rand(nq, 3, generator=g) * 3
    inds = torch.randint(0, ns + 1, (nq, h + 3), generator=g)
    x = torch.randn(ns, cin, generator=g); x[::3] = -x[::3].abs()
    kp = torch.randn(15, 3, generator=g) * 0.4; w = torch.randn(15, cin, cout, generator=g) * 0.1
    s, q, inds, x, kp, w = [t.to(dev) for t in (s, q, inds, x, kp, w)]
    for (key, buf) in ops._ws_cache.items():
        buf.view(torch.float32)[: buf.numel() // 4].fill_(float('nan'))
    view = inds[:, :h]
    want = ref(q, s, view, x, kp, w, 0.7)
    for mode in ((1, 2) if (15 * cin) % 32 == 0 and cout % 16 == 0 else (1,)):
        prep = ops.kpconv_prepare_weights(w) if mode == 2 else None
        got = ops.kpconv(q, s, view, x, kp, w, 0.7, wprep=prep, mode=mode).double()
        e = ((got - want).norm() / want.norm().clamp_min(1e-30)).item()
        tol = 2e-5 if mode == 1 else 2e-3
        if not (e < tol):
            bad += 1
            d = (got - want).norm(dim=1) / want.norm(dim=1).clamp_min(1e-20)
            rows = torch.nonzero(d > tol).flatten().tolist()
            print(f"it {it} mode {mode} ns {ns} nq {nq} h {h} cin {cin} cout {cout}: rel {e:.3e}, {len(rows)} bad rows {rows[:8]} nan={torch.isnan(got).any().item()}")
print("done, mismatches:", bad)
