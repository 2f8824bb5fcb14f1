import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from apr_b200 import ops
from oracle import blocks_ref
cuda = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(5)
ns, nq, h, cin, cout = 300, 200, 19, 24, 40
s = torch.rand(ns, 3, generator=gen) * 2
q = torch.rand(nq, 3, generator=gen) * 2
inds = torch.randint(0, ns + 1, (nq, h + 5), generator=gen)
inds[:7] = ns
x = torch.randn(ns, cin, generator=gen)
x[::3] = -x[::3].abs()
x[5] = 0
kp = torch.randn(15, 3, generator=gen) * 0.4
w = torch.randn(15, cin, cout, generator=gen) * 0.1
view = inds[:, :h]
want = blocks_ref.kpconv_ref(q, s, view, x, kp, w, 0.7)
for it in range(6):
    got = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda)[:, :h], x.to(cuda), kp.to(cuda), w.to(cuda), 0.7, mode=1).cpu()
    d = (got - want).norm(dim=1) / want.norm(dim=1).clamp_min(1e-20)
    bad = torch.nonzero(d > 1e-5).flatten().tolist()
    print(it, "rel", ((got - want).norm() / want.norm()).item(), "bad rows", bad[:10], [round(d[i].item(), 6) for i in bad[:10]])
    for i in bad[:3]:
        r = (got[i] / want[i])
        print("   ratio row", i, r[:6].tolist())
sums = x.sum(1)
print("min |rowsum|", sums.abs().sort()[0][:5].tolist())
