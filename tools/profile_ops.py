"""GPU diagnostic: per-op, per-kernel CUDA-event times for one KITTI-shaped pair (pyramid + 11 encoder blocks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import _native, blocks, dataloader, ops, synth
from apr_b200.architectures import KPFCNNEncoder
from apr_b200.config import kitti_config

dev = torch.device("cuda", 0)
cfg = kitti_config()
blocks.LINEAR_MODE = os.environ.get("LINEAR_MODE", "tf32")
a, b = synth.pair_raw(0)
raw = torch.from_numpy(np.concatenate([a, b])).to(dev); lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
p0, l0 = ops.grid_subsample(raw, lens, 0.3)
torch.manual_seed(0); np.random.seed(0)
enc = KPFCNNEncoder(cfg).to(dev).eval()
limits = [56, 55, 56, 58]

def run():
    return enc(dataloader.build_pyramid_device(p0, l0, cfg, limits))

for _ in range(3): run()
torch.cuda.synchronize()
rows = []
def wrap(name):
    f = getattr(ops, name)
    def g(*a, **k):
        r = f(*a, **k)
        rep = _native.prof_report()
        shp = [tuple(t.shape) for t in a if isinstance(t, torch.Tensor)][:4]
        rows.append((name, shp, rep))
        return r
    setattr(ops, name, g)
for n in ("grid_subsample", "radius_neighbors", "kpconv", "max_pool", "instnorm_lrelu", "linear_tf32"):
    wrap(n)
_native.prof_enable(True); _native.prof_report()
run()
_native.prof_enable(False)
tot = {}
for name, shp, rep in rows:
    ms = sum(v[1] for v in rep.values())
    tot[name] = tot.get(name, 0) + ms
    det = " ".join(f"{k.replace('_kernel','')}={v[1]*1e3:.0f}" for k, v in rep.items())
    print(f"{name:18s} {ms*1e3:7.0f} us  {shp}  | {det}")
print({k: round(v, 3) for k, v in tot.items()}, "total", round(sum(tot.values()), 3), "ms")
