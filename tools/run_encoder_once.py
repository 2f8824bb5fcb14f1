"""GPU helper for ncu: a few passes of pyramid + encoder on one KITTI-shaped pair (no timers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import blocks, dataloader, ops, synth
from apr_b200.architectures import KPFCNNEncoder
from apr_b200.config import kitti_config
dev = torch.device("cuda", 0)
cfg = kitti_config(); blocks.LINEAR_MODE = "tf32"
a, b = synth.pair_raw(0)
raw = torch.from_numpy(np.concatenate([a, b])).to(dev); lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
p0, l0 = ops.grid_subsample(raw, lens, 0.3)
torch.manual_seed(0); np.random.seed(0)
enc = KPFCNNEncoder(cfg).to(dev).eval()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    y = enc(dataloader.build_pyramid_device(p0, l0, cfg, [56, 55, 56, 58]))
torch.cuda.synchronize()
print("ok", tuple(y.shape))
