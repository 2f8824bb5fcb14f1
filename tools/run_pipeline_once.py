"""GPU helper for ncu: a few super-batched passes (P pairs per call) of the native pipeline (no timers).
usage: run_pipeline_once.py [P=8] [iters=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import blocks, ops, synth
from apr_b200.architectures import KPFCNNEncoder
from apr_b200.config import kitti_config
from apr_b200.pipeline import KFEPipeline
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
cfg = kitti_config(); blocks.LINEAR_MODE = "tf32"
pts, lens = [], []
for sd in range(P):
    a, b = synth.pair_raw(sd)
    raw = torch.from_numpy(np.concatenate([a, b])).to(dev); ln = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
    p0, l0 = ops.grid_subsample(raw, ln, 0.3)
    pts.append(p0); lens.append(l0)
p0, l0 = torch.cat(pts).contiguous(), torch.cat(lens).contiguous()
torch.manual_seed(0); np.random.seed(0)
enc = KPFCNNEncoder(cfg).to(dev).eval()
pipe = KFEPipeline(enc, cfg, [57, 53, 54, 55], clouds_per_segment=2 if P > 1 else 0)
torch.cuda.synchronize()
print("setup done, launches so far", __import__("apr_b200._native", fromlist=["x"]).launch_count(), flush=True)
for _ in range(iters):
    y = pipe.forward(p0, l0)
torch.cuda.synchronize()
print("ok", tuple(y.shape), "points", p0.shape[0])
