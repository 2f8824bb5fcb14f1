"""GPU diagnostic: where one KPFCNNPipeline.forward (BASELINE config 3, P stacked LoKITTI-shaped pairs) spends its time.
Prints (a) host time of the call vs device time (CUDA events) and (b) a torch.profiler table of the CUDA kernels and the CPU
ops of one call.   usage: kpfcnn_sections.py [P=8]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import blocks, ops, synth
from apr_b200.architectures import KPFCNN
from apr_b200.config import kitti_config
from apr_b200.pipeline import KPFCNNPipeline

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
cfg = kitti_config(); blocks.LINEAR_MODE = "tf32"
pts, lens = [], []
for sd in range(P):
    a, b = synth.pair_raw(sd, "kitti", True)
    raw = torch.from_numpy(np.concatenate([a, b])).to(dev); ln = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
    p0, l0 = ops.grid_subsample(raw, ln, 0.3)
    pts.append(p0); lens.append(l0)
p0, l0 = torch.cat(pts).contiguous(), torch.cat(lens).contiguous()
torch.manual_seed(0); np.random.seed(0)
net = KPFCNN(cfg).to(dev).eval()
pipe = KPFCNNPipeline(net, cfg, [57, 53, 54, 55], clouds_per_segment=2)
for _ in range(3):
    out = pipe.forward(p0, l0)
torch.cuda.synchronize()
hs, ds = [], []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(pipe.stream):
        e0.record()
        t0 = time.perf_counter()
        out = pipe.forward(p0, l0)
        t1 = time.perf_counter()
        e1.record()
    torch.cuda.synchronize()
    hs.append((t1 - t0) * 1e3); ds.append(e0.elapsed_time(e1))
print(f"P={P}: host time of the call {np.median(hs):.2f} ms, device time {np.median(ds):.2f} ms (points {p0.shape[0]})")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    out = pipe.forward(p0, l0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
