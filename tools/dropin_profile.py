"""GPU diagnostic: where the strict drop-in path (numpy in / numpy out, one pair per step: bench.py `dropin_e2e`) spends its
time — the 13 cpp_wrappers-compatible calls of collate_fn_descriptor one by one, the H2D of the collated batch, the
module-path encoder, the D2H.   usage: dropin_profile.py [reps=5]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import blocks, dataloader as dl, ops, synth
from apr_b200.architectures import KPFCNNEncoder
from apr_b200.config import kitti_config
import apr_b200.cpp_wrappers.cpp_neighbors.radius_neighbors as cn
import apr_b200.cpp_wrappers.cpp_subsampling.grid_subsampling as cs

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
cfg = kitti_config(); blocks.LINEAR_MODE = "tf32"
a, b = synth.pair_raw(0)
raw = torch.from_numpy(np.concatenate([a, b])).to(dev); ln = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
p0, l0 = ops.grid_subsample(raw, ln, 0.3)
h = p0.cpu().numpy(); n0 = int(l0[0]); src, tgt = h[:n0].copy(), h[n0:].copy()
limits = [57, 53, 54, 55]
torch.manual_seed(0); np.random.seed(0)
enc = KPFCNNEncoder(cfg).to(dev).eval()

times = {}
def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); times[name] = times.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w
orig_q, orig_s = cn.batch_query, cs.subsample_batch
def run():
    t0 = time.perf_counter()
    batch = dl.collate_fn_descriptor(dl.make_list_data(src, tgt), cfg, limits)
    t1 = time.perf_counter()
    gpu = {k: ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v)
           for k, v in batch.items() if k in ("points", "neighbors", "pools", "upsamples", "stack_lengths")}
    gpu["features"] = batch["features"].to(dev)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    with torch.no_grad():
        y = enc(gpu)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    yh = y.cpu()
    t4 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2, t4 - t3
run(); run()
import apr_b200.dataloader as dmod
dmod.cpp_neighbors.batch_query = timed("batch_query (10 calls)", orig_q)
dmod.cpp_subsampling.subsample_batch = timed("subsample_batch (3 calls)", orig_s)
acc = np.zeros(4)
for _ in range(reps):
    acc += np.array(run())
acc /= reps
print(f"per pair: collate {acc[0]*1e3:.2f} ms | H2D of the batch {acc[1]*1e3:.2f} | encoder (module path) {acc[2]*1e3:.2f} | D2H {acc[3]*1e3:.2f} "
      f"| total {acc.sum()*1e3:.2f} ms = {2/acc.sum():.1f} clouds/s")
for k, v in times.items():
    print(f"   {k}: {v/reps*1e3:.2f} ms per pair")
# inside one batch_query: host conversions vs device work
q = np.concatenate([src, tgt]); lens = np.array([len(src), len(tgt)], np.int32)
for name, fn in (("batch_query L0 conv", lambda: orig_q(q, q, lens, lens, radius=1.275)),):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    print(f"   {name}: {(time.perf_counter()-t0)/reps*1e3:.2f} ms, out {out.shape} {out.dtype}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    orig_q(q, q, lens, lens, radius=1.275)
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=50))
