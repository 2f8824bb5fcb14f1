"""GPU diagnostic: host-side cost of one aprb_kfe_forward call (enqueue only) vs its device time, 1..8 threads."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apr_b200 import blocks, dataloader, ops, synth
from apr_b200.architectures import KPFCNNEncoder
from apr_b200.config import kitti_config
from apr_b200.pipeline import KFEPipeline
dev = torch.device("cuda", 0); cfg = kitti_config(); blocks.LINEAR_MODE = "tf32"
a, b = synth.pair_raw(0)
raw = torch.from_numpy(np.concatenate([a, b])).to(dev); lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
p0, l0 = ops.grid_subsample(raw, lens, 0.3)
torch.manual_seed(0); np.random.seed(0)
enc = KPFCNNEncoder(cfg).to(dev).eval()
lim = [56, 55, 56, 58]
for S in (1, 2, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    pipes = [KFEPipeline(enc, cfg, lim, stream=streams[k]) for k in range(S)]
    for p in pipes: p.forward(p0, l0)
    torch.cuda.synchronize()
    host = [0.0] * S
    def work(k, reps):
        for _ in range(reps):
            t = time.perf_counter(); pipes[k].forward(p0, l0); host[k] += time.perf_counter() - t
    reps = 20
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k, reps)) for k in range(S)]
    [t.start() for t in th]; [t.join() for t in th]
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"S={S}: host time inside forward {1e3*sum(host)/(S*reps):.2f} ms/call; all enqueued after {1e3*t_enq:.1f} ms; "
          f"GPU done after {1e3*t_all:.1f} ms -> {1e3*t_all/(S*reps):.3f} ms/pair", flush=True)
