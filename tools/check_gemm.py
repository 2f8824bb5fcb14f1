"""GPU diagnostic: aprb_linear_tf32 (tcgen05 TF32 GEMM) vs fp64 matmul on a few shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from apr_b200 import ops
torch.manual_seed(0)
dev = torch.device("cuda", 0)
from apr_b200 import _native
cl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
_native.check(_native.lib().aprb_set_option(b"gemm_cluster", cl)); print("gemm_cluster =", cl)
for n, cin, cout in [(128, 32, 64), (128, 64, 64), (256, 128, 128), (1000, 64, 128), (4255, 256, 64), (129, 32, 16),
                     (1567, 512, 2048), (35000, 960, 64), (1567, 7680, 512), (2051, 1024, 2048), (13995, 256, 512), (5471, 3840, 256), (13995, 1920, 128)]:
    x = torch.randn(n, cin, device=dev); w = torch.randn(cout, cin, device=dev) / cin ** 0.5
    y = ops.linear_tf32(x, w)
    torch.cuda.synchronize()
    ref = (x.double() @ w.double().t())
    err = ((y.double() - ref).norm() / ref.norm()).item()
    mx = (y.double() - ref).abs().max().item()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): ops.linear_tf32(x, w)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(f"{n}x{cin}x{cout}: rel {err:.3e} maxabs {mx:.3e}  {ms*1e3:.1f} us  {2*n*cin*cout/ms/1e9:.1f} TFLOP/s", flush=True)
