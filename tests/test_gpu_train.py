"""GPU parity of the training path (BASELINE config 5): gradients of our KPConv / max-pool autograd functions and of the
whole KPFCNN vs torch autograd through the fp32 CPU oracle (oracle/blocks_ref.py is plain torch, so autograd through it
IS the reference gradient). Tolerances: fp32 everywhere, scatter-add order not fixed -> 1e-4 relative (Frobenius)."""
import numpy as np
import pytest
import torch

from apr_b200 import ops, train
from apr_b200.architectures import KPFCNN
from apr_b200.config import kitti_config
from oracle import blocks_ref
from oracle.ref import collate_ref

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("cin,cout,h", [(1, 16, 20), (24, 40, 19), (64, 64, 35), (128, 32, 33)])
def test_kpconv_gradients_vs_oracle_autograd(cuda, cin, cout, h):
    gen = torch.Generator().manual_seed(11)
    ns, nq = 500, 350
    s = torch.rand(ns, 3, generator=gen) * 2.5
    q = torch.rand(nq, 3, generator=gen) * 2.5
    inds = torch.randint(0, ns + 1, (nq, h), generator=gen)
    inds[:5] = ns
    x = torch.randn(ns, cin, generator=gen)
    kp = torch.randn(15, 3, generator=gen) * 0.4
    w = torch.randn(15, cin, cout, generator=gen) * 0.1
    gout = torch.randn(nq, cout, generator=gen)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    yr = blocks_ref.kpconv_ref(q, s, inds, xr, kp, wr, 0.7)
    yr.backward(gout)
    xg, wg = x.to(cuda).requires_grad_(), w.to(cuda).requires_grad_()
    try:
        # fp32 contractions (torch.matmul), then the tcgen05 TF32 GEMMs (10-bit operand mantissa: the KPConv feature bar 1e-3)
        for tensor_gemm, tol_y, tol_g in ((False, 2e-5, 1e-4), (True, 1e-3, 1.5e-3)):
            train.TENSOR_GEMM = tensor_gemm
            for idx in (inds.to(cuda), inds.to(cuda).int()):
                xg.grad = wg.grad = None
                y = train._KPConvFn.apply(xg, wg, q.to(cuda), s.to(cuda), idx, kp.to(cuda), 0.7)
                y.backward(gout.to(cuda))
                print(f"Cin {cin} Cout {cout} tensor_gemm {tensor_gemm}: y {rel(y, yr):.1e} dx {rel(xg.grad, xr.grad):.1e} dW {rel(wg.grad, wr.grad):.1e}")
                assert rel(y, yr) < tol_y
                assert rel(xg.grad, xr.grad) < tol_g and rel(wg.grad, wr.grad) < tol_g
    finally:
        train.TENSOR_GEMM = True


def test_max_pool_backward_vs_oracle_autograd(cuda):
    gen = torch.Generator().manual_seed(12)
    x = torch.randn(400, 24, generator=gen)
    inds = torch.randint(0, 401, (150, 17), generator=gen)
    inds[:3] = 400
    gout = torch.randn(150, 24, generator=gen)
    xr = x.clone().requires_grad_()
    blocks_ref.max_pool_ref(xr, inds).backward(gout)
    xg = x.to(cuda).requires_grad_()
    train._MaxPoolFn.apply(xg, inds.to(cuda)).backward(gout.to(cuda))
    assert rel(xg.grad, xr.grad) < 1e-6


def test_kpfcnn_train_step_gradients_vs_oracle(cuda, oracle, gold_kpfcnn):
    """Full network: differentiable forward == the inference forward == the reference's golden outputs, and the
    parameter gradients of a scalar loss == autograd through the CPU restatement of the reference."""
    g = gold_kpfcnn
    cfg = kitti_config(first_feats_dim=16, gnn_feats_dim=32, final_feats_dim=8)
    pyr = collate_ref(g["p0"], g["l0"], cfg, list(g["limits"]), oracle.subsample_batch, oracle.batch_query)
    cpu = dict(points=[torch.from_numpy(p) for p in pyr["points"]], neighbors=[torch.from_numpy(n).long() for n in pyr["neighbors"]],
               pools=[torch.from_numpy(n).long() for n in pyr["pools"]], upsamples=[torch.from_numpy(n).long() for n in pyr["upsamples"]],
               stack_lengths=[torch.from_numpy(l) for l in pyr["stack_lengths"]], features=torch.ones(len(g["p0"]), 1))
    gpu = {k: ([t.to(cuda) for t in v] if isinstance(v, list) else v.to(cuda)) for k, v in cpu.items()}
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    net = KPFCNN(cfg)
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda)
    train.TENSOR_GEMM = False          # the whole-network gradient check runs the fp32 contractions (first_feats_dim = 16: most
    try:                               # KPConvs of this small golden network do not fit the tensor path's shape rules anyway)
        _check_kpfcnn_gradients(net, gpu, cpu, sd, cfg, g, cuda)
    finally:
        train.TENSOR_GEMM = True


def _check_kpfcnn_gradients(net, gpu, cpu, sd, cfg, g, cuda):
    ff, so, ss = train.kpfcnn_forward_train(net, gpu)
    for got, key in ((ff, "feats_f"), (so, "scores_overlap"), (ss, "scores_saliency")):
        assert rel(got, torch.from_numpy(g[key])) < 2e-4, key
    gen = torch.Generator().manual_seed(3)
    wf_, wo, ws_ = torch.randn(ff.shape, generator=gen), torch.randn(so.shape, generator=gen), torch.randn(ss.shape, generator=gen)
    loss = (ff * wf_.to(cuda)).sum() + (so * wo.to(cuda)).sum() + (ss * ws_.to(cuda)).sum()
    loss.backward()
    sdr = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'kernel_points' not in k) for k, v in sd.items()}
    rf, ro, rs = blocks_ref.kpfcnn_ref(cpu, sdr, cfg)
    ((rf * wf_).sum() + (ro * wo).sum() + (rs * ws_).sum()).backward()
    gmax = max(v.grad.norm().item() for v in sdr.values() if v.grad is not None)
    worst, checked, fails = 0.0, 0, []
    for name, p in net.named_parameters():
        if not p.requires_grad:
            continue
        ref = sdr[name].grad
        assert ref is not None and p.grad is not None, name
        # parameters whose gradient is analytically zero (a bias in front of an InstanceNorm) hold only rounding noise
        # on both sides: bound their absolute size instead of their ratio
        diff = (p.grad.detach().cpu().double() - ref.double()).norm().item()
        ok = diff < 2e-3 * ref.norm().item() + 1e-6 * gmax
        if ref.norm().item() > 1e-5 * gmax:
            worst = max(worst, diff / ref.norm().item()); checked += 1
        if not ok:
            fails.append((name, diff, ref.norm().item()))
    print(f"KPFCNN parameter gradients: {checked} tensors with non-trivial gradient, worst rel err {worst:.2e}")
    assert not fails, fails
    assert checked > 30


def test_train_step_runs_and_learns(cuda, oracle):
    """A few SGD steps of the config-5 objective (NPR generative loss + descriptor/score surrogate) on one small pair:
    finite losses, decreasing trend, all parameters updated through the bucketed reducer (world size 1 here)."""
    from apr_b200 import dataloader, synth
    cfg = kitti_config(first_feats_dim=32, gnn_feats_dim=64, final_feats_dim=32)
    a = synth.small_cloud(81, 2600)
    b = synth.small_cloud(81, 2500) - np.array([2.0, 0.0, 0.0], np.float32)
    p0, l0 = oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3)
    dp, dl_ = torch.from_numpy(p0).to(cuda), torch.from_numpy(l0).to(cuda)
    batch = dataloader.build_pyramid_device(dp, dl_, cfg, [30, 30, 30, 30])
    n_src = int(l0[0])
    src, tgt = dp[:n_src], dp[n_src:]
    shift = torch.tensor([2.0, 0.0, 0.0], device=cuda)
    nn_idx = ops.radius_neighbors(src, tgt + shift, dl_[:1], dl_[1:], 0.45, 1)       # correspondences within 0.45 m
    has = nn_idx[:, 0] < tgt.shape[0]
    corr = torch.stack([torch.nonzero(has).flatten(), nn_idx[has, 0].long()], 1)
    torch.manual_seed(0); np.random.seed(0)
    net, head = KPFCNN(cfg).to(cuda), train.NPRHead(cfg.final_feats_dim, cfg.point_generation_ratio).to(cuda)
    params = list(net.parameters()) + list(head.parameters())
    red = train.GradBucketReducer(params, bucket_mb=1.0)
    opt = torch.optim.SGD([p for p in params if p.requires_grad], lr=0.01, momentum=0.9, weight_decay=1e-6)
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        ff, so, ss = train.kpfcnn_forward_train(net, batch)
        loss = train.surrogate_desc_loss(ff[:n_src], ff[n_src:], corr, so, ss, n_src) \
            + train.npr_loss(head, ff[:n_src], src, src) + train.npr_loss(head, ff[n_src:], tgt, tgt)
        loss.backward()
        red.finish()
        opt.step()
        losses.append(loss.item())
    print("losses", [round(v, 4) for v in losses])
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_instnorm_lrelu_backward_vs_autograd(cuda):
    """aprb_instnorm_lrelu_backward (gradient of LeakyReLU(InstanceNorm(x)) from y and rstd) vs torch autograd in float64."""
    gen = torch.Generator().manual_seed(21)
    for n, c, slope in ((5000, 64, 0.1), (1567, 512, 0.1), (300, 7, 1.0), (40000, 128, 1.0), (2, 5, 0.1)):
        x = torch.randn(n, c, generator=gen) * 2 + 3
        g = torch.randn(n, c, generator=gen)
        xr = x.double().requires_grad_()
        yr = blocks_ref.instnorm_ref(xr)
        yr = yr if slope == 1.0 else torch.nn.functional.leaky_relu(yr, slope)
        yr.backward(g.double())
        xg = x.to(cuda).requires_grad_()
        y = train._NormActFn.apply(xg, slope)
        y.backward(g.to(cuda))
        assert rel(y, yr) < 1e-5
        assert rel(xg.grad, xr.grad) < 2e-4, (n, c, slope, rel(xg.grad, xr.grad))
