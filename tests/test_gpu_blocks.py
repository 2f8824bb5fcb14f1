"""GPU parity: KPConv (K5), pooling gathers (K4), InstanceNorm(+LeakyReLU) (K6), the block modules and the KFE encoder
vs the reference golden vectors and the fp32 CPU oracle (oracle/blocks_ref.py).
Tolerances: fp32 CUDA-core path 2e-5 relative (Frobenius); tcgen05 TF32 path 1e-3 relative (TF32 operands rounded to
nearest, fp32 accumulation in TMEM) — the bar north_star states for KPConv features."""
import numpy as np
import pytest
import torch

from apr_b200 import blocks, ops
from apr_b200.architectures import KPFCNN, KPFCNNEncoder
from apr_b200.config import kitti_config
from oracle import blocks_ref
from oracle.ref import collate_ref

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-5
TOL_TF32 = 1e-3


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _t(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


@pytest.mark.parametrize("tag,strided", [("c1", False), ("c8", False), ("c32s", True), ("c64", False)])
@pytest.mark.parametrize("idx_dtype", [torch.int32, torch.int64])
def test_kpconv_golden_fp32(cuda, gold_kpconv, tag, strided, idx_dtype):
    g = gold_kpconv
    q, s, inds = (g["p1"], g["p0"], g["pool"]) if strided else (g["p0"], g["p0"], g["conv"])
    y = ops.kpconv(_t(q, cuda), _t(s, cuda), _t(inds, cuda).to(idx_dtype), _t(g[f"{tag}_x"], cuda),
                   _t(g[f"{tag}_kp"], cuda), _t(g[f"{tag}_W"], cuda), 0.6, mode=1)
    assert rel(y, torch.from_numpy(g[f"{tag}_y"])) < TOL_FP32


def test_kpconv_quirks_vs_oracle(cuda):
    """Shadow rows, neighbor_num on the SIGN of the feature sum, all-pad rows, non-contiguous index views."""
    gen = torch.Generator().manual_seed(5)
    ns, nq, h, cin, cout = 300, 200, 19, 24, 40
    s = torch.rand(ns, 3, generator=gen) * 2
    q = torch.rand(nq, 3, generator=gen) * 2
    inds = torch.randint(0, ns + 1, (nq, h + 5), generator=gen)
    inds[:7] = ns                                                         # queries with no neighbour at all
    x = torch.randn(ns, cin, generator=gen)
    x[::3] = -x[::3].abs()                                                # rows with negative sums (not counted)
    x[5] = 0
    kp = torch.randn(15, 3, generator=gen) * 0.4
    w = torch.randn(15, cin, cout, generator=gen) * 0.1
    view = inds[:, :h]                                                    # column-sliced view, like neighbors[:, :limit]
    want = blocks_ref.kpconv_ref(q, s, view, x, kp, w, 0.7)
    got = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda)[:, :h], x.to(cuda), kp.to(cuda), w.to(cuda), 0.7, mode=1)
    assert rel(got, want) < TOL_FP32
    assert torch.all(got[:7] == 0)


@pytest.mark.parametrize("cin,h,kp_scale", [(24, 19, 0.4), (64, 57, 0.4), (96, 70, 0.4), (128, 33, 0.4), (256, 40, 0.4),
                                            (512, 20, 0.4), (64, 40, 0.0), (36, 100, 0.01), (16, 140, 0.4)])
def test_kpconv_list_kernel_paths(cuda, cin, h, kp_scale):
    """The CSR-list producer (kp_weighted4_kernel): every lane-group configuration, partial channel slabs, 2 and 4
    neighbour groups per row, the H > 128 fallback, and rows whose list overflows (kp_scale 0: all 15 kernel points
    coincide, every neighbour inside the extent influences all of them -> direct path)."""
    gen = torch.Generator().manual_seed(cin * 1000 + h)
    ns, nq, cout = 700, 333, 32
    s = torch.rand(ns, 3, generator=gen) * 1.5
    q = torch.rand(nq, 3, generator=gen) * 1.5
    inds = torch.randint(0, ns + 1, (nq, h), generator=gen)
    inds[:3] = ns
    x = torch.randn(ns, cin, generator=gen)
    kp = torch.randn(15, 3, generator=gen) * kp_scale
    w = torch.randn(15, cin, cout, generator=gen) * 0.1
    want = blocks_ref.kpconv_ref(q, s, inds, x, kp, w, 0.7)
    got = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda).int(), x.to(cuda), kp.to(cuda), w.to(cuda), 0.7, mode=1)
    assert rel(got, want) < TOL_FP32
    got64 = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda), x.to(cuda), kp.to(cuda), w.to(cuda), 0.7, mode=1)
    assert torch.equal(got, got64)


def test_kpconv_tensor_path_vs_oracle(cuda, gold_kpconv):
    """tcgen05 TF32 contraction (mode 2) on shapes the tensor path accepts."""
    g = gold_kpconv
    gen = torch.Generator().manual_seed(9)
    p0, conv = torch.from_numpy(g["p0"]), torch.from_numpy(g["conv"]).long()
    for cin, cout in ((32, 32), (64, 64), (128, 128), (64, 256)):
        x = torch.randn(len(p0), cin, generator=gen)
        kp = torch.from_numpy(g["c8_kp"])
        w = torch.randn(15, cin, cout, generator=gen) / np.sqrt(15 * cin)
        want = blocks_ref.kpconv_ref(p0, p0, conv, x, kp, w, 0.6)
        wd = w.to(cuda)
        prep = ops.kpconv_prepare_weights(wd)
        got = ops.kpconv(p0.to(cuda), p0.to(cuda), conv.to(cuda).int(), x.to(cuda), kp.to(cuda), wd, 0.6, wprep=prep, mode=2)
        e = rel(got, want)
        print(f"kpconv tensor path Cin={cin} Cout={cout}: rel err {e:.2e}")
        assert e < TOL_TF32, f"Cin={cin} Cout={cout}: rel err {e:.2e}"


@pytest.mark.parametrize("cin,h,kp_scale,nq", [(64, 57, 0.4, 1000), (128, 33, 0.4, 517), (64, 20, 0.4, 129), (64, 40, 0.0, 300),
                                               (128, 64, 0.02, 260)])
def test_kpconv_fused_kernel_vs_unfused_and_oracle(cuda, cin, h, kp_scale, nq):
    """kpconv_fused_kernel (gather -> influence -> swizzled A tiles in shared memory -> tcgen05, no [Nq, K*Cin]
    intermediate) against the two-kernel tensor path and the fp32 oracle: ragged last tile, 1 and 2 neighbour groups,
    pads inside rows, all-pad rows, rows whose influence list overflows the shared-memory slot (kp_scale ~ 0: every
    neighbour influences all 15 kernel points -> direct path), int64 indices, and the epilogue's group statistics."""
    from apr_b200 import _native
    gen = torch.Generator().manual_seed(cin * 7 + h)
    ns = 900
    s = torch.rand(ns, 3, generator=gen) * 1.5
    q = torch.rand(nq, 3, generator=gen) * 1.5
    inds = torch.randint(0, ns + 1, (nq, h), generator=gen)
    inds[:3] = ns
    inds[5, : h // 2] = ns
    x = torch.randn(ns, cin, generator=gen)
    kp = torch.randn(15, 3, generator=gen) * kp_scale
    w = torch.randn(15, cin, cin, generator=gen) / np.sqrt(15 * cin)
    want = blocks_ref.kpconv_ref(q, s, inds, x, kp, w, 0.7)
    dev = [t.to(cuda) for t in (q, s)]
    wd, xd, kpd = w.to(cuda), x.to(cuda), kp.to(cuda)
    prep = ops.kpconv_prepare_weights(wd)
    setopt = lambda v: _native.check(_native.lib().aprb_set_option(b"kpconv_fused", v), "aprb_set_option")
    try:
        setopt(0)
        two = ops.kpconv(*dev, inds.to(cuda).int(), xd, kpd, wd, 0.7, wprep=prep, mode=2)
        setopt(1)
        launches = _native.launch_count()
        one = ops.kpconv(*dev, inds.to(cuda).int(), xd, kpd, wd, 0.7, wprep=prep, mode=2)
        assert _native.launch_count() - launches == 2            # rowsum/pack + the fused kernel
        one64 = ops.kpconv(*dev, inds.to(cuda), xd, kpd, wd, 0.7, wprep=prep, mode=2)
    finally:
        setopt(0)                                                     # library default (see profiles/r01_kpconv_fused.txt)
    assert torch.equal(one, one64)
    assert torch.all(one[:3] == 0)
    # the fused producer keeps 17 mantissa bits of each influence weight (2^-18 relative): a few A elements then round to
    # the neighbouring TF32 value (2^-11), hence 1e-4 rather than round-off between the two tensor paths
    assert rel(one, two) < 1e-4, rel(one, two)
    assert rel(one, want) < TOL_TF32, rel(one, want)
    if nq >= 32:                                                      # group statistics written by the fused epilogue
        gs = ops.group_stats(one).view(-1, 2, cin)
        g = nq // 32
        blk = one[: g * 32].view(g, 32, cin)
        assert rel(gs[:g, 0], blk.mean(1)) < 1e-5
        assert rel(gs[:g, 1], ((blk - blk.mean(1, keepdim=True)) ** 2).sum(1)) < 1e-4


def test_kpconv_fp16_operand_path_vs_oracle(cuda, gold_kpconv):
    """mode 3: the weighted tile and the weights in fp16 (10-bit mantissa like TF32), fp32 accumulation in TMEM."""
    g = gold_kpconv
    gen = torch.Generator().manual_seed(19)
    p0, conv = torch.from_numpy(g["p0"]), torch.from_numpy(g["conv"]).long()
    for cin, cout in ((64, 64), (128, 128), (64, 256), (256, 256), (36, 48)):
        x = torch.randn(len(p0), cin, generator=gen)
        kp = torch.from_numpy(g["c8_kp"])
        w = torch.randn(15, cin, cout, generator=gen) / np.sqrt(15 * cin)
        want = blocks_ref.kpconv_ref(p0, p0, conv, x, kp, w, 0.6)
        wd = w.to(cuda)
        if not ops.kpconv_f16_supported(15, cin, cout, conv.shape[1]):
            assert (15 * cin) % 64 != 0
            continue
        got = ops.kpconv(p0.to(cuda), p0.to(cuda), conv.to(cuda).int(), x.to(cuda), kp.to(cuda), wd, 0.6,
                         wprep=ops.kpconv_prepare_weights_f16(wd), mode=3)
        e = rel(got, want)
        print(f"kpconv fp16 operands Cin={cin} Cout={cout}: rel err {e:.2e}")
        assert e < TOL_TF32, f"Cin={cin} Cout={cout}: rel err {e:.2e}"
        assert ops.group_stats(got) is not None


def test_linear_tf32_vs_fp32(cuda):
    gen = torch.Generator().manual_seed(2)
    for n, cin, cout in ((1000, 64, 128), (4255, 256, 64), (129, 32, 16), (1567, 512, 2048)):
        x = torch.randn(n, cin, generator=gen)
        w = torch.randn(cout, cin, generator=gen) / np.sqrt(cin)
        want = x.double() @ w.double().t()
        got = ops.linear_tf32(x.to(cuda), w.to(cuda))
        e = rel(got, want)
        assert e < TOL_TF32, f"{n}x{cin}x{cout}: {e:.2e}"


def test_pools_golden_bit_exact(cuda, gold_kpconv):
    g = gold_kpconv
    for dt in (torch.int32, torch.int64):
        y = ops.max_pool(_t(g["pool_x"], cuda), _t(g["pool"], cuda).to(dt))
        assert torch.equal(y.cpu(), torch.from_numpy(g["max_pool_y"]))
        y = ops.closest_pool(_t(g["closest_x"], cuda), _t(g["up"], cuda).to(dt))
        assert torch.equal(y.cpu(), torch.from_numpy(g["closest_pool_y"]))
    # all-negative neighbourhood with a shadow entry pools to 0 (the zero row takes part in the max, blocks.py:95-101)
    x = -torch.rand(10, 6) - 1
    inds = torch.tensor([[0, 1, 10], [2, 3, 4]])
    y = ops.max_pool(x.to(cuda), inds.to(cuda)).cpu()
    assert torch.all(y[0] == 0) and torch.equal(y[1], x[2:5].max(0)[0])
    # odd channel count -> scalar kernel
    x = torch.randn(50, 7); inds = torch.randint(0, 51, (20, 9))
    assert torch.equal(ops.max_pool(x.to(cuda), inds.to(cuda)).cpu(), blocks_ref.max_pool_ref(x, inds))
    # multiples of 128 channels -> warp-per-query kernel (pads anywhere in the row, all-pad rows, H > 32, int64 indices)
    gen = torch.Generator().manual_seed(3)
    for c, h in ((128, 9), (256, 56), (512, 33), (1024, 70), (384, 5)):
        x = torch.randn(300, c, generator=gen)
        inds = torch.randint(0, 301, (77, h), generator=gen)
        inds[3] = 300
        inds[4, : h // 2] = 300
        want = blocks_ref.max_pool_ref(x, inds)
        assert torch.equal(ops.max_pool(x.to(cuda), inds.to(cuda).int()).cpu(), want)
        assert torch.equal(ops.max_pool(x.to(cuda), inds.to(cuda)).cpu(), want)
        assert torch.equal(ops.max_pool(-x.abs().to(cuda), inds.to(cuda).int()).cpu(), blocks_ref.max_pool_ref(-x.abs(), inds))


def test_instnorm_lrelu_vs_torch(cuda):
    gen = torch.Generator().manual_seed(3)
    for n, c in ((1567, 512), (27782, 64), (300, 7), (1, 5)):
        x = torch.randn(n, c, generator=gen) * 3 + 50                      # large mean: stresses the variance formula
        want = torch.nn.functional.leaky_relu(blocks_ref.instnorm_ref(x.double()), 0.1).float() if n > 1 else None
        got = ops.instnorm_lrelu(x.to(cuda), slope=0.1).cpu()
        if n > 1:
            assert (got - want).abs().max() < 2e-4
        r = torch.randn(n, c, generator=gen)
        got = ops.instnorm_lrelu(x.to(cuda), slope=1.0, residual=r.to(cuda)).cpu()
        if n > 1:
            assert (got - (blocks_ref.instnorm_ref(x.double()) + r).float()).abs().max() < 2e-4
            got = ops.instnorm_lrelu(x.to(cuda), slope=0.1, residual=r.to(cuda), norm_residual=True).cpu()
            want = torch.nn.functional.leaky_relu(blocks_ref.instnorm_ref(x.double()) + blocks_ref.instnorm_ref(r.double()), 0.1)
            assert (got - want.float()).abs().max() < 2e-4
    # same semantics as the reference module (nn.InstanceNorm1d on [1,C,N])
    x = torch.randn(500, 16, generator=gen)
    ref = torch.nn.InstanceNorm1d(16)(x.unsqueeze(2).transpose(0, 2)).transpose(0, 2).squeeze()
    assert (ops.instnorm_lrelu(x.to(cuda), slope=1.0).cpu() - ref).abs().max() < 1e-5


def test_instnorm_segmented_vs_torch(cuda):
    """Segmented InstanceNorm (super-batched pairs): every segment equals a separate BatchNormBlock call."""
    gen = torch.Generator().manual_seed(4)
    for lens, c in (([700, 650, 800, 30, 1, 900], 64), ([5000, 4000], 128), ([35000, 31000, 33000, 36000], 64), ([300], 2048)):
        n = sum(lens)
        x = torch.randn(n, c, generator=gen) * 2 + torch.linspace(-30, 30, n).unsqueeze(1)   # segment-dependent means
        r = torch.randn(n, c, generator=gen)
        off = ops.segment_offsets(torch.tensor(lens, dtype=torch.int32, device=cuda), 1)
        assert off.cpu().tolist() == [0] + list(np.cumsum(lens))
        got = ops.instnorm_lrelu_seg(x.to(cuda), off, slope=0.1).cpu()
        got_r = ops.instnorm_lrelu_seg(x.to(cuda), off, slope=0.1, residual=r.to(cuda), norm_residual=True).cpu()
        got_p = ops.instnorm_lrelu_seg(x.to(cuda), off, slope=1.0, residual=r.to(cuda)).cpu()
        a = 0
        for ln in lens:
            xs, rs = x[a:a + ln].double(), r[a:a + ln].double()
            if ln > 1:
                want = torch.nn.functional.leaky_relu(blocks_ref.instnorm_ref(xs), 0.1).float()
                assert (got[a:a + ln] - want).abs().max() < 2e-4
                want = torch.nn.functional.leaky_relu(blocks_ref.instnorm_ref(xs) + blocks_ref.instnorm_ref(rs), 0.1).float()
                assert (got_r[a:a + ln] - want).abs().max() < 2e-4
                assert (got_p[a:a + ln] - (blocks_ref.instnorm_ref(xs) + rs).float()).abs().max() < 2e-4
            a += ln
    # two clouds per segment (a collated pair), and the unsegmented form == the three-launch kernel's result
    off = ops.segment_offsets(torch.tensor([10, 20, 30, 40, 50], dtype=torch.int32, device=cuda), 2)
    assert off.cpu().tolist() == [0, 30, 100, 150]
    x = torch.randn(4000, 256, generator=gen) + 7
    a, b = ops.instnorm_lrelu_seg(x.to(cuda), None, slope=0.1).cpu(), ops.instnorm_lrelu(x.to(cuda), slope=0.1).cpu()
    assert (a - b).abs().max() < 1e-5


def _pyramid(oracle, cfg, p0, l0, limits, cuda):
    pyr = collate_ref(p0, l0, cfg, limits, oracle.subsample_batch, oracle.batch_query)
    cpu = dict(points=[torch.from_numpy(p) for p in pyr["points"]],
               neighbors=[torch.from_numpy(n).long() for n in pyr["neighbors"]],
               pools=[torch.from_numpy(n).long() for n in pyr["pools"]],
               upsamples=[torch.from_numpy(n).long() for n in pyr["upsamples"]],
               features=torch.ones(len(p0), 1))
    gpu = {k: ([t.to(cuda) for t in v] if isinstance(v, list) else v.to(cuda)) for k, v in cpu.items()}
    return cpu, gpu


def test_instnorm_from_gemm_group_stats(cuda):
    """Linear on tcgen05 whose epilogue records per-32-row-group column statistics, consumed by the segmented
    InstanceNorm instead of its own pass over the tensor: equal to the scanning path up to fp32 reassociation, for
    segment bounds that are unaligned to the groups, shorter than a group, and empty."""
    gen = torch.Generator().manual_seed(11)
    for n, cin, cout, bounds in ((5000, 64, 256, [0, 1234, 1250, 1250, 3001, 5000]), (700, 128, 64, [0, 700]),
                                 (4097, 32, 128, [0, 31, 64, 4097]), (40, 64, 64, [0, 7, 40])):
        x = (torch.randn(n, cin, generator=gen) * 2 + 0.5).to(cuda)
        w = (torch.randn(cout, cin, generator=gen) / np.sqrt(cin)).to(cuda)
        r = (torch.randn(n, cout, generator=gen)).to(cuda)
        seg = torch.tensor(bounds, dtype=torch.int32, device=cuda)
        ops.FUSE_STATS = True
        try:
            y = ops.linear_tf32(x, w)
            wrote = ops.group_stats(y) is not None
            a = ops.instnorm_lrelu_seg(y, seg, slope=0.1, residual=r, norm_residual=False)
            y2 = ops.linear_tf32(x, w)
            b2 = ops.instnorm_lrelu_seg(y, seg, slope=0.1, residual=y2, norm_residual=True)
        finally:
            ops.FUSE_STATS = True
        yc = y.clone()                                               # a clone carries no statistics: scanning path
        b = ops.instnorm_lrelu_seg(yc, seg, slope=0.1, residual=r, norm_residual=False)
        b3 = ops.instnorm_lrelu_seg(yc, seg, slope=0.1, residual=y2.clone(), norm_residual=True)
        assert wrote or n < 128 * 148                                # split-K shapes legitimately skip the statistics
        assert rel(a, b) < 2e-6, (n, cin, cout, rel(a, b))
        assert rel(b2, b3) < 2e-6, (n, cin, cout, rel(b2, b3))
        # and against torch, segment by segment
        for s0, s1 in zip(bounds[:-1], bounds[1:]):
            if s1 > s0:
                ys = yc[s0:s1]
                ref = (ys - ys.mean(0)) / torch.sqrt(ys.var(0, unbiased=False) + 1e-5) + r[s0:s1]
                ref = torch.nn.functional.leaky_relu(ref, 0.1)
                assert rel(a[s0:s1], ref) < 1e-5


@pytest.mark.parametrize("mode,tol", [(1, 5e-5), (0, 2e-3)])
def test_encoder_golden_small(cuda, oracle, gold_encoder, mode, tol):
    """Whole KFE encoder (first_feats_dim=16) vs the REAL reference's output; weights loaded through load_state_dict
    with the reference's own keys."""
    g = gold_encoder
    cfg = kitti_config(first_feats_dim=16)
    cpu, gpu = _pyramid(oracle, cfg, g["p0"], g["l0"], list(g["limits"]), cuda)
    enc = KPFCNNEncoder(cfg)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    missing, unexpected = enc.load_state_dict(sd, strict=True), None
    enc = enc.to(cuda).eval()
    blocks.KPCONV_MODE = mode
    try:
        y = enc(gpu)
    finally:
        blocks.KPCONV_MODE = 0
    assert y.shape == g["y_final"].shape
    assert rel(y, torch.from_numpy(g["y_final"])) < tol


@pytest.mark.parametrize("linear_mode", ["fp32", "tf32"])
def test_blocks_vs_oracle_kitti_width(cuda, oracle, linear_mode):
    """Each block type at KITTI channel widths, fed IDENTICAL inputs on both sides (per-block parity, tensor path on;
    unary Linear either cuBLAS fp32 or our tcgen05 TF32 GEMM)."""
    from apr_b200 import synth
    cfg = kitti_config()
    a, b = synth.small_cloud(41, 2600), synth.small_cloud(42, 2400)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p0, l0 = oracle.subsample_batch(raw, lens, sampleDl=0.3)
    cpu, gpu = _pyramid(oracle, cfg, p0, l0, [35, 35, 35, 35], cuda)
    torch.manual_seed(0); np.random.seed(0)
    gen = torch.Generator().manual_seed(1)
    cases = [("simple", 1, 128, 0), ("resnetb", 64, 128, 0), ("resnetb_strided", 128, 128, 0), ("resnetb", 128, 256, 1),
             ("resnetb_strided", 256, 256, 1), ("resnetb", 512, 1024, 2), ("resnetb", 2048, 2048, 3)]
    worst = 0.0
    for name, cin, cout, layer in cases:
        r = cfg.first_subsampling_dl * cfg.conv_radius * 2 ** layer
        blk = blocks.block_decider(name, r, cin, cout, layer, cfg)
        sd = {"encoder_blocks.0." + k: v.detach().clone() for k, v in blk.state_dict().items()}
        x = torch.ones(len(cpu["points"][layer]), 1) if cin == 1 else torch.randn(len(cpu["points"][layer]), cin, generator=gen)
        extent = r * cfg.KP_extent / cfg.conv_radius
        if "simple" in name:
            want = blocks_ref.simple_ref(x, cpu, sd, "encoder_blocks.0.", name, layer, extent)
        else:
            want = blocks_ref.resnetb_ref(x, cpu, sd, "encoder_blocks.0.", name, layer, extent)
        blocks.LINEAR_MODE = linear_mode
        try:
            with torch.no_grad():
                got = blk.to(cuda)(x.to(cuda), gpu)
        finally:
            blocks.LINEAR_MODE = 'fp32'
        e = rel(got, want)
        print(f"{linear_mode} {name} {cin}->{cout}: rel err {e:.2e}")
        worst = max(worst, e)
        assert e < TOL_TF32, f"{name} {cin}->{cout}: rel err {e:.2e}"
    print(f"worst per-block rel err {worst:.2e}")


def test_native_pipeline_matches_module_path(cuda, oracle):
    """aprb_kfe_forward (C++ driver: pyramid + encoder in one call) == Python module path, and its pyramid == oracle."""
    from apr_b200 import dataloader, synth
    from apr_b200.pipeline import KFEPipeline
    cfg = kitti_config()
    a, b = synth.small_cloud(51, 3000), synth.small_cloud(52, 2600)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p0, l0 = oracle.subsample_batch(raw, lens, sampleDl=0.3)
    limits = [30, 31, 32, 33]
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    dp, dl_ = _t(p0, cuda), _t(l0, cuda)
    blocks.LINEAR_MODE = 'tf32'
    try:
        want = enc(dataloader.build_pyramid_device(dp, dl_, cfg, limits))
    finally:
        blocks.LINEAR_MODE = 'fp32'
    pipe = KFEPipeline(enc, cfg, limits)
    got = pipe.forward(dp, dl_).clone()                             # the result is a view into the arena: copy it
    torch.cuda.synchronize()
    assert got.shape == want.shape
    # The native path keeps normalised activations in fp16 (exact: they are TF32-rounded) and runs every contraction with
    # kind::f16; the module path keeps them in fp32 and runs the Linears with kind::tf32. Same operand values, different
    # accumulation grouping inside the tensor core (K = 16 vs 8 per instruction): fp32 round-off, which a TF32 re-rounding
    # of a stored activation occasionally turns into 2^-11 — the same drift class as super-batched vs single pairs.
    e_paths = rel(got, want)
    print(f"native (fp16 activations) vs module path (fp32 activations): {e_paths:.2e}")
    assert e_paths < 5e-3
    from apr_b200 import _native
    try:                                                            # fp32 activations + TF32 Linears natively: bit-identical
        _native.check(_native.lib().aprb_set_option(b"act_f16", 0), "aprb_set_option")
        got32 = pipe.forward(dp, dl_).clone()
    finally:
        _native.check(_native.lib().aprb_set_option(b"act_f16", 1), "aprb_set_option")
    assert torch.equal(got32, want)                                 # same kernels, same order -> bit-identical
    ref = collate_ref(p0, l0, cfg, limits, oracle.subsample_batch, oracle.batch_query)
    pyr = pipe.pyramid()
    for k in ("points", "neighbors", "pools", "upsamples"):
        for lvl in range(4):
            w = ref[k][lvl]
            d = pyr[k][lvl].cpu().numpy()
            if w.shape[0] == 0:
                assert d.shape[0] == 0
            elif k == "points":
                assert np.array_equal(d, w)
            else:
                assert np.array_equal(d[:, :w.shape[1]], w), (k, lvl)
    # host-buffer entry point (H2D + path + D2H)
    hp, hl = torch.from_numpy(p0).pin_memory(), torch.from_numpy(l0).pin_memory()
    host = pipe.forward_host(hp, hl)
    assert torch.equal(host, got.cpu())
    # asynchronous host entry point: two calls in flight, results awaited one call later, ping-pong device/host buffers
    bufs = [torch.empty((got.shape[0] + 7, got.shape[1]), dtype=torch.float32).pin_memory() for _ in range(2)]
    pending = []
    for c in range(5):
        y, ticket = pipe.forward_host_async(hp, hl, bufs[c & 1])
        if pending:
            pipe.wait_host(pending[1])
            assert torch.equal(pending[0], got.cpu())
        pending = (y, ticket)
    pipe.wait_host(pending[1])
    assert torch.equal(pending[0], got.cpu())
    # fp16 host output: the final activation is already rounded to a 10-bit mantissa -> same values, half the bytes
    bufs16 = [torch.empty((got.shape[0] + 7, got.shape[1]), dtype=torch.float16).pin_memory() for _ in range(2)]
    for c in range(3):
        y16, ticket = pipe.forward_host_async(hp, hl, bufs16[c & 1])
        pipe.wait_host(ticket)
        assert torch.equal(y16, got.cpu().half())
        big = got.cpu().abs() >= 2.0 ** -14
        assert torch.equal(y16.float()[big], got.cpu()[big])
    y32, ticket = pipe.forward_host_async(hp, hl, bufs[0])           # and back to fp32 on the same handle
    pipe.wait_host(ticket)
    assert torch.equal(y32, got.cpu())
    # vs the fp32 CPU oracle end to end (drift over 11 blocks, TF32 tensor path): reported, loose bound
    cpu = dict(points=[torch.from_numpy(p) for p in ref["points"]], neighbors=[torch.from_numpy(n).long() for n in ref["neighbors"]],
               pools=[torch.from_numpy(n).long() for n in ref["pools"]], features=torch.ones(len(p0), 1))
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    y = blocks_ref.encoder_ref(cpu, sd, cfg)
    e = rel(got, y)
    print(f"end-of-encoder drift vs fp32 oracle (11 blocks, TF32): {e:.2e}")
    assert e < 1e-2


def test_linear_norm_apply_recompute_equals_three_kernel_sequence(cuda):
    """Recompute path of the block-closing Linears (aprb_linear_f16_stats_ragged -> aprb_instnorm_seg_stats ->
    aprb_linear_f16_norm_apply: the fp32 product is never materialised) == Linear -> segmented InstanceNorm (+ plain
    fp16 residual | + standardised shortcut product) -> LeakyReLU through the stored fp32 tensor. Same accumulators; the
    group statistics are summed in a different order and the shortcut is added unfused (as the reference does), so
    the two agree up to fp32 round-off, which the 10-bit rounding of the stored activation turns into a rare one-ulp
    (2^-10) difference. Segment bounds unaligned to the 32-row groups, shorter than a group, empty; fp16 and fp32
    outputs. The product buffer is poisoned with NaN first: only the ragged rows may be read."""
    from apr_b200 import _native
    L = _native.lib()
    P, sp = _native.ptr, _native.stream_ptr
    gen = torch.Generator().manual_seed(5)
    cases = ((5000, 64, 256, 128, [0, 1234, 1250, 1250, 3001, 5000]), (700, 128, 512, 256, [0, 700]),
             (4097, 64, 128, 64, [0, 31, 64, 4097]), (40, 64, 64, 64, [0, 7, 40]), (20000, 256, 1024, 512, [0, 9000, 20000]))
    for n, cin, cout, csc, bounds in cases:
        S = len(bounds) - 1
        x = torch.nn.functional.leaky_relu(torch.randn(n, cin, generator=gen), 0.1).half().to(cuda)
        w = (torch.randn(cout, cin, generator=gen) / np.sqrt(cin)).half().to(cuda)
        xs = torch.randn(n, csc, generator=gen).half().to(cuda)
        ws = (torch.randn(cout, csc, generator=gen) / np.sqrt(csc)).half().to(cuda)
        res = torch.randn(n, cout, generator=gen).half().to(cuda)
        seg = torch.tensor(bounds, dtype=torch.int32, device=cuda)
        gbytes = int(L.aprb_group_stats_bytes(n, cout))
        ws_norm = torch.empty(int(L.aprb_instnorm_seg_ws_bytes(n, cout, S)), dtype=torch.uint8, device=cuda)

        def old(variant, out16):
            y = torch.empty(n, cout, device=cuda); g = torch.empty(gbytes // 4, device=cuda)
            wr = _native.C.c_int(0)
            _native.check(L.aprb_linear_f16_stats(P(x), P(w), n, cin, cout, P(y), P(g), _native.C.byref(wr), sp()), "linear")
            assert wr.value == 1
            y2 = g2 = None
            if variant == "dual":
                y2 = torch.empty(n, cout, device=cuda); g2 = torch.empty(gbytes // 4, device=cuda)
                _native.check(L.aprb_linear_f16_stats(P(xs), P(ws), n, csc, cout, P(y2), P(g2), _native.C.byref(wr), sp()), "linear")
            out = torch.empty(n, cout, dtype=torch.float16 if out16 else torch.float32, device=cuda)
            r = {"dual": y2, "res": res, "none": None}[variant]
            _native.check(L.aprb_instnorm_lrelu_seg_f16(P(y), n, cout, P(seg), S, 1e-5, 0.1, P(r), 1 if variant == "res" else 0,
                                                        1 if variant == "dual" else 0, 1, P(out), out16, P(g), P(g2),
                                                        P(ws_norm), ws_norm.numel(), sp()), "norm")
            return out

        GUARD = 4096                                                 # canary elements on both sides of every written buffer

        def guarded(numel, dtype, fill):
            big = torch.full((numel + 2 * GUARD,), fill, dtype=dtype, device=cuda)
            return big, big[GUARD:GUARD + numel]

        def intact(big, numel, fill):
            head, tail = big[:GUARD], big[GUARD + numel:]
            same = (lambda t: torch.isnan(t).all()) if fill != fill else (lambda t: (t == fill).all())
            return bool(same(head)) and bool(same(tail))

        def new(variant, out16):
            ybig, yflat = guarded(n * cout, torch.float32, float("nan")); y = yflat.view(n, cout)
            gbig, g = guarded(gbytes // 4, torch.float32, 12345.0)
            _native.check(L.aprb_linear_f16_stats_ragged(P(x), P(w), n, cin, cout, P(y), P(g), P(seg), S, sp()), "ragged")
            y2 = g2 = None
            if variant == "dual":
                y2 = torch.full((n, cout), float("nan"), device=cuda); g2 = torch.empty(gbytes // 4, device=cuda)
                _native.check(L.aprb_linear_f16_stats_ragged(P(xs), P(ws), n, csc, cout, P(y2), P(g2), P(seg), S, sp()), "ragged")
            nt = 2 if variant == "dual" else 1
            stbig, st = guarded(S * nt * 2 * cout, torch.float32, 777.0)
            _native.check(L.aprb_instnorm_seg_stats(P(y), P(y2), n, cout, P(seg), S, 1e-5, P(g), P(g2), P(st), sp()), "stats")
            obig, oflat = guarded(n * cout, torch.float16 if out16 else torch.float32, 321.0)
            out = oflat.view(n, cout)
            _native.check(L.aprb_linear_f16_norm_apply(P(x), P(w), n, cin, cout, P(xs) if variant == "dual" else None,
                                                       P(ws) if variant == "dual" else None, csc,
                                                       P(res) if variant == "res" else None, P(seg), S, P(st), 0.1, P(out),
                                                       out16, sp()), "apply")
            frac_nan = torch.isnan(y).float().mean().item()
            torch.cuda.synchronize()
            # no kernel of the sequence wrote outside its buffers (compute-sanitizer is closed on this pool)
            assert intact(ybig, n * cout, float("nan")) and intact(gbig, gbytes // 4, 12345.0), (n, variant, "stats pass")
            assert intact(stbig, S * nt * 2 * cout, 777.0) and intact(obig, n * cout, 321.0), (n, variant, "apply pass")
            return out, frac_nan

        for variant in ("none", "res", "dual"):
            for out16 in (1, 0):
                a = old(variant, out16)
                b, frac_nan = new(variant, out16)
                torch.cuda.synchronize()
                assert torch.isfinite(b.float()).all(), (n, variant, out16)
                af, bf = a.float(), b.float()
                assert rel(bf, af) < 2e-5, (n, cin, cout, variant, out16, rel(bf, af))
                assert ((af - bf).abs() <= 2.0 ** -10 * torch.maximum(af.abs(), bf.abs()) + 1e-6).all(), (n, variant, out16)
                if n >= 4097:
                    assert frac_nan > 0.9                            # the product really was not materialised
        # and against torch for the plain-residual variant, segment by segment (fp16 operands, fp32 math)
        b, _ = new("res", 0)
        yf = x.float() @ w.float().t()
        for s0, s1 in zip(bounds[:-1], bounds[1:]):
            if s1 > s0:
                ys = yf[s0:s1]
                ref = (ys - ys.mean(0)) / torch.sqrt(ys.var(0, unbiased=False) + 1e-5) + res[s0:s1].float()
                ref = torch.nn.functional.leaky_relu(ref, 0.1)
                assert rel(b[s0:s1], ref) < 6e-4                     # 10-bit mantissa rounding of the stored activation


def test_native_pipeline_recompute_path_equals_stored_path(cuda, oracle):
    """aprb_kfe_forward with the recompute path (default) vs with the stored fp32 products (gemm_apply = 0), single pair
    and super-batch of three pairs (segment bounds inside tiles): equal up to the re-rounding drift of 11 blocks."""
    from apr_b200 import _native, synth
    from apr_b200.pipeline import KFEPipeline
    cfg = kitti_config()
    limits = [30, 31, 32, 33]
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    pairs = []
    for sd, (na, nb) in enumerate([(3000, 2600), (1800, 3300), (2500, 2500)]):
        a, b = synth.small_cloud(81 + 2 * sd, na), synth.small_cloud(82 + 2 * sd, nb)
        pairs.append(oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3))
    setopt = lambda v: _native.check(_native.lib().aprb_set_option(b"gemm_apply", v), "aprb_set_option")
    for cps, sel in ((0, pairs[:1]), (2, pairs)):
        P0 = np.concatenate([p for p, _ in sel]); L0 = np.concatenate([l for _, l in sel])
        pipe = KFEPipeline(enc, cfg, limits, clouds_per_segment=cps)
        got = pipe.forward(_t(P0, cuda), _t(L0, cuda)).clone()
        n1 = _native.launch_count()
        pipe.forward(_t(P0, cuda), _t(L0, cuda))
        n1 = _native.launch_count() - n1
        try:
            setopt(0)
            want = pipe.forward(_t(P0, cuda), _t(L0, cuda)).clone()
            n0 = _native.launch_count()
            pipe.forward(_t(P0, cuda), _t(L0, cuda))
            n0 = _native.launch_count() - n0
        finally:
            setopt(1)
        torch.cuda.synchronize()
        e = rel(got, want)
        print(f"launches per forward: recompute {n1}, stored {n0}; recompute vs stored {e:.2e}")
        assert e < 5e-3                                              # the TF32 re-rounding drift class (see above)


def test_native_pipeline_arena_estimate_and_fallback(cuda):
    """The arena is sized for level sizes N * 0.6^l; a pyramid that shrinks less (here: estimate forced to 0.05) makes
    the native driver return APRB_ERR_WORKSPACE before writing past the arena, and the call is retried with the
    unconditional bound: same result either way."""
    from apr_b200 import synth
    from apr_b200.pipeline import KFEPipeline
    cfg = kitti_config()
    a, b = synth.small_cloud(71, 2500), synth.small_cloud(72, 2300)
    raw = torch.from_numpy(np.concatenate([a, b])).to(cuda)
    lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=cuda)
    p0, l0 = ops.grid_subsample(raw, lens, 0.3)
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    ref_pipe = KFEPipeline(enc, cfg, [30, 31, 32, 33])
    want = ref_pipe.forward(p0, l0).clone()
    est_bytes = ref_pipe.arena.numel()
    tight = KFEPipeline(enc, cfg, [30, 31, 32, 33])
    tight.LEVEL_RATIO = 0.05
    got = tight.forward(p0, l0).clone()
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert tight.arena.numel() > est_bytes                           # it had to grow to the unconditional bound


def test_native_pipeline_super_batch_equals_separate_pairs(cuda, oracle):
    """P collated pairs stacked in ONE aprb_kfe_forward call (clouds_per_segment = 2: per-pair InstanceNorm statistics)
    give, pair by pair, the result of P separate single-pair calls: pyramid bit-exact (indices shifted by the pair's row
    offset); encoder features equal up to reassociation (norm partial sums and split-K choices depend on the row count,
    and a 1-ulp change can flip the TF32 rounding of a stored activation: 2^-11 relative, amplified over 11 blocks), so
    the bound is the TF32 drift bound, and every pair is also checked against the fp32 CPU oracle."""
    from apr_b200 import synth
    from apr_b200.pipeline import KFEPipeline
    cfg = kitti_config()
    limits = [30, 31, 32, 33]
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    pairs = []
    for sd, (na, nb) in enumerate([(3000, 2600), (1800, 3300), (2500, 2500)]):
        a, b = synth.small_cloud(61 + 2 * sd, na), synth.small_cloud(62 + 2 * sd, nb)
        pairs.append(oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3))
    single = KFEPipeline(enc, cfg, limits)
    outs, pyrs = [], []
    for p0, l0 in pairs:
        outs.append(single.forward(_t(p0, cuda), _t(l0, cuda)).clone())
        pyrs.append({k: [t.clone() for t in v] for k, v in single.pyramid().items()})
    torch.cuda.synchronize()
    batch = KFEPipeline(enc, cfg, limits, clouds_per_segment=2)
    P0 = np.concatenate([p for p, _ in pairs]); L0 = np.concatenate([l for _, l in pairs])
    got = batch.forward(_t(P0, cuda), _t(L0, cuda)).clone()
    pyr = batch.pyramid()
    torch.cuda.synchronize()
    assert got.shape[0] == sum(o.shape[0] for o in outs)
    for lvl in range(4):
        lens = pyr["stack_lengths"][lvl].cpu().numpy()
        assert np.array_equal(lens, np.concatenate([p["stack_lengths"][lvl].cpu().numpy() for p in pyrs]))
        assert torch.equal(pyr["points"][lvl], torch.cat([p["points"][lvl] for p in pyrs]))
        n_tot = pyr["points"][lvl].shape[0]
        for key, sup_lvl in (("neighbors", lvl), ("pools", lvl), ("upsamples", lvl + 1)):
            if pyr[key][lvl].shape[0] == 0:
                continue
            ns_tot = pyr["points"][sup_lvl].shape[0]
            rows = []
            soff = 0
            for p in pyrs:
                t = p[key][lvl].clone()
                ns = p["points"][sup_lvl].shape[0]
                t = torch.where(t == ns, torch.full_like(t, ns_tot), t + soff)        # shadow index = stacked total
                rows.append(t); soff += ns
            assert torch.equal(pyr[key][lvl], torch.cat(rows)), (key, lvl)
    a = 0
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    for i, o in enumerate(outs):
        e = rel(got[a:a + o.shape[0]], o)
        assert e < 5e-3, f"super-batched pair differs from its single-pair run: {e:.2e}"
        if i < 2:
            p0, l0 = pairs[i]
            ref = collate_ref(p0, l0, cfg, limits, oracle.subsample_batch, oracle.batch_query)
            cpu = dict(points=[torch.from_numpy(p) for p in ref["points"]], neighbors=[torch.from_numpy(n).long() for n in ref["neighbors"]],
                       pools=[torch.from_numpy(n).long() for n in ref["pools"]], features=torch.ones(len(p0), 1))
            y = blocks_ref.encoder_ref(cpu, sd, cfg)
            eb, es = rel(got[a:a + o.shape[0]], y), rel(o, y)
            print(f"pair {i}: drift vs fp32 oracle: super-batched {eb:.2e}, single {es:.2e}; batched vs single {e:.2e}")
            assert eb < 1e-2 and es < 1e-2
        a += o.shape[0]
    host = batch.forward_host(torch.from_numpy(P0).pin_memory(), torch.from_numpy(L0).pin_memory())
    assert torch.equal(host, got.cpu())


@pytest.mark.parametrize("mode,tol", [(1, 2e-4), (0, 5e-3)])
def test_kpfcnn_full_forward_golden(cuda, oracle, gold_kpfcnn, mode, tol):
    """BASELINE config 3's network: full KPFCNN forward (encoder + bottleneck GNN + decoder) vs the outputs of the REAL
    reference; the reference's checkpoint keys load with strict=True."""
    g = gold_kpfcnn
    cfg = kitti_config(first_feats_dim=16, gnn_feats_dim=32, final_feats_dim=8)
    pyr = collate_ref(g["p0"], g["l0"], cfg, list(g["limits"]), oracle.subsample_batch, oracle.batch_query)
    gpu = dict(points=[_t(p, cuda) for p in pyr["points"]], neighbors=[_t(n, cuda).long() for n in pyr["neighbors"]],
               pools=[_t(n, cuda).long() for n in pyr["pools"]], upsamples=[_t(n, cuda).long() for n in pyr["upsamples"]],
               stack_lengths=[_t(l, cuda) for l in pyr["stack_lengths"]], features=torch.ones(len(g["p0"]), 1, device=cuda))
    net = KPFCNN(cfg)
    net.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}, strict=True)
    net = net.to(cuda).eval()
    blocks.KPCONV_MODE = mode
    try:
        ff, so, ss = net(gpu)
    finally:
        blocks.KPCONV_MODE = 0
    for got, key in ((ff, "feats_f"), (so, "scores_overlap"), (ss, "scores_saliency")):
        e = rel(got, torch.from_numpy(g[key]))
        print(f"KPFCNN {key} (mode {mode}): rel err {e:.2e}")
        assert got.shape == g[key].shape and e < tol, key


def test_kpfcnn_kitti_width_device_pyramid_vs_oracle(cuda, oracle):
    """Full KPFCNN at the KITTI widths on a small LoKITTI-like pair, pyramid built on the device (collate on the GPU),
    vs the fp32 CPU restatement of the whole reference forward."""
    from apr_b200 import dataloader, synth
    cfg = kitti_config()
    a = synth.small_cloud(71, 3200)
    b = synth.small_cloud(71, 3000) + np.array([6.0, 0.5, 0.0], np.float32)          # same scene, shifted sensor
    p0, l0 = oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3)
    limits = [32, 32, 32, 32]
    torch.manual_seed(0); np.random.seed(0)
    net = KPFCNN(cfg).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ref = collate_ref(p0, l0, cfg, limits, oracle.subsample_batch, oracle.batch_query)
    cpu = dict(points=[torch.from_numpy(p) for p in ref["points"]], neighbors=[torch.from_numpy(n).long() for n in ref["neighbors"]],
               pools=[torch.from_numpy(n).long() for n in ref["pools"]], upsamples=[torch.from_numpy(n).long() for n in ref["upsamples"]],
               stack_lengths=[torch.from_numpy(l) for l in ref["stack_lengths"]], features=torch.ones(len(p0), 1))
    want = blocks_ref.kpfcnn_ref(cpu, sd, cfg)
    net = net.to(cuda)
    pyr = dataloader.build_pyramid_device(_t(p0, cuda), _t(l0, cuda), cfg, limits)
    blocks.LINEAR_MODE = 'tf32'
    try:
        got = net(pyr)
    finally:
        blocks.LINEAR_MODE = 'fp32'
    for gt, wt, name in zip(got, want, ("feats_f", "scores_overlap", "scores_saliency")):
        e = rel(gt, wt)
        print(f"KPFCNN kitti width {name}: rel err {e:.2e}")
        assert gt.shape == wt.shape and e < 2e-2, name
