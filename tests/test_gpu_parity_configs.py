"""Parity of the path bench.py times — aprb_kfe_forward in its default mode (fp16 activation storage, KPConv mode 4,
recompute path for the block-closing Linears, super-batched pairs with per-pair InstanceNorm) — on the shapes BASELINE
names: a full KITTI-shaped pair, a nuScenes-shaped pair and a LoKITTI-shaped distant pair, with calibrated limits.

Every block is checked on its own with IDENTICAL inputs on both sides: the device copies (taps) of block b-1's output and
of KPConv b's input feed the fp32 CPU oracle (oracle/blocks_ref.py <- models/blocks.py:229-374, :653-681), whose result
is compared with the tap of block b's output / KPConv b's raw output. Bars: KPConv <= 1e-3 (north_star), block <= TOL_BLOCK.
The kernels that exist only on this path get direct tests: KPConv mode 4, fp16 max_pool, fp16 segmented InstanceNorm,
and the `dual` / `none` variants of the recompute-apply GEMM against torch."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from apr_b200 import _native, blocks, dataloader, ops, synth
from apr_b200.architectures import KPFCNNEncoder
from apr_b200.config import kitti_config, nuscenes_config
from apr_b200.pipeline import KFEPipeline
from oracle import blocks_ref
from oracle.ref import calibrate_ref, collate_ref

pytestmark = pytest.mark.gpu

TOL_KPCONV = 1e-3     # north_star: KPConv features within 1e-3 relative of the fp32 reference
TOL_BLOCK = 1.5e-3    # a whole ResnetBottleneckBlock: three fp16-operand contractions + two 10-bit re-roundings of stored
                      # activations (measured 5-9e-4 on B200, printed below)
# End of encoder, free-running (no re-synchronisation of inputs): the per-block errors above are amplified by the 10 blocks
# that follow them — a property of the random-init network (InstanceNorm rescales, lidar-shaped clouds have heavy-tailed
# channels), not of the kernels. The yardstick is therefore the fp32 oracle ITSELF with gaussian noise of the measured
# per-block size injected after every block: our drift must stay within DRIFT_FACTOR of that, and under DRIFT_CAP.
#   (a) vs the fp32 oracle: reported, capped at DRIFT_CAP;
#   (b) vs the fp32 oracle given gaussian noise of the measured per-block size after every block (uncorrelated noise
#       averages out in the next KPConv's neighbourhood sums, a rounded WEIGHT does not): reported;
#   (c) the fp32 oracle against ITSELF run in the product's number format (quant = round to a 10-bit mantissa applied to
#       every contraction operand and stored activation; statistics and accumulation fp32): the drift the format alone
#       causes (measured on B200, r02: 2.1-2.9e-2 on these pairs; ours vs fp32: 1.9-2.6e-2; ours vs the format oracle:
#       1.0-1.3e-2 — two runs in the same format differ in which values fall on the other side of a rounding boundary,
#       and the network amplifies that like any other perturbation). Bar: our drift <= DRIFT_VS_FORMAT x (c).
DRIFT_CAP = 6e-2
DRIFT_VS_FORMAT = 1.5


def _q10(t):
    """Round to fp16 and back: the operand / storage format of the product path (10-bit mantissa, as TF32)."""
    return t.half().float()


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _t(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def _voxel_pair(oracle, seed, kind, distant):
    a, b = synth.pair_raw(seed, kind, distant)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    return oracle.subsample_batch(raw, lens, sampleDl=0.3)               # first-level voxelisation (stand-in)


def _pair_batch(pyr, j, nlev):
    """CPU batch dict (reference layout: local indices, pad = the pair's own support count) of pair j of a super-batched
    device pyramid, plus the pair's row range per level."""
    rows, out = [], dict(points=[], neighbors=[], pools=[])
    for l in range(nlev):
        lens = pyr["stack_lengths"][l].cpu().numpy().astype(np.int64)
        off = int(lens[:2 * j].sum()); n = int(lens[2 * j] + lens[2 * j + 1])
        rows.append((off, n))
    for l in range(nlev):
        off, n = rows[l]
        out["points"].append(pyr["points"][l][off:off + n].cpu())
        ns_tot = pyr["points"][l].shape[0]
        t = pyr["neighbors"][l][off:off + n].cpu().long()
        out["neighbors"].append(torch.where(t >= ns_tot, torch.full_like(t, n), t - off))
        if l + 1 < nlev and pyr["pools"][l].shape[0]:
            qo, qn = rows[l + 1]
            t = pyr["pools"][l][qo:qo + qn].cpu().long()
            out["pools"].append(torch.where(t >= ns_tot, torch.full_like(t, n), t - off))
        else:
            out["pools"].append(torch.zeros((0, 1), dtype=torch.long))
    return out, rows


def _cut_to_reference_width(m, pad):
    """The reference's matrix has min(max_count, limit) columns: drop the trailing columns that are pad in every row."""
    valid = (m != pad).any(0)
    w = int(valid.nonzero().max()) + 1 if bool(valid.any()) else 1
    return m[:, :w]


def _noisy_oracle_drift(batch, sd, cfg, errs, y_clean):
    """Drift of the fp32 oracle encoder when gaussian noise of relative size errs[b] is added to block b's output."""
    gen = torch.Generator().manual_seed(123)
    x = batch["features"].clone()
    r = cfg.first_subsampling_dl * cfg.conv_radius
    layer = 0
    for bi, name in enumerate(b for b in cfg.architecture if "upsample" not in b and "unary" not in b):
        prefix, extent = f"encoder_blocks.{bi}.", r * cfg.KP_extent / cfg.conv_radius
        fn = blocks_ref.simple_ref if "simple" in name else blocks_ref.resnetb_ref
        x = fn(x, batch, sd, prefix, name, layer, extent)
        nz = torch.randn(x.shape, generator=gen)
        x = x + nz * (errs[bi] * x.norm() / nz.norm())
        if "strided" in name:
            layer += 1; r *= 2
    return rel(x, y_clean)


@pytest.mark.parametrize("kind,distant", [("kitti", False), ("nuscenes", False), ("kitti", True)],
                         ids=["kitti_pair", "nuscenes_pair", "lokitti_distant_pair"])
def test_benchmarked_path_per_block_vs_oracle(cuda, oracle, kind, distant):
    cfg = kitti_config() if kind == "kitti" else nuscenes_config()
    nlev = cfg.num_layers
    pairs = [_voxel_pair(oracle, seed, kind, distant) for seed in (3, 4, 5)]
    dev = [(_t(p, cuda), _t(l, cuda)) for p, l in pairs]
    limits = [int(v) for v in dataloader.calibrate_neighbors_device(dev, cfg)]
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    pipe = KFEPipeline(enc, cfg, limits, clouds_per_segment=2)            # the bench's configuration: super-batched pairs
    pipe.set_tap(3 << 30)
    P0 = torch.cat([p for p, _ in dev]); L0 = torch.cat([l for _, l in dev])
    out = pipe.forward(P0, L0).clone()
    torch.cuda.synchronize()
    pyr, taps = pipe.pyramid(), pipe.taps()
    j = 1                                                                 # the middle pair: segment bounds inside tiles
    cpu, rows = _pair_batch(pyr, j, nlev)
    # the oracle sees what the reference would: matrices cut to min(max_count, limit) columns
    cpu_ref = dict(points=cpu["points"],
                   neighbors=[_cut_to_reference_width(m, len(cpu["points"][l])) for l, m in enumerate(cpu["neighbors"])],
                   pools=[_cut_to_reference_width(m, len(cpu["points"][l])) if m.shape[0] else m for l, m in enumerate(cpu["pools"])])
    print(f"\n{kind}{' distant' if distant else ''}: stacked pair rows per level {[n for _, n in rows]}, limits {limits}, "
          f"reference widths conv {[m.shape[1] for m in cpu_ref['neighbors']]} pool {[m.shape[1] for m in cpu_ref['pools'][:-1]]}")
    r = cfg.first_subsampling_dl * cfg.conv_radius
    layer, worst_kp, worst_blk, errs = 0, 0.0, 0.0, []
    feats = torch.ones(rows[0][1], 1)
    arch = [b for b in cfg.architecture if "upsample" not in b and "unary" not in b]
    for bi, name in enumerate(arch):
        prefix = f"encoder_blocks.{bi}."
        extent = r * cfg.KP_extent / cfg.conv_radius
        strided = "strided" in name
        off_o, n_o = rows[layer + 1] if strided else rows[layer]
        off_i, n_i = rows[layer]
        got = taps[(bi, "out")][off_o:off_o + n_o].float().cpu()
        if "simple" in name:
            want = blocks_ref.simple_ref(feats, cpu_ref, sd, prefix, name, layer, extent)
            e_kp = float("nan")
        else:
            want = blocks_ref.resnetb_ref(feats, cpu_ref, sd, prefix, name, layer, extent)
            x1 = taps[(bi, "kp_in")][off_i:off_i + n_i].float().cpu()   # fp16 storage of a 10-bit value: exact widening
            q, s, inds = blocks_ref._select(name, layer, cpu_ref)
            want_kp = blocks_ref.kpconv_ref(q, s, inds, x1, sd[prefix + "KPConv.kernel_points"], sd[prefix + "KPConv.weights"], extent)
            e_kp = rel(taps[(bi, "kp_out")][off_o:off_o + n_o], want_kp)
            worst_kp = max(worst_kp, e_kp)
        e = rel(got, want)
        worst_blk = max(worst_blk, e)
        errs.append(e)
        print(f"  block {bi:2d} {name:16s} L{layer} rows {n_o:6d}  KPConv rel err {e_kp:.2e}   block rel err {e:.2e}")
        feats = got                                                        # next block: identical input on both sides
        if strided:
            layer += 1; r *= 2
    # end of encoder: free-running oracle (no re-synchronisation of inputs)
    cpu_ref["features"] = torch.ones(rows[0][1], 1)
    y = blocks_ref.encoder_ref(cpu_ref, sd, cfg)
    off, n = rows[nlev - 1]
    drift = rel(out[off:off + n], y)
    model = _noisy_oracle_drift(cpu_ref, sd, cfg, errs, y)
    yq = blocks_ref.encoder_ref(cpu_ref, sd, cfg, quant=_q10)
    fmt, left = rel(yq, y), rel(out[off:off + n], yq)
    print(f"  worst KPConv {worst_kp:.2e}, worst block {worst_blk:.2e}; end-of-encoder drift vs fp32 oracle {drift:.2e} "
          f"[fp32 oracle + per-block gaussian noise: {model:.2e}; oracle in the 10-bit operand/storage format vs fp32 oracle: "
          f"{fmt:.2e}]; vs the oracle in the product's format: {left:.2e}")
    assert worst_kp < TOL_KPCONV
    assert worst_blk < TOL_BLOCK
    assert drift < DRIFT_CAP and drift < DRIFT_VS_FORMAT * fmt
    assert left < DRIFT_VS_FORMAT * fmt
    pipe.set_tap(0)


@pytest.mark.parametrize("cin,cout,h", [(64, 64, 57), (128, 128, 33), (256, 256, 40), (512, 512, 20), (64, 128, 56)])
def test_kpconv_mode4_fp16_features_vs_oracle(cuda, cin, cout, h):
    """KPConv with fp16 feature rows in (mode 4, what aprb_kfe_forward runs): the oracle is fed the same fp16 values."""
    gen = torch.Generator().manual_seed(cin + h)
    ns, nq = 2000, 1500
    s = torch.rand(ns, 3, generator=gen) * 2
    q = torch.rand(nq, 3, generator=gen) * 2
    inds = torch.randint(0, ns + 1, (nq, h), generator=gen)
    inds[:3] = ns
    inds[5, h // 2:] = ns
    x16 = F.leaky_relu(torch.randn(ns, cin, generator=gen), 0.1).half()
    x16[::7] = -x16[::7].abs()                                            # rows with negative sums: not counted in neighbor_num
    kp = torch.randn(15, 3, generator=gen) * 0.4
    w = torch.randn(15, cin, cout, generator=gen) / np.sqrt(15 * cin)
    want = blocks_ref.kpconv_ref(q, s, inds, x16.float(), kp, w, 0.7)
    wd = w.to(cuda)
    got = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda).int(), x16.to(cuda), kp.to(cuda), wd, 0.7,
                     wprep=ops.kpconv_prepare_weights_f16(wd), mode=4)
    e = rel(got, want)
    print(f"kpconv mode 4 Cin={cin} Cout={cout} H={h}: rel err {e:.2e}")
    assert e < TOL_KPCONV
    assert torch.all(got[:3] == 0)
    # mode 5: the weighting stage itself on tcgen05 (kpconv_tc.cu, optional: slower than the list kernel, see
    # profiles/r02_kpconv_tc.txt), weighted tile channel-major
    setopt = lambda v: _native.check(_native.lib().aprb_set_option(b"kpconv_tc", v), "aprb_set_option")
    try:
        setopt(1)
        if ops.kpconv_tc_supported(h, 15, cin, cout, ns):
            got5 = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda).int(), x16.to(cuda), kp.to(cuda), wd, 0.7,
                              wprep=ops.kpconv_prepare_weights_f16_ck(wd), mode=5)
            e5 = rel(got5, want)
            print(f"kpconv mode 5 Cin={cin} Cout={cout} H={h}: rel err {e5:.2e}")
            assert e5 < TOL_KPCONV
            assert torch.all(got5[:3] == 0)
            # the weighted tile itself, against its fp32 definition (blocks.py:269-354), query ranges that end inside a tile
            for nq_part in (nq, 1, 37):
                wf, inv = ops.kpconv_weighted_f16(q[:nq_part].to(cuda), s.to(cuda), inds[:nq_part].to(cuda).int(), x16.to(cuda), kp.to(cuda), 0.7,
                                                  layout_ck=True)
                sp = torch.cat((s, torch.zeros(1, 3) + 1e6)); nb = sp[inds[:nq_part]] - q[:nq_part].unsqueeze(1)
                w = torch.clamp(1 - torch.sqrt(((nb.unsqueeze(2) - kp) ** 2).sum(3)) / 0.7, min=0).transpose(1, 2)
                xz = torch.cat((x16.float(), torch.zeros(1, cin)))
                want_wf = torch.matmul(w, xz[inds[:nq_part]])                       # [nq, 15, cin]
                got_wf = wf.float().cpu().view(nq_part, cin, 16)
                assert rel(got_wf[:, :, :15].permute(0, 2, 1), want_wf) < 6e-4
                assert bool((got_wf[:, :, 15] == 0).all())
                nn_ = (xz[inds[:nq_part]].sum(-1) > 0).sum(-1).clamp_min(1).float()
                assert torch.equal(inv.cpu(), 1.0 / nn_)
        else:
            assert cin == 512
    finally:
        setopt(0)


def test_fp16_operand_path_saturates_instead_of_overflowing(cuda):
    """Mode 3 takes arbitrary fp32 features: a weighted sum outside the fp16 range clamps to +-65504 in the weighted tile
    (pack_half2_sat) — the output stays finite, and rows with in-range features are unaffected."""
    gen = torch.Generator().manual_seed(3)
    ns, nq, h, cin, cout = 600, 400, 20, 64, 64
    s = torch.rand(ns, 3, generator=gen); q = torch.rand(nq, 3, generator=gen)
    inds = torch.randint(0, ns, (nq, h), generator=gen)
    inds[: nq // 2] = inds[: nq // 2] % (ns // 2)                           # first half of the queries: supports < ns/2 only
    x = torch.randn(ns, cin, generator=gen)
    kp = torch.randn(15, 3, generator=gen) * 0.3
    w = torch.randn(15, cin, cout, generator=gen) / np.sqrt(15 * cin)
    wd = w.to(cuda); prep = ops.kpconv_prepare_weights_f16(wd)
    base = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda).int(), x.to(cuda), kp.to(cuda), wd, 0.5, wprep=prep, mode=3)
    xb = x.clone(); xb[ns // 2:] *= 1e6                                     # second half of the supports: far outside fp16
    got = ops.kpconv(q.to(cuda), s.to(cuda), inds.to(cuda).int(), xb.to(cuda), kp.to(cuda), wd, 0.5, wprep=prep, mode=3)
    assert bool(torch.isfinite(got).all())
    assert torch.equal(got[: nq // 2], base[: nq // 2])
    h16 = ops.f32_to_f16(torch.tensor([1e9, -1e9, 65504.0, 1.0, float("inf")], device=cuda)) if hasattr(ops, "f32_to_f16") else None
    if h16 is not None:
        assert h16.cpu().tolist() == [65504.0, -65504.0, 65504.0, 1.0, 65504.0]


def test_max_pool_f16_vs_oracle(cuda):
    """fp16 max_pool (product path): the max of fp16 values is exact, so the result is bit-identical to the oracle's on
    the same values — pads anywhere, all-pad rows, H > 32, and per-segment reference widths."""
    gen = torch.Generator().manual_seed(8)
    for c, h in ((128, 9), (256, 56), (512, 33), (1024, 70)):
        x = torch.randn(300, c, generator=gen).half()
        inds = torch.randint(0, 301, (77, h), generator=gen)
        inds[3] = 300
        inds[4, : h // 2] = 300
        for xx in (x, -x.abs()):
            want = blocks_ref.max_pool_ref(xx.float(), inds).half()
            got = ops.max_pool(xx.to(cuda), inds.to(cuda).int())
            assert got.dtype == torch.float16 and torch.equal(got.cpu(), want)
    # per-segment widths: trailing all-pad columns of a segment do not inject the zero row (negative features stay negative)
    ns, c, h = 200, 128, 12
    x = (-torch.rand(ns, c, generator=gen) - 0.5).half()
    inds = torch.full((60, h), ns, dtype=torch.int64)
    for n in range(60):
        k = 5 if n < 30 else 9                                            # segment 0 rows hold <= 5, segment 1 rows <= 9 valid entries
        cnt = int(torch.randint(1, k + 1, (1,), generator=gen))
        if n in (0, 30):
            cnt = k                                                       # one row per segment attains the segment's width
        inds[n, :cnt] = torch.randint(0, ns, (cnt,), generator=gen)
    seg = torch.tensor([0, 30, 60], dtype=torch.int32, device=cuda)
    widths = ops.pool_seg_widths(inds.to(cuda).int(), ns, seg)
    assert widths.cpu().tolist() == [5, 9]
    want = torch.cat([blocks_ref.max_pool_ref(x.float(), inds[:30, :5]), blocks_ref.max_pool_ref(x.float(), inds[30:, :9])])
    for xd, cast in ((x.to(cuda), lambda t: t.half()), (x.float().to(cuda), lambda t: t)):
        got = ops.max_pool(xd, inds.to(cuda).int(), width_dev=widths, seg_off=seg)
        assert torch.equal(got.cpu(), cast(want))
        assert bool((got[0] < 0).all()) and bool((got[30] < 0).all())    # full rows: no phantom zero
        full = ops.max_pool(xd, inds.to(cuda).int())                      # without the widths every row sees a pad
        assert bool((full == 0).all())


def test_instnorm_seg_f16_vs_torch(cuda):
    """aprb_instnorm_lrelu_seg_f16 (fp32 GEMM output in, fp16 activation out, optional fp16 residual) vs torch per segment;
    the stored value is the 10-bit rounding of the fp32 result."""
    L, P, sp = _native.lib(), _native.ptr, _native.stream_ptr
    gen = torch.Generator().manual_seed(21)
    for bounds, c in (([0, 700, 1350, 2150, 2180, 2181, 3081], 64), ([0, 5000, 9000], 128), ([0, 300], 2048)):
        n, S = bounds[-1], len(bounds) - 1
        x = (torch.randn(n, c, generator=gen) * 2 + torch.linspace(-20, 20, n).unsqueeze(1)).to(cuda)
        res = torch.randn(n, c, generator=gen).half().to(cuda)
        seg = torch.tensor(bounds, dtype=torch.int32, device=cuda)
        ws = torch.empty(int(L.aprb_instnorm_seg_ws_bytes(n, c, S)), dtype=torch.uint8, device=cuda)
        for use_res in (False, True):
            for out16 in (1, 0):
                y = torch.empty(n, c, dtype=torch.float16 if out16 else torch.float32, device=cuda)
                _native.check(L.aprb_instnorm_lrelu_seg_f16(P(x), n, c, P(seg), S, 1e-5, 0.1, P(res) if use_res else None, 1, 0, 1,
                                                            P(y), out16, None, None, P(ws), ws.numel(), sp()), "norm16")
                for s0, s1 in zip(bounds[:-1], bounds[1:]):
                    if s1 - s0 > 1:
                        xs = x[s0:s1].double()
                        ref = (xs - xs.mean(0)) / torch.sqrt(xs.var(0, unbiased=False) + 1e-5)
                        if use_res:
                            ref = ref + res[s0:s1].double()
                        ref = F.leaky_relu(ref, 0.1).float()
                        got = y[s0:s1].float()
                        assert ((got - ref).abs() <= 2.0 ** -10 * ref.abs() + 2e-4).all(), (bounds, c, use_res, out16)
                        assert rel(got, ref) < 6e-4


def test_linear_norm_apply_all_variants_vs_torch(cuda):
    """The recompute-apply GEMM (gemm_nrm_f16_kernel) in its three forms — no shortcut, fp16 residual rows, and `dual`
    (the shortcut product accumulated and standardised in the same CTA) — against plain torch on the same fp16 operands."""
    L, P, sp = _native.lib(), _native.ptr, _native.stream_ptr
    gen = torch.Generator().manual_seed(15)
    for n, cin, cout, csc, bounds in ((5000, 64, 256, 128, [0, 1234, 1250, 3001, 5000]), (20000, 256, 1024, 512, [0, 9000, 20000]),
                                      (700, 128, 512, 256, [0, 700])):
        S = len(bounds) - 1
        x = F.leaky_relu(torch.randn(n, cin, generator=gen), 0.1).half().to(cuda)
        w = (torch.randn(cout, cin, generator=gen) / np.sqrt(cin)).half().to(cuda)
        xs = torch.randn(n, csc, generator=gen).half().to(cuda)
        wsc = (torch.randn(cout, csc, generator=gen) / np.sqrt(csc)).half().to(cuda)
        res = torch.randn(n, cout, generator=gen).half().to(cuda)
        seg = torch.tensor(bounds, dtype=torch.int32, device=cuda)
        gbytes = int(L.aprb_group_stats_bytes(n, cout))
        yf, y2f = x.float() @ w.float().t(), xs.float() @ wsc.float().t()
        for variant in ("none", "res", "dual"):
            y = torch.full((n, cout), float("nan"), device=cuda); g = torch.empty(gbytes // 4, device=cuda)
            _native.check(L.aprb_linear_f16_stats_ragged(P(x), P(w), n, cin, cout, P(y), P(g), P(seg), S, sp()), "ragged")
            y2 = g2 = None
            if variant == "dual":
                y2 = torch.full((n, cout), float("nan"), device=cuda); g2 = torch.empty(gbytes // 4, device=cuda)
                _native.check(L.aprb_linear_f16_stats_ragged(P(xs), P(wsc), n, csc, cout, P(y2), P(g2), P(seg), S, sp()), "ragged")
            st = torch.empty(S * (2 if variant == "dual" else 1) * 2 * cout, device=cuda)
            _native.check(L.aprb_instnorm_seg_stats(P(y), P(y2), n, cout, P(seg), S, 1e-5, P(g), P(g2), P(st), sp()), "stats")
            out = torch.empty(n, cout, device=cuda)
            _native.check(L.aprb_linear_f16_norm_apply(P(x), P(w), n, cin, cout, P(xs) if variant == "dual" else None,
                                                       P(wsc) if variant == "dual" else None, csc, P(res) if variant == "res" else None,
                                                       P(seg), S, P(st), 0.1, P(out), 0, sp()), "apply")
            for s0, s1 in zip(bounds[:-1], bounds[1:]):
                ys = yf[s0:s1]
                ref = (ys - ys.mean(0)) / torch.sqrt(ys.var(0, unbiased=False) + 1e-5)
                if variant == "res":
                    ref = ref + res[s0:s1].float()
                elif variant == "dual":
                    y2s = y2f[s0:s1]
                    ref = ref + (y2s - y2s.mean(0)) / torch.sqrt(y2s.var(0, unbiased=False) + 1e-5)
                ref = F.leaky_relu(ref, 0.1)
                e = rel(out[s0:s1], ref)
                assert e < 6e-4, (n, cin, cout, variant, e)              # 10-bit mantissa rounding of the stored activation


def test_pool_width_guard_sparse_cloud(cuda, oracle):
    """a pool search whose max_count is below the limit leaves all-pad columns in the fixed-width device matrix;
    the max_pool of a strided shortcut must cut at the reference's width min(max_count, limit) per collated pair, or rows
    the reference sees full gain a phantom zero. Sparse clouds + wide limits, single pair and super-batch, native and
    module path, against the oracle on the reference-width matrices (collate_ref)."""
    cfg = kitti_config()
    limits = [45, 45, 45, 45]
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    pairs = []
    for sdn, (na, nb) in enumerate([(700, 600), (1500, 400)]):
        a, b = synth.small_cloud(91 + sdn, na, extent=(25.0, 15.0, 3.0)), synth.small_cloud(95 + sdn, nb, extent=(25.0, 15.0, 3.0))
        pairs.append(oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3))
    refs, narrow = [], 0
    for p0, l0 in pairs:
        ref = collate_ref(p0, l0, cfg, limits, oracle.subsample_batch, oracle.batch_query)
        narrow += sum(int(m.shape[0] > 0 and m.shape[1] < lim) for m, lim in zip(ref["pools"], limits))
        cpu = dict(points=[torch.from_numpy(p) for p in ref["points"]], neighbors=[torch.from_numpy(n).long() for n in ref["neighbors"]],
                   pools=[torch.from_numpy(n).long() for n in ref["pools"]], features=torch.ones(len(p0), 1))
        refs.append(blocks_ref.encoder_ref(cpu, sd, cfg, return_all=True))
    assert narrow >= 2, "test clouds are not sparse enough to exercise the width cut"
    # native, super-batch of both pairs: per-pair widths
    pipe = KFEPipeline(enc, cfg, limits, clouds_per_segment=2)
    pipe.set_tap(1 << 30)
    P0 = np.concatenate([p for p, _ in pairs]); L0 = np.concatenate([l for _, l in pairs])
    for act16 in (1, 0):
        _native.check(_native.lib().aprb_set_option(b"act_f16", act16), "aprb_set_option")
        try:
            pipe.forward(_t(P0, cuda), _t(L0, cuda))
            torch.cuda.synchronize()
            pyr, taps = pipe.pyramid(), pipe.taps()
        finally:
            _native.check(_native.lib().aprb_set_option(b"act_f16", 1), "aprb_set_option")
        for j in range(2):
            _, rows = _pair_batch(pyr, j, cfg.num_layers)
            layer = 0
            for bi, name in enumerate(b for b in cfg.architecture if "upsample" not in b and "unary" not in b):
                if "strided" in name:
                    layer += 1
                off, n = rows[layer]
                e = rel(taps[(bi, "out")][off:off + n], refs[j][bi])
                assert e < 1e-2, (act16, j, bi, name, e)                   # a phantom zero row costs > 5e-2 at the strided blocks
    pipe.set_tap(0)
    # module path, one pair, widths from build_pyramid_device
    p0, l0 = pairs[0]
    pyr = dataloader.build_pyramid_device(_t(p0, cuda), _t(l0, cuda), cfg, limits)
    blocks.LINEAR_MODE = 'tf32'
    try:
        got = enc(pyr)
    finally:
        blocks.LINEAR_MODE = 'fp32'
    assert rel(got, refs[0][-1]) < 1e-2


def test_calibrate_neighbors_device_vs_oracle(cuda, oracle):
    """calibrate_neighbors_device (what bench.py uses) == the reference's calibrate_neighbors (dataloader.py:200-232)
    restated on the oracle, on two small pairs."""
    cfg = kitti_config()
    pairs = []
    for sdn in (0, 1):
        a, b = synth.small_cloud(31 + sdn, 2500), synth.small_cloud(41 + sdn, 2300)
        pairs.append(oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3))
    want = calibrate_ref(pairs, cfg, oracle.subsample_batch, oracle.batch_query, samples_threshold=10 ** 9)
    got = dataloader.calibrate_neighbors_device([(_t(p, cuda), _t(l, cuda)) for p, l in pairs], cfg, samples_threshold=10 ** 9)
    assert np.array_equal(np.asarray(got), np.asarray(want)), (got, want)


def test_pipeline_refreshes_stale_weights(cuda, oracle):
    """the native handle snapshots the encoder's weights; a later load_state_dict / in-place update must not
    leave the tensor path on stale operands."""
    cfg = kitti_config()
    a, b = synth.small_cloud(11, 1500), synth.small_cloud(12, 1300)
    p0, l0 = oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3)
    limits = [30, 30, 30, 30]
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    torch.manual_seed(1); np.random.seed(1)
    other = KPFCNNEncoder(cfg).to(cuda).eval()
    pipe = KFEPipeline(enc, cfg, limits)
    y0 = pipe.forward(_t(p0, cuda), _t(l0, cuda)).clone()
    want = KFEPipeline(other, cfg, limits).forward(_t(p0, cuda), _t(l0, cuda)).clone()
    enc.load_state_dict(other.state_dict())
    y1 = pipe.forward(_t(p0, cuda), _t(l0, cuda)).clone()
    torch.cuda.synchronize()
    assert not torch.equal(y0, y1)
    assert torch.equal(y1, want)


def test_pipeline_nearest_upsamples_equal_column0_of_the_full_matrices(cuda, oracle):
    """build_upsamples="nearest" (the upsample search only needs column 0): [N_l, 1] matrices bit-identical to
    column 0 of the full ones and of the oracle's collate; everything else of the pyramid and the encoder output unchanged."""
    cfg = kitti_config()
    clouds = [synth.small_cloud(21 + i, 1400 + 150 * i) for i in range(4)]          # two collated pairs in one call
    p0, l0 = oracle.subsample_batch(np.concatenate(clouds), np.array([len(c) for c in clouds], np.int32), sampleDl=0.3)
    limits = [30, 30, 30, 30]
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(cuda).eval()
    full = KFEPipeline(enc, cfg, limits, build_upsamples=True, clouds_per_segment=2)
    near = KFEPipeline(enc, cfg, limits, build_upsamples="nearest", clouds_per_segment=2)
    yf = full.forward(_t(p0, cuda), _t(l0, cuda)).clone()
    yn = near.forward(_t(p0, cuda), _t(l0, cuda)).clone()
    torch.cuda.synchronize()
    assert torch.equal(yf, yn)
    pf, pn = full.pyramid(), near.pyramid()
    for l in range(cfg.num_layers - 1):
        assert pn["upsamples"][l].shape == (pf["points"][l].shape[0], 1)
        assert torch.equal(pn["upsamples"][l][:, 0], pf["upsamples"][l][:, 0])
        want = oracle.batch_query(pf["points"][l].cpu().numpy(), pf["points"][l + 1].cpu().numpy(), pf["stack_lengths"][l].cpu().numpy(),
                                  pf["stack_lengths"][l + 1].cpu().numpy(), radius=2 * 1.275 * 2 ** l)
        assert np.array_equal(pn["upsamples"][l][:, 0].cpu().numpy(), want[:, 0])
        for key in ("neighbors", "pools"):
            assert torch.equal(pn[key][l], pf[key][l])


def test_kpfcnn_pipeline_config3_vs_oracle(cuda, oracle):
    """BASELINE config 3's network through KPFCNNPipeline (native encoder + bottleneck / GCN / decoder stream-ordered in the
    same call, two LoKITTI-like pairs super-batched) at the KITTI widths:
      (a) the bottleneck / GNN / decoder half on IDENTICAL inputs — the oracle (blocks_ref.kpfcnn_ref <- architectures.py:
          155-212, models/gcn.py) is fed the product's own encoder activations — <= 5e-3 (the bar VERDICT r1 set);
      (b) end to end against the free-running fp32 oracle: reported, bounded by the operand-format drift (see above)."""
    from apr_b200.architectures import KPFCNN
    from apr_b200.pipeline import KPFCNNPipeline
    cfg = kitti_config()
    limits = [32, 32, 32, 32]
    torch.manual_seed(0); np.random.seed(0)
    net = KPFCNN(cfg).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(cuda)
    pairs = []
    for sdn in (0, 1):
        a = synth.small_cloud(71 + sdn, 3200)
        b = synth.small_cloud(71 + sdn, 3000) + np.array([6.0, 0.5, 0.0], np.float32)      # same scene, shifted sensor
        pairs.append(oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), sampleDl=0.3))
    P0 = np.concatenate([p for p, _ in pairs]); L0 = np.concatenate([l for _, l in pairs])
    blocks.LINEAR_MODE = 'tf32'
    try:
        pipe = KPFCNNPipeline(net, cfg, limits, clouds_per_segment=2)
        ff, so, ss = pipe.forward(_t(P0, cuda), _t(L0, cuda))
        torch.cuda.synchronize()
        pyr = pipe.enc.pyramid()
        enc_outs = [pipe.enc.block_output(i).float().cpu() for i in range(len(net.encoder_blocks))]
    finally:
        blocks.LINEAR_MODE = 'fp32'
    n_enc = len(net.encoder_blocks)
    arch = [b for b in cfg.architecture if "upsample" not in b and "unary" not in b]
    for j, (p0, l0) in enumerate(pairs):
        ref = collate_ref(p0, l0, cfg, limits, oracle.subsample_batch, oracle.batch_query)
        cpu = dict(points=[torch.from_numpy(p) for p in ref["points"]], neighbors=[torch.from_numpy(n).long() for n in ref["neighbors"]],
                   pools=[torch.from_numpy(n).long() for n in ref["pools"]], upsamples=[torch.from_numpy(n).long() for n in ref["upsamples"]],
                   stack_lengths=[torch.from_numpy(l) for l in ref["stack_lengths"]], features=torch.ones(len(p0), 1))
        _, rows = _pair_batch(pyr, j, cfg.num_layers)
        # this pair's rows of every encoder block output
        outs_j, layer = [], 0
        for bi, name in enumerate(arch):
            if "strided" in name:
                layer += 1
            off, n = rows[layer]
            outs_j.append(enc_outs[bi][off:off + n])
        off0, n0 = rows[0]
        got = (ff[off0:off0 + n0], so[off0:off0 + n0], ss[off0:off0 + n0])
        half = blocks_ref.kpfcnn_ref(cpu, sd, cfg, enc_outs=outs_j)           # decoder half on identical inputs
        full = blocks_ref.kpfcnn_ref(cpu, sd, cfg)                            # free-running fp32 oracle
        for gt, wh, wf_, name in zip(got, half, full, ("feats_f", "scores_overlap", "scores_saliency")):
            e_half, e_full = rel(gt, wh), rel(gt, wf_)
            print(f"pair {j} KPFCNN {name}: decoder half on identical inputs {e_half:.2e}; end to end vs fp32 oracle {e_full:.2e}")
            assert gt.shape == wh.shape
            assert e_half < 1e-4, (j, name, e_half)                          # measured 6e-8 .. 2.5e-7 (fp32 library ops)
            assert e_full < 5e-3, (j, name, e_full)                          # measured 2e-4 .. 1.1e-3 on these clouds
