"""GPU parity: grid subsampling (K1) and radius neighbour search (K2+K3) through the C-ABI vs the CPU oracle and the
reference golden vectors. Index lists are compared BIT-EXACT (order and shadow padding included); subsampled points
bit-exact in canonical voxel order (the bar is 1e-6 relative)."""
import numpy as np
import pytest
import torch

from apr_b200 import ops, synth
from apr_b200.config import kitti_config
from oracle.ref import d2_rows, equal_modulo_ties, lexsort_rows_per_cloud

pytestmark = pytest.mark.gpu


def _dev(a, cuda, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    return t.to(dtype) if dtype is not None else t


def test_subsample_golden_bit_exact(cuda, gold_l1):
    raw, lens = gold_l1["raw"], gold_l1["lens"]
    for dl in (0.3, 0.6):
        p, l = ops.grid_subsample(_dev(raw, cuda), _dev(lens, cuda), dl)
        assert np.array_equal(l.cpu().numpy(), gold_l1[f"sub_{dl}_lens"])
        assert np.array_equal(p.cpu().numpy(), gold_l1[f"sub_{dl}_points_canonical"])
        assert np.array_equal(lexsort_rows_per_cloud(p.cpu().numpy(), l.cpu().numpy()),
                              gold_l1[f"sub_{dl}_points_ref_lexsorted"])


def test_subsample_kitti_size_vs_oracle(cuda, oracle):
    a, b = synth.pair_raw(0)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    pts, ln = raw, lens
    for dl in (0.3, 0.6, 1.2, 2.4):                                     # first-level voxelisation + the 3 pyramid levels
        po, lo = oracle.subsample_batch(pts, ln, sampleDl=dl)
        pg, lg = ops.grid_subsample(_dev(pts, cuda), _dev(ln, cuda), dl)
        assert np.array_equal(lg.cpu().numpy(), lo)
        assert np.array_equal(pg.cpu().numpy(), po), f"dl={dl}"
        pts, ln = po, lo


def test_subsample_edge_cases(cuda, oracle):
    rng = np.random.default_rng(0)
    # single point, single cloud
    p, l = ops.grid_subsample(_dev(np.array([[1.0, 2.0, 3.0]], np.float32), cuda), _dev(np.array([1], np.int32), cuda), 0.5)
    assert p.shape == (1, 3) and l.tolist() == [1] and np.allclose(p.cpu().numpy(), [[1, 2, 3]])
    # ragged batch with an empty cloud in the middle, negative coordinates, duplicates
    clouds = [rng.normal(0, 3, (500, 3)).astype(np.float32), np.zeros((0, 3), np.float32),
              np.repeat(rng.normal(0, 1, (50, 3)).astype(np.float32), 4, axis=0), rng.uniform(-50, 50, (1200, 3)).astype(np.float32)]
    raw = np.concatenate(clouds); lens = np.array([len(c) for c in clouds], np.int32)
    for dl in (0.05, 0.7, 40.0):
        po, lo = oracle.subsample_batch(raw, lens, sampleDl=dl)
        pg, lg = ops.grid_subsample(_dev(raw, cuda), _dev(lens, cuda), dl)
        assert np.array_equal(lg.cpu().numpy(), lo) and np.array_equal(pg.cpu().numpy(), po), f"dl={dl}"
    # max_p truncation keeps the first max_p voxels per cloud in canonical order
    po, lo = oracle.subsample_batch(raw, lens, sampleDl=0.7, max_p=100)
    pg, lg = ops.grid_subsample(_dev(raw, cuda), _dev(lens, cuda), 0.7, max_p=100)
    assert np.array_equal(lg.cpu().numpy(), lo) and np.array_equal(pg.cpu().numpy(), po)
    # idempotence-like property: every barycentre lies in its voxel, so re-subsampling keeps the count
    pg2, lg2 = ops.grid_subsample(pg, lg, 0.7)
    assert lg2.tolist() == lg.tolist()


def test_subsample_features_mean(cuda):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-5, 5, (3000, 3)).astype(np.float32)
    feats = rng.normal(0, 1, (3000, 5)).astype(np.float32)
    lens = np.array([1000, 2000], np.int32)
    p, l, f = ops.grid_subsample(_dev(pts, cuda), _dev(lens, cuda), 1.0, features=_dev(feats, cuda))
    # reference semantics: per-voxel mean of features in the same voxel grouping as the points
    # check via augmenting: the feature mean of the xyz columns themselves equals the barycentre
    p2, l2, f2 = ops.grid_subsample(_dev(pts, cuda), _dev(lens, cuda), 1.0, features=_dev(pts, cuda))
    assert torch.allclose(p2, f2, rtol=1e-6, atol=1e-6)
    assert f.shape == (p.shape[0], 5) and torch.isfinite(f).all()


@pytest.mark.parametrize("name,r", [("conv0", 1.275), ("pool0", 1.275), ("up0", 2.55)])
def test_neighbors_golden_bit_exact(cuda, gold_l1, name, r):
    p0, l0, p1, l1 = gold_l1["p0"], gold_l1["l0"], gold_l1["p1"], gold_l1["l1"]
    q, s, ql, sl = {"conv0": (p0, p0, l0, l0), "pool0": (p1, p0, l1, l0), "up0": (p0, p1, l0, l1)}[name]
    gold = gold_l1[f"nn_{name}_ordered"]
    w = gold.shape[1]
    idx, counts, maxc = ops.radius_neighbors(_dev(q, cuda), _dev(s, cuda), _dev(ql, cuda), _dev(sl, cuda), r, w,
                                             want_counts=True)
    assert int(maxc.item()) == w                                         # reference output width (neighbors.cpp:296-304)
    assert np.array_equal(idx.cpu().numpy(), gold)                       # bit-exact incl. order and shadow padding
    assert np.array_equal(counts.cpu().numpy(), (gold < len(s)).sum(1))
    ok, _, nontie = equal_modulo_ties(idx.cpu().numpy(), gold_l1[f"nn_{name}_nanoflann"], q, s)
    assert ok and nontie == 0
    # truncated (the hot-path call): leading columns of the full result
    for lim in (1, 7, 20):
        t = ops.radius_neighbors(_dev(q, cuda), _dev(s, cuda), _dev(ql, cuda), _dev(sl, cuda), r, lim)
        assert np.array_equal(t.cpu().numpy(), gold[:, :lim])


def test_neighbors_kitti_pyramid_vs_oracle(cuda, oracle):
    """Full KITTI-shaped pair, all 10 searches of the pyramid at the calibrated width, bit-exact vs the oracle."""
    cfg = kitti_config()
    a, b = synth.pair_raw(1)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p, l = oracle.subsample_batch(raw, lens, sampleDl=0.3)
    r, lim = 0.3 * 4.25, 57
    for level in range(4):
        dq, dl_ = _dev(p, cuda), _dev(l, cuda)
        got = ops.radius_neighbors(dq, dq, dl_, dl_, r, lim).cpu().numpy()
        want = oracle.batch_query(p, p, l, l, radius=r, max_neighbors=lim)
        assert got.shape[0] == want.shape[0]
        assert np.array_equal(got[:, :want.shape[1]], want), f"conv level {level}"
        assert np.all(got[:, want.shape[1]:] == len(p))
        if level == 3:
            break
        p2, l2 = oracle.subsample_batch(p, l, sampleDl=0.6 * 2 ** level)
        dq2, dl2 = _dev(p2, cuda), _dev(l2, cuda)
        got = ops.radius_neighbors(dq2, dq, dl2, dl_, r, lim).cpu().numpy()
        want = oracle.batch_query(p2, p, l2, l, radius=r, max_neighbors=lim)
        assert np.array_equal(got[:, :want.shape[1]], want), f"pool level {level}"
        got = ops.radius_neighbors(dq, dq2, dl_, dl2, 2 * r, lim).cpu().numpy()
        want = oracle.batch_query(p, p2, l, l2, radius=2 * r, max_neighbors=lim)
        assert np.array_equal(got[:, :want.shape[1]], want), f"upsample level {level}"
        p, l, r = p2, l2, 2 * r


def test_neighbors_edge_cases(cuda, oracle):
    rng = np.random.default_rng(3)
    # ragged batch, one element with no supports, one with no queries; queries far outside the support bbox
    qs = [rng.uniform(-4, 4, (300, 3)), rng.uniform(-4, 4, (40, 3)), np.zeros((0, 3)), rng.uniform(50, 60, (25, 3))]
    ss = [rng.uniform(-4, 4, (500, 3)), np.zeros((0, 3)), rng.uniform(-1, 1, (30, 3)), rng.uniform(-4, 4, (100, 3))]
    q = np.concatenate(qs).astype(np.float32); s = np.concatenate(ss).astype(np.float32)
    ql = np.array([len(x) for x in qs], np.int32); sl = np.array([len(x) for x in ss], np.int32)
    for r in (0.3, 1.0, 3.0):
        want, cnt = oracle.batch_query(q, s, ql, sl, radius=r, return_counts=True)
        w = max(want.shape[1], 1)
        got, counts, maxc = ops.radius_neighbors(_dev(q, cuda), _dev(s, cuda), _dev(ql, cuda), _dev(sl, cuda), r, w,
                                                 want_counts=True)
        assert int(maxc.item()) == want.shape[1]
        if want.shape[1]:
            assert np.array_equal(got.cpu().numpy(), want)
        assert np.array_equal(counts.cpu().numpy(), cnt)
    # radius larger than the cloud: every support of the element is a neighbour (wide rows, multi-round selection)
    q1 = rng.uniform(-1, 1, (64, 3)).astype(np.float32); s1 = rng.uniform(-1, 1, (700, 3)).astype(np.float32)
    l1, l2 = np.array([64], np.int32), np.array([700], np.int32)
    want = oracle.batch_query(q1, s1, l1, l2, radius=10.0)
    assert want.shape[1] == 700
    got = ops.radius_neighbors(_dev(q1, cuda), _dev(s1, cuda), _dev(l1, cuda), _dev(l2, cuda), 10.0, 700)
    assert np.array_equal(got.cpu().numpy(), want)
    got = ops.radius_neighbors(_dev(q1, cuda), _dev(s1, cuda), _dev(l1, cuda), _dev(l2, cuda), 10.0, 33)  # forces truncation rounds
    assert np.array_equal(got.cpu().numpy(), want[:, :33])
    # exact duplicates -> exact d2 ties, broken by ascending support index
    sd = np.repeat(rng.uniform(-1, 1, (40, 3)).astype(np.float32), 3, axis=0)
    ld = np.array([len(sd)], np.int32)
    want = oracle.batch_query(sd, sd, ld, ld, radius=0.5)
    got = ops.radius_neighbors(_dev(sd, cuda), _dev(sd, cuda), _dev(ld, cuda), _dev(ld, cuda), 0.5, want.shape[1])
    assert np.array_equal(got.cpu().numpy(), want)


def test_neighbors_invariants_at_full_size(cuda):
    """Size-independent properties on a full KITTI pair (no oracle needed)."""
    a, b = synth.pair_raw(2)
    raw = _dev(np.concatenate([a, b]), cuda); lens = _dev(np.array([len(a), len(b)], np.int32), cuda)
    p, l = ops.grid_subsample(raw, lens, 0.3)
    idx, counts, maxc = ops.radius_neighbors(p, p, l, l, 1.275, 57, want_counts=True)
    n = p.shape[0]
    i, pn, ln = idx.cpu().numpy(), p.cpu().numpy(), l.cpu().numpy()
    assert np.array_equal(i[:, 0], np.arange(n))                          # self first
    d2 = d2_rows(pn, pn, i)
    assert np.all(np.diff(np.where(np.isinf(d2), np.float32(3e38), d2), axis=1) >= 0)
    assert np.all(d2[i < n] < np.float32(1.275) * np.float32(1.275))
    valid = (i < n).sum(1)
    assert np.array_equal(valid, np.minimum(counts.cpu().numpy(), 57))
    assert np.all((i < n) == (np.arange(57)[None, :] < valid[:, None]))   # pads contiguous at the row end
    assert np.all(i[:ln[0]][i[:ln[0]] < n] < ln[0]) and np.all(i[ln[0]:] >= ln[0])
    assert int(maxc.item()) == counts.max().item()


def test_dropin_modules_numpy_in_numpy_out(cuda, gold_l1):
    import cpp_wrappers.cpp_neighbors.radius_neighbors as cpp_neighbors
    import cpp_wrappers.cpp_subsampling.grid_subsampling as cpp_subsampling
    raw, lens = gold_l1["raw"], gold_l1["lens"]
    p, l = cpp_subsampling.subsample_batch(torch.from_numpy(raw).double(), torch.from_numpy(lens), sampleDl=0.3)
    assert isinstance(p, np.ndarray) and p.dtype == np.float32 and l.dtype == np.int32
    assert np.array_equal(p, gold_l1["sub_0.3_points_canonical"]) and np.array_equal(l, gold_l1["sub_0.3_lens"])
    single = cpp_subsampling.subsample(raw[:lens[0]], sampleDl=0.3)
    assert np.array_equal(single, p[:l[0]])
    nn = cpp_neighbors.batch_query(p, p, l, l, radius=1.275)
    assert nn.dtype == np.int32 and np.array_equal(nn, gold_l1["nn_conv0_ordered"])
    # error behaviour of the reference wrappers (cpp_neighbors/wrapper.cpp:127-171, cpp_subsampling/wrapper.cpp:92-96)
    with pytest.raises(RuntimeError, match="query.shape is not"):
        cpp_neighbors.batch_query(p[:, :2], p, l, l, radius=1.0)
    with pytest.raises(RuntimeError, match="different for queries and supports"):
        cpp_neighbors.batch_query(p, p, l, l[:1], radius=1.0)
    with pytest.raises(RuntimeError, match="^Error$"):
        cpp_neighbors.batch_query(p + 1000.0, p, l, l, radius=0.01)       # no query has any neighbour
    with pytest.raises(RuntimeError, match="Error parsing method"):
        cpp_subsampling.subsample_batch(raw, lens, sampleDl=0.3, method="centroids")
    with pytest.raises(RuntimeError, match="points.shape is not"):
        cpp_subsampling.subsample_batch(raw[:, :2], lens, sampleDl=0.3)
    with pytest.raises(TypeError):
        cpp_neighbors.batch_query(p, p, l, l, 1.0)                        # radius is keyword-only ("OOOO|$f")


def test_collate_host_and_device_pyramids_agree_with_oracle(cuda, oracle):
    from apr_b200 import dataloader as dl
    from oracle.ref import collate_ref
    cfg = kitti_config()
    a, b = synth.small_cloud(31, 2500), synth.small_cloud(32, 2200)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p0, l0 = oracle.subsample_batch(raw, lens, sampleDl=0.3)
    limits = [24, 26, 28, 30]
    want = collate_ref(p0, l0, cfg, limits, oracle.subsample_batch, oracle.batch_query)
    host = dl.collate_fn_descriptor(dl.make_list_data(p0[:l0[0]], p0[l0[0]:]), cfg, limits)
    devp = dl.build_pyramid_device(_dev(p0, cuda), _dev(l0, cuda), cfg, limits)
    for k in ("points", "neighbors", "pools", "upsamples"):
        for lvl in range(4):
            w = want[k][lvl]
            h = host[k][lvl].numpy()
            assert h.shape == w.shape and np.array_equal(h, w), (k, lvl)
            d = devp[k][lvl].cpu().numpy()
            if k == "points" or w.shape[0] == 0:
                assert d.shape[0] == w.shape[0] and (k != "points" or np.array_equal(d, w))
            else:                                                          # device matrices have fixed width = limit
                assert np.array_equal(d[:, :w.shape[1]], w) and np.all(d[:, w.shape[1]:] == len(want["points"][lvl if k != "upsamples" else lvl + 1]))
    assert host["neighbors"][0].dtype == torch.int64 and host["features"].shape == (len(p0), 1)
    lim = dl.calibrate_neighbors([dl.make_list_data(p0[:l0[0]], p0[l0[0]:])[0]], cfg, dl.collate_fn_descriptor)
    from oracle.ref import calibrate_ref
    assert np.array_equal(lim, calibrate_ref([(p0, l0)], cfg, oracle.subsample_batch, oracle.batch_query))


def test_voxel_downsample_raw_open3d_semantics(cuda):
    """First-level voxelisation of raw [n,4] scans: the CUDA path vs the numpy restatement of open3d 0.10's
    VoxelDownSample (oracle/o3d_voxel_ref.py — parity unpinned: open3d itself is absent) — bit-exact, raw KITTI-shaped
    scan pair with a reflectance column, plus an empty cloud in the batch; and the voxel grid differs from
    grid_subsampling's (origin at min - 0.5 voxel instead of floor(min/dl)*dl)."""
    import torch
    from apr_b200 import ops, synth
    from oracle.o3d_voxel_ref import voxel_down_sample_batch_ref
    a, b = synth.pair_raw(2, "kitti")
    a, b = a[:30000], b[:25000]
    rng = np.random.default_rng(0)
    raw = np.concatenate([a, np.zeros((0, 3), np.float32), b])
    raw4 = np.concatenate([raw, rng.random((len(raw), 1), dtype=np.float32)], 1)          # x, y, z, reflectance
    lens = np.array([len(a), 0, len(b)], np.int32)
    want_p, want_l = voxel_down_sample_batch_ref(raw4, lens, 0.3)
    got_p, got_l = ops.voxel_downsample_raw(torch.from_numpy(raw4).to(cuda), torch.from_numpy(lens).to(cuda), 0.3)
    assert np.array_equal(got_l.cpu().numpy(), want_l)
    assert np.array_equal(got_p.cpu().numpy(), want_p)
    ref_p, ref_l = ops.grid_subsample(torch.from_numpy(raw).to(cuda), torch.from_numpy(lens).to(cuda), 0.3)
    assert abs(int(ref_l.sum()) - int(want_l.sum())) < 0.05 * int(want_l.sum())              # same density, different grid
    assert ref_p.shape != got_p.shape or not torch.equal(ref_p, got_p)


def test_subsample_label_vote_vs_reference(cuda, ref_l1):
    """subsample_batch(classes=...) (grid_subsampling.cpp:63-68, :96-101): the drop-in's labels against the reference's
    object code, rows matched through their barycentres; where the vote is tied the smallest label must win here."""
    from apr_b200.cpp_wrappers.cpp_subsampling import grid_subsampling as gs
    from test_oracle import _vote_case, check_votes, voxel_votes
    if not hasattr(ref_l1.lib, "ref_batch_grid_subsampling_full"):
        pytest.skip("oracle/_ref predates the label shim")
    for seed, n, lens, ldim, nl in ((0, 6000, [2500, 3500], 1, 3), (1, 5000, [5000], 2, 4), (2, 64, [64], 1, 2)):
        pts, lens, feats, cls = _vote_case(seed, n, lens, ldim, nl)
        rp, rl, rf, rc = ref_l1.subsample_batch_full(pts, lens, feats, cls, sampleDl=1.5)
        p, l, f, c = gs.subsample_batch(pts, lens, features=feats, classes=cls if ldim > 1 else cls[:, 0], sampleDl=1.5)
        assert c.dtype == np.int32 and c.shape == (len(p), ldim) and l.tolist() == rl.tolist()
        check_votes(pts, lens, cls, 1.5, p, l, c)
        votes = voxel_votes(pts, lens, cls, 1.5)
        # order-free comparison with the reference: sort both by barycentre
        off, n_differ = 0, 0
        for m in l:
            a = np.lexsort((p[off:off + m, 2], p[off:off + m, 1], p[off:off + m, 0])) + off
            b = np.lexsort((rp[off:off + m, 2], rp[off:off + m, 1], rp[off:off + m, 0])) + off
            assert np.array_equal(p[a], rp[b])
            assert np.array_equal(f[a].view(np.int32), rf[b].view(np.int32))
            differ = (c[a] != rc[b])
            assert (c[a][differ] < rc[b][differ]).all()          # only on ties, and then the smaller label
            n_differ += int(differ.sum())
            off += m
        tied = sum(1 for _, _, cnts in votes.values() for cn in cnts if sorted(cn.values())[-1:] == sorted(cn.values())[-2:-1])
        assert n_differ <= tied
    # single-cloud entry point: (points, classes)
    pts, lens, feats, cls = _vote_case(5, 300, [300], 1, 3)
    sp, sc = gs.subsample(pts, classes=cls, sampleDl=1.0)
    assert sc.shape == (len(sp), 1)


def test_nearest_only_search_is_column0_of_the_full_search(cuda, oracle):
    """aprb_cell_grid_query_nearest (the upsample search of the device-resident pyramid): bit-identical to
    column 0 of the full sorted matrix — upsample shapes (queries = finer level, supports = coarser level, radius 2r), empty
    balls (pad), queries far outside the support bounding box, an empty cloud, exact d2 ties (duplicated supports)."""
    rng = np.random.default_rng(3)
    a, b = synth.pair_raw(1)
    p0, l0 = oracle.subsample_batch(np.concatenate([a, b]), np.array([len(a), len(b)], np.int32), 0.3)
    p1, l1 = oracle.subsample_batch(p0, l0, 0.6)
    cases = [(p0, l0, p1, l1, 2.55), (p1, l1, p0, l0, 1.275), (p0, l0, p1, l1, 0.4)]
    far = np.concatenate([p0[:50] + 500.0, p0[50:100]]).astype(np.float32)
    cases.append((far, np.array([60, 40], np.int32), p1, l1, 2.55))
    dup = np.concatenate([p1[:200], p1[:200]]).astype(np.float32)                  # every support twice: d2 ties, lower index wins
    cases.append((p0[:300], np.array([300], np.int32), dup, np.array([400], np.int32), 2.55))
    cases.append((p0[:100], np.array([60, 40], np.int32), p1[:80], np.array([0, 80], np.int32), 2.55))   # first cloud has no supports
    for q, ql, s, sl, r in cases:
        grid = ops.CellGrid(_dev(s, cuda), _dev(sl, cuda), r)
        full = grid.query(_dev(q, cuda), _dev(ql, cuda), r, 40)
        near = grid.query_nearest(_dev(q, cuda), _dev(ql, cuda), r)
        assert near.shape == (len(q), 1) and torch.equal(near[:, 0], full[:, 0])
        want = oracle.batch_query(q, s, ql, sl, radius=r)
        col0 = want[:, 0] if want.shape[1] else np.full(len(q), len(s), np.int32)
        assert np.array_equal(near[:, 0].cpu().numpy(), col0)
