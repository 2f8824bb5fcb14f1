"""CPU: pins the oracle (oracle/oracle_l1.c, oracle/blocks_ref.py) against the golden vectors produced by the real
reference (oracle/make_golden.py) and, when oracle/_ref is present, against the reference's object code directly."""
import numpy as np
import pytest
import torch

from apr_b200 import synth
from apr_b200.config import kitti_config
from oracle import blocks_ref
from oracle.ref import collate_ref, calibrate_ref, equal_modulo_ties, lexsort_rows_per_cloud, d2_rows


def test_subsample_matches_reference_golden(oracle, gold_l1):
    raw, lens = gold_l1["raw"], gold_l1["lens"]
    for dl in (0.3, 0.6):
        p, l = oracle.subsample_batch(raw, lens, sampleDl=dl)
        assert np.array_equal(l, gold_l1[f"sub_{dl}_lens"])
        assert np.array_equal(p, gold_l1[f"sub_{dl}_points_canonical"])
        # as a set, bit-identical to the reference's (unordered_map-ordered) output
        assert np.array_equal(lexsort_rows_per_cloud(p, l), gold_l1[f"sub_{dl}_points_ref_lexsorted"])


@pytest.mark.parametrize("name,r", [("conv0", 1.275), ("pool0", 1.275), ("up0", 2.55)])
def test_neighbors_match_reference_golden(oracle, gold_l1, name, r):
    p0, l0, p1, l1 = gold_l1["p0"], gold_l1["l0"], gold_l1["p1"], gold_l1["l1"]
    q, s, ql, sl = {"conv0": (p0, p0, l0, l0), "pool0": (p1, p0, l1, l0), "up0": (p0, p1, l0, l1)}[name]
    nn = oracle.batch_query(q, s, ql, sl, radius=r)
    assert np.array_equal(nn, gold_l1[f"nn_{name}_ordered"])              # bit-exact vs batch_ordered_neighbors
    ok, rows, nontie = equal_modulo_ties(nn, gold_l1[f"nn_{name}_nanoflann"], q, s)
    assert ok and nontie == 0                                             # nanoflann differs only inside d2 ties


def test_neighbor_invariants(oracle, gold_l1):
    p0, l0 = gold_l1["p0"], gold_l1["l0"]
    nn, counts = oracle.batch_query(p0, p0, l0, l0, radius=1.275, return_counts=True)
    ns = len(p0)
    d2 = d2_rows(p0, p0, nn)
    assert np.all(np.diff(np.where(np.isinf(d2), np.float32(3e38), d2), axis=1) >= 0)   # ascending, pads last
    assert np.array_equal(nn[:, 0], np.arange(ns))                        # self is the nearest neighbour
    assert np.array_equal((nn < ns).sum(1), counts)
    r2 = np.float32(1.275) * np.float32(1.275)
    assert np.all(d2[nn < ns] < r2)
    off = np.concatenate([[0], np.cumsum(l0)])
    for b in range(len(l0)):                                              # indices stay inside their own cloud
        blk = nn[off[b]:off[b + 1]]
        v = blk[blk < ns]
        assert v.min() >= off[b] and v.max() < off[b + 1]
    # truncation == leading columns (dataloader.py:66-70)
    assert np.array_equal(oracle.batch_query(p0, p0, l0, l0, radius=1.275, max_neighbors=10), nn[:, :10])


def test_oracle_vs_reference_object_code(oracle, ref_l1):
    a, b = synth.small_cloud(11, 900), synth.small_cloud(12, 700)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    pr, lr = ref_l1.subsample_batch(raw, lens, sampleDl=0.45)
    po, lo = oracle.subsample_batch(raw, lens, sampleDl=0.45)
    assert np.array_equal(lr, lo)
    assert np.array_equal(lexsort_rows_per_cloud(pr, lr), lexsort_rows_per_cloud(po, lo))
    nn_o = oracle.batch_query(po, po, lo, lo, radius=1.9)
    assert np.array_equal(nn_o, ref_l1.batch_query(po, po, lo, lo, radius=1.9, variant="ordered"))
    ok, _, nontie = equal_modulo_ties(nn_o, ref_l1.batch_query(po, po, lo, lo, radius=1.9), po, po)
    assert ok and nontie == 0


def test_kpconv_restatement_matches_reference_golden(gold_kpconv):
    g = gold_kpconv
    p0, p1 = torch.from_numpy(g["p0"]), torch.from_numpy(g["p1"])
    conv, pool = torch.from_numpy(g["conv"]).long(), torch.from_numpy(g["pool"]).long()
    for tag, strided in (("c1", False), ("c8", False), ("c32s", True), ("c64", False)):
        q, s, inds = (p1, p0, pool) if strided else (p0, p0, conv)
        y = blocks_ref.kpconv_ref(q, s, inds, torch.from_numpy(g[f"{tag}_x"]), torch.from_numpy(g[f"{tag}_kp"]),
                                  torch.from_numpy(g[f"{tag}_W"]), 0.6)
        ref = torch.from_numpy(g[f"{tag}_y"])
        assert (y - ref).norm() / ref.norm() < 1e-6
    assert torch.equal(blocks_ref.max_pool_ref(torch.from_numpy(g["pool_x"]), pool), torch.from_numpy(g["max_pool_y"]))
    assert torch.equal(blocks_ref.closest_pool_ref(torch.from_numpy(g["closest_x"]), torch.from_numpy(g["up"]).long()),
                       torch.from_numpy(g["closest_pool_y"]))


def test_encoder_restatement_matches_reference_golden(oracle, gold_encoder):
    g = gold_encoder
    cfg = kitti_config(first_feats_dim=16)
    pyr = collate_ref(g["p0"], g["l0"], cfg, list(g["limits"]), oracle.subsample_batch, oracle.batch_query)
    batch = dict(points=[torch.from_numpy(p) for p in pyr["points"]],
                 neighbors=[torch.from_numpy(n).long() for n in pyr["neighbors"]],
                 pools=[torch.from_numpy(n).long() for n in pyr["pools"]],
                 features=torch.ones(len(g["p0"]), 1))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    outs = blocks_ref.encoder_ref(batch, sd, cfg, return_all=True)
    ref = torch.from_numpy(g["y_final"])
    assert (outs[-1] - ref).norm() / ref.norm() < 1e-5
    assert np.allclose([o.norm().item() for o in outs], g["block_norms"], rtol=1e-5)


def _kpfcnn_batch(oracle, g, cfg):
    pyr = collate_ref(g["p0"], g["l0"], cfg, list(g["limits"]), oracle.subsample_batch, oracle.batch_query)
    return dict(points=[torch.from_numpy(p) for p in pyr["points"]],
                neighbors=[torch.from_numpy(n).long() for n in pyr["neighbors"]],
                pools=[torch.from_numpy(n).long() for n in pyr["pools"]],
                upsamples=[torch.from_numpy(n).long() for n in pyr["upsamples"]],
                stack_lengths=[torch.from_numpy(l) for l in pyr["stack_lengths"]],
                features=torch.ones(len(g["p0"]), 1))


def test_kpfcnn_restatement_matches_reference_golden(oracle, gold_kpfcnn):
    """Full KPFCNN.forward restatement (encoder + bottleneck GNN + decoder) vs the outputs of the REAL reference."""
    g = gold_kpfcnn
    cfg = kitti_config(first_feats_dim=16, gnn_feats_dim=32, final_feats_dim=8)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    ff, so, ss = blocks_ref.kpfcnn_ref(_kpfcnn_batch(oracle, g, cfg), sd, cfg)
    for got, key in ((ff, "feats_f"), (so, "scores_overlap"), (ss, "scores_saliency")):
        ref = torch.from_numpy(g[key])
        assert got.shape == ref.shape and (got - ref).norm() / ref.norm() < 1e-4, key


def test_gcn_module_matches_restatement_on_cpu(gold_kpfcnn):
    """apr_b200.gcn (stock torch, factored edge convolution, row-major points) == the reference-shaped restatement, with
    the reference's own `gnn.*` parameters loaded by name."""
    from apr_b200.gcn import GCN
    g = gold_kpfcnn
    cfg = kitti_config(first_feats_dim=16, gnn_feats_dim=32, final_feats_dim=8)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    net = GCN(cfg.num_head, cfg.gnn_feats_dim, cfg.dgcnn_k, cfg.nets)
    net.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("gnn.")}, strict=True)
    gen = torch.Generator().manual_seed(0)
    c0, c1 = torch.rand(90, 3, generator=gen) * 20, torch.rand(70, 3, generator=gen) * 20
    d0, d1 = torch.randn(90, 32, generator=gen), torch.randn(70, 32, generator=gen)
    with torch.no_grad():
        a0, a1 = net(c0, c1, d0, d1)
    r0, r1 = d0.t(), d1.t()
    for li, name in enumerate(cfg.nets):
        p = f"gnn.layers.{li}."
        if name == "self":
            r0 = blocks_ref.self_attention_ref(c0, r0, sd, p, cfg.dgcnn_k)
            r1 = blocks_ref.self_attention_ref(c1, r1, sd, p, cfg.dgcnn_k)
        else:
            r0 = r0 + blocks_ref.cross_attention_ref(r0, r1, sd, p, cfg.num_head)
            r1 = r1 + blocks_ref.cross_attention_ref(r1, r0, sd, p, cfg.num_head)
    assert (a0 - r0.t()).abs().max() < 1e-4 and (a1 - r1.t()).abs().max() < 1e-4


def test_pyramid_schedule_and_calibration(oracle):
    cfg = kitti_config()
    a, b = synth.small_cloud(21, 1500), synth.small_cloud(22, 1500)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p0, l0 = oracle.subsample_batch(raw, lens, sampleDl=0.3)
    pyr = collate_ref(p0, l0, cfg, [20, 20, 20, 20], oracle.subsample_batch, oracle.batch_query)
    assert len(pyr["points"]) == 4 and pyr["pools"][3].shape[0] == 0 and pyr["upsamples"][3].shape[0] == 0
    for l in range(3):
        assert pyr["pools"][l].shape[0] == len(pyr["points"][l + 1])
        assert pyr["upsamples"][l].shape[0] == len(pyr["points"][l])
        assert pyr["neighbors"][l].shape[1] <= 20
    lims = calibrate_ref([(p0, l0)], cfg, oracle.subsample_batch, oracle.batch_query)
    assert lims.shape == (4,) and np.all(lims > 0)


def test_load_kernels_matches_reference_golden():
    """apr_b200.kernel_points.load_kernels == the reference's load_kernels (kernels/kernel_points.py:388-470) bit for bit
    under the same numpy global seed and call order (fixture: oracle/make_golden_kernel_points.py, run on the reference)."""
    import os
    from apr_b200.kernel_points import load_kernels
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kernel_points.npz"))
    for seed in (0, 1, 12345):
        np.random.seed(seed)
        for i, radius in enumerate((1.275, 2.55, 5.1, 10.2)):
            kp = load_kernels(radius, 15, dimension=3, fixed='center')
            want = g[f"s{seed}_{i}"]
            assert kp.dtype == np.float32 and kp.shape == (15, 3)
            assert np.array_equal(kp, want), (seed, i)
    with pytest.raises(NotImplementedError):
        load_kernels(1.0, 13, dimension=3, fixed='center')


def _vote_case(seed, n, lens, ldim, n_labels):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    feats = rng.normal(0, 1, (n, 3)).astype(np.float32)
    cls = rng.integers(0, n_labels, (n, ldim)).astype(np.int32)
    return pts, np.asarray(lens, np.int32), feats, cls


def voxel_votes(pts, lens, cls, dl):
    """numpy restatement of the label vote (grid_subsampling.cpp:27, :49-53 voxel of a point; grid_subsampling.h:62-75
    counts; grid_subsampling.cpp:96-101 max count): {(cloud, voxel) -> (barycentre fp64, [Counter per label dim])}."""
    from collections import Counter
    out, off = {}, 0
    for b, n in enumerate(lens):
        p = pts[off:off + n]
        if n:
            dlf = np.float32(dl)
            origin = np.floor(p.min(0) * (np.float32(1) / dlf)) * dlf
            vox = np.floor((p - origin) / dlf).astype(np.int64)
            for i in range(n):
                key = (b,) + tuple(vox[i])
                ent = out.setdefault(key, [np.zeros(3), 0, [Counter() for _ in range(cls.shape[1])]])
                ent[0] += p[i]; ent[1] += 1
                for d in range(cls.shape[1]):
                    ent[2][d][int(cls[off + i, d])] += 1
        off += n
    return out


def check_votes(pts, lens, cls, dl, out_pts, out_lens, out_cls):
    """Every output row's label is one of the most frequent labels of the voxel its barycentre lies in."""
    votes = voxel_votes(pts, lens, cls, dl)
    by_cloud = {}
    for key, (s, c, cnts) in votes.items():
        by_cloud.setdefault(key[0], []).append((s / c, cnts))
    assert int(np.sum(out_lens)) == len(out_pts) == len(out_cls) == len(votes)
    off = 0
    for b, m in enumerate(out_lens):
        cand = by_cloud.get(b, [])
        cen = np.stack([c[0] for c in cand]) if cand else np.zeros((0, 3))
        for r in range(off, off + m):
            j = int(np.argmin(np.abs(cen - out_pts[r]).sum(1)))
            assert np.abs(cen[j] - out_pts[r]).max() < 1e-4
            for d in range(out_cls.shape[1]):
                cnt = cand[j][1][d]
                assert cnt[int(out_cls[r, d])] == max(cnt.values())
        off += m


def test_reference_label_vote_and_feature_mean_semantics(ref_l1):
    """Pins the restated semantics of update_all / the max_element vote on the reference's own object code."""
    if not hasattr(ref_l1.lib, "ref_batch_grid_subsampling_full"):
        pytest.skip("oracle/_ref predates the label shim")
    for seed, n, lens, ldim, nl in ((0, 600, [250, 350], 1, 3), (1, 500, [500], 2, 4), (2, 64, [64], 1, 2)):
        pts, lens, feats, cls = _vote_case(seed, n, lens, ldim, nl)
        op, ol, of, oc = ref_l1.subsample_batch_full(pts, lens, feats, cls, sampleDl=1.5)
        check_votes(pts, lens, cls, 1.5, op, ol, oc)
        op2, ol2 = ref_l1.subsample_batch(pts, lens, sampleDl=1.5)
        assert np.array_equal(op, op2) and np.array_equal(ol, ol2) and of.shape == (len(op), 3)
