"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the REAL reference in this container:
  * the reference's object code (oracle/_ref/libapr_ref.so) for subsampling / radius search;
  * the reference's Python modules imported from /root/reference/Predator_APR (models/blocks.py,
    models/architectures.py) for KPConv / blocks / the KFE encoder.
While doing so it pins the oracle restatements (oracle/oracle_l1.c, oracle/blocks_ref.py) against the reference and
aborts on any disagreement. Run from the repo root:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
The GPU box has no /root/reference; tests there use the committed fixtures.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Predator_APR"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from apr_b200 import synth                                   # noqa: E402
from apr_b200.config import AttrDict, kitti_config           # noqa: E402
from oracle import blocks_ref                                # noqa: E402
from oracle.ref import Oracle, RefL1, collate_ref, equal_modulo_ties, lexsort_rows_per_cloud  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    os.chdir(REF)                                            # load_kernels opens a relative path (kernel_points.py:391)
    sys.path.insert(0, REF)
    import models.blocks as rblocks
    import models.architectures as rarch
    return rblocks, rarch


def small_pair(seed, n=1500):
    a = synth.small_cloud(seed, n)
    b = synth.small_cloud(seed + 100, n - 200) + np.array([1.5, 0.3, 0.0], np.float32)
    return a, b


def golden_l1(O, R):
    a, b = small_pair(1)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    out = {}
    for dl in (0.3, 0.6):
        pr, lr = R.subsample_batch(raw, lens, sampleDl=dl)
        po, lo = O.subsample_batch(raw, lens, sampleDl=dl)
        assert np.array_equal(lr, lo)
        assert np.array_equal(lexsort_rows_per_cloud(pr, lr), lexsort_rows_per_cloud(po, lo)), "oracle subsample != reference"
        out[f"sub_{dl}_points_canonical"] = po               # ascending voxel order (oracle == reference as a set)
        out[f"sub_{dl}_points_ref_lexsorted"] = lexsort_rows_per_cloud(pr, lr)
        out[f"sub_{dl}_lens"] = lr
    p0, l0 = O.subsample_batch(raw, lens, sampleDl=0.3)
    p1, l1 = O.subsample_batch(p0, l0, sampleDl=0.6)
    cases = [("conv0", p0, p0, l0, l0, 1.275), ("pool0", p1, p0, l1, l0, 1.275), ("up0", p0, p1, l0, l1, 2.55)]
    for name, q, s, ql, sl, r in cases:
        nn_nano = R.batch_query(q, s, ql, sl, radius=r)
        nn_ord = R.batch_query(q, s, ql, sl, radius=r, variant="ordered")
        nn_orc = O.batch_query(q, s, ql, sl, radius=r)
        assert np.array_equal(nn_ord, nn_orc), f"{name}: oracle != batch_ordered_neighbors"
        ok, rows, nontie = equal_modulo_ties(nn_orc, nn_nano, q, s)
        assert ok, f"{name}: oracle vs nanoflann differ outside tie runs ({nontie})"
        out[f"nn_{name}_ordered"] = nn_ord.astype(np.int32)
        out[f"nn_{name}_nanoflann"] = nn_nano.astype(np.int32)
        print(f"  {name}: width {nn_ord.shape[1]}, rows differing from nanoflann only inside ties: {rows}")
    out.update(raw=raw, lens=lens, p0=p0, l0=l0, p1=p1, l1=l1)
    np.savez_compressed(os.path.join(GOLD, "l1_small.npz"), **out)


def batch_to_torch(b):
    return dict(points=[torch.from_numpy(p) for p in b['points']],
                neighbors=[torch.from_numpy(n.astype(np.int64)) for n in b['neighbors']],
                pools=[torch.from_numpy(n.astype(np.int64)) for n in b['pools']],
                upsamples=[torch.from_numpy(n.astype(np.int64)) for n in b['upsamples']],
                stack_lengths=[torch.from_numpy(l) for l in b['stack_lengths']])


def golden_kpconv(O, rblocks):
    a, b = small_pair(2, 1200)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p0, l0 = O.subsample_batch(raw, lens, sampleDl=0.3)
    p1, l1 = O.subsample_batch(p0, l0, sampleDl=0.6)
    conv = O.batch_query(p0, p0, l0, l0, radius=1.275, max_neighbors=40)
    pool = O.batch_query(p1, p0, l1, l0, radius=1.275, max_neighbors=40)
    out = dict(p0=p0, p1=p1, conv=conv, pool=pool)
    g = torch.Generator().manual_seed(0)
    for tag, (cin, cout, strided) in {"c1": (1, 16, False), "c8": (8, 12, False), "c32s": (32, 32, True),
                                      "c64": (64, 48, False)}.items():
        torch.manual_seed(1); np.random.seed(1)
        m = rblocks.KPConv(15, 3, cin, cout, 0.6, 1.275)
        q, s, inds = (p1, p0, pool) if strided else (p0, p0, conv)
        x = torch.randn(len(s), cin, generator=g) if cin > 1 else torch.ones(len(s), 1)
        qt, st, it = torch.from_numpy(q), torch.from_numpy(s), torch.from_numpy(inds.astype(np.int64))
        with torch.no_grad():
            y = m(qt, st, it, x)
            y2 = blocks_ref.kpconv_ref(qt, st, it, x, m.kernel_points, m.weights, 0.6)
        err = (y - y2).norm() / y.norm()
        assert err < 1e-6, f"kpconv_ref != reference KPConv ({err})"
        out.update({f"{tag}_x": x.numpy(), f"{tag}_kp": m.kernel_points.detach().numpy(),
                    f"{tag}_W": m.weights.detach().numpy(), f"{tag}_y": y.numpy()})
        print(f"  kpconv {tag}: restatement rel err {err:.2e}")
    # pooling goldens
    x = torch.randn(len(p0), 24, generator=g)
    pool_t = torch.from_numpy(pool.astype(np.int64))
    out["pool_x"] = x.numpy()
    out["max_pool_y"] = rblocks.max_pool(x, pool_t).numpy()
    up = O.batch_query(p0, p1, l0, l1, radius=2.55, max_neighbors=40)
    xc = torch.randn(len(p1), 24, generator=g)
    out["up"] = up; out["closest_x"] = xc.numpy()
    out["closest_pool_y"] = rblocks.closest_pool(xc, torch.from_numpy(up.astype(np.int64))).numpy()
    assert torch.equal(blocks_ref.max_pool_ref(x, pool_t), torch.from_numpy(out["max_pool_y"]))
    np.savez_compressed(os.path.join(GOLD, "kpconv_small.npz"), **out)


def golden_encoder(O, rarch):
    cfg = kitti_config(first_feats_dim=16)
    a, b = small_pair(3, 1400)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p0, l0 = O.subsample_batch(raw, lens, sampleDl=0.3)
    limits = [30, 30, 30, 30]
    pyr = collate_ref(p0, l0, cfg, limits, O.subsample_batch, O.batch_query)
    batch = batch_to_torch(pyr)
    batch['features'] = torch.ones(len(p0), 1)
    torch.manual_seed(0); np.random.seed(0)
    model = rarch.KPFCNN(cfg).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items() if k.startswith('encoder_blocks.')}
    with torch.no_grad():
        x = batch['features'].clone()
        ref_outs = []
        for blk in model.encoder_blocks:
            x = blk(x, batch)
            ref_outs.append(x)
        mine = blocks_ref.encoder_ref(batch, sd, cfg, return_all=True)
    for i, (r, m) in enumerate(zip(ref_outs, mine)):
        err = ((r - m).norm() / r.norm()).item()
        assert err < 1e-5, f"encoder_ref block {i} != reference ({err})"
    print(f"  encoder: {len(ref_outs)} blocks, restatement max rel err "
          f"{max(((r - m).norm() / r.norm()).item() for r, m in zip(ref_outs, mine)):.2e}")
    out = {"p0": p0, "l0": l0, "limits": np.array(limits), "y_final": ref_outs[-1].numpy(),
           "block_norms": np.array([o.norm().item() for o in ref_outs]),
           "block0_y": ref_outs[0].numpy(), "block2_y": ref_outs[2].numpy()}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "encoder_small.npz"), **out)

    # also pin the restatement at the real KITTI width on a small cloud (not stored: 80 MB of weights)
    cfg = kitti_config(first_feats_dim=128)
    torch.manual_seed(0); np.random.seed(0)
    model = rarch.KPFCNN(cfg).eval()
    sd = {k: v.detach() for k, v in model.state_dict().items() if k.startswith('encoder_blocks.')}
    with torch.no_grad():
        x = batch['features'].clone()
        for blk in model.encoder_blocks:
            x = blk(x, batch)
        y = blocks_ref.encoder_ref(batch, sd, cfg)
    err = ((x - y).norm() / x.norm()).item()
    assert err < 1e-5, f"encoder_ref (first_feats_dim=128) != reference ({err})"
    print(f"  encoder first_feats_dim=128: restatement rel err {err:.2e}")


def golden_kpfcnn(O, rarch):
    """Full KPFCNN forward (encoder + bottleneck GNN + decoder) of the REAL reference on a small pair; pins
    blocks_ref.kpfcnn_ref and freezes the outputs (small widths so the fixture stays small)."""
    cfg = kitti_config(first_feats_dim=16, gnn_feats_dim=32, final_feats_dim=8)
    a, b = small_pair(5, 1500)
    raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
    p0, l0 = O.subsample_batch(raw, lens, sampleDl=0.3)
    limits = [30, 30, 30, 30]
    pyr = collate_ref(p0, l0, cfg, limits, O.subsample_batch, O.batch_query)
    batch = batch_to_torch(pyr)
    batch['features'] = torch.ones(len(p0), 1)
    torch.manual_seed(0); np.random.seed(0)
    model = rarch.KPFCNN(cfg).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ff, so, ss = model(batch)
        mf, mo, ms = blocks_ref.kpfcnn_ref(batch, sd, cfg)
    for name, r, m in (("feats_f", ff, mf), ("scores_overlap", so, mo), ("scores_saliency", ss, ms)):
        err = ((r - m).norm() / r.norm()).item()
        assert err < 1e-4, f"kpfcnn_ref {name} != reference ({err})"
        print(f"  kpfcnn {name}: restatement rel err {err:.2e}")
    out = {"p0": p0, "l0": l0, "limits": np.array(limits), "feats_f": ff.numpy(), "scores_overlap": so.numpy(),
           "scores_saliency": ss.numpy()}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "kpfcnn_small.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    O, R = Oracle(), RefL1()
    print("L1 goldens"); golden_l1(O, R)
    rblocks, rarch = import_reference()
    print("KPConv goldens"); golden_kpconv(O, rblocks)
    print("encoder goldens"); golden_encoder(O, rarch)
    print("KPFCNN goldens"); golden_kpfcnn(O, rarch)
    print("golden fixtures written to", GOLD)
