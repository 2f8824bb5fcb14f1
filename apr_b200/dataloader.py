"""Python faces of the native neighbourhood ops and the pyramid builder — mirror of
/root/reference/Predator_APR/datasets/dataloader.py:15-232 (same function names, arguments and results).

Two flavours:
* host drop-ins `batch_grid_subsampling_kpconv`, `batch_neighbors_kpconv`, `collate_fn_descriptor`,
  `calibrate_neighbors` — numpy / CPU-torch in, CPU torch tensors out, like the reference (the GPU does the work,
  every call pays its own H2D/D2H);
* `build_pyramid_device` — the same schedule (dataloader.py:93-176) kept on the device: int32 index matrices of fixed
  width = neighbourhood limit go straight to the KPConv / pooling kernels, no int64 widening, no PCIe traffic.
"""
import numpy as np
import torch

from . import ops
from .cpp_wrappers.cpp_neighbors import radius_neighbors as cpp_neighbors
from .cpp_wrappers.cpp_subsampling import grid_subsampling as cpp_subsampling


def batch_grid_subsampling_kpconv(points, batches_len, features=None, labels=None, sampleDl=0.1, max_p=0, verbose=0,
                                  random_grid_orient=True):
    """dataloader.py:15-53"""
    if (features is None) and (labels is None):
        s_points, s_len = cpp_subsampling.subsample_batch(points, batches_len, sampleDl=sampleDl, max_p=max_p,
                                                          verbose=verbose)
        return torch.from_numpy(s_points), torch.from_numpy(s_len)
    elif labels is None:
        s_points, s_len, s_features = cpp_subsampling.subsample_batch(points, batches_len, features=features,
                                                                      sampleDl=sampleDl, max_p=max_p, verbose=verbose)
        return torch.from_numpy(s_points), torch.from_numpy(s_len), torch.from_numpy(s_features)
    elif features is None:
        s_points, s_len, s_labels = cpp_subsampling.subsample_batch(points, batches_len, classes=labels,
                                                                    sampleDl=sampleDl, max_p=max_p, verbose=verbose)
        return torch.from_numpy(s_points), torch.from_numpy(s_len), torch.from_numpy(s_labels)
    s_points, s_len, s_features, s_labels = cpp_subsampling.subsample_batch(
        points, batches_len, features=features, classes=labels, sampleDl=sampleDl, max_p=max_p, verbose=verbose)
    return (torch.from_numpy(s_points), torch.from_numpy(s_len), torch.from_numpy(s_features),
            torch.from_numpy(s_labels))


def batch_neighbors_kpconv(queries, supports, q_batches, s_batches, radius, max_neighbors):
    """dataloader.py:55-70 — the [:, :max_neighbors] cut happens on the device (same result: rows are sorted)."""
    neighbors = cpp_neighbors.batch_query(queries, supports, q_batches, s_batches, radius=radius,
                                          max_neighbors=max(int(max_neighbors), 0))
    return torch.from_numpy(neighbors)


def _pyramid_schedule(config):
    """Yields per layer (conv?, strided?, deformable-conv?, deformable-pool?) following dataloader.py:104-176."""
    arch = config.architecture
    layer_blocks = []
    for block_i, block in enumerate(arch):
        if 'global' in block or 'upsample' in block:
            break
        if not ('pool' in block or 'strided' in block):
            layer_blocks += [block]
            if block_i < len(arch) - 1 and not ('upsample' in arch[block_i + 1]):
                continue
        deform_conv = bool(layer_blocks) and any('deformable' in b for b in layer_blocks[:-1])
        yield bool(layer_blocks), ('pool' in block or 'strided' in block), deform_conv, 'deformable' in block
        layer_blocks = []


def _long(t):
    """`.long()` of dataloader.py:164-166, written straight into page-locked memory when a CUDA context exists: the int64 index
    matrices are 40 MB per KITTI pair and the caller ships them to the device next (trainer.py:299-305) — pageable, that copy
    ran at ~6 GB/s (tools/dropin_profile.py)."""
    if t.dtype == torch.int64 or t.numel() == 0 or not torch.cuda.is_available():
        return t.long()
    out = torch.empty(t.shape, dtype=torch.int64, pin_memory=True)
    out.copy_(t)
    return out


def collate_fn_descriptor(list_data, config, neighborhood_limits):
    """dataloader.py:72-198. list_data = [(src_pcd, tgt_pcd, src_feats, tgt_feats, rot, trans, matching_inds,
    src_pcd_raw, tgt_pcd_raw, src_nghb, tgt_nghb, sample)] with exactly one pair."""
    assert len(list_data) == 1
    (src_pcd, tgt_pcd, src_feats, tgt_feats, rot, trans, matching_inds, src_pcd_raw, tgt_pcd_raw, src_nghb, tgt_nghb,
     sample) = list_data[0]
    batched_features = torch.from_numpy(np.concatenate([src_feats, tgt_feats], axis=0))
    batched_points = torch.from_numpy(np.concatenate([src_pcd, tgt_pcd], axis=0))
    batched_lengths = torch.from_numpy(np.array([len(src_pcd), len(tgt_pcd)])).int()

    r_normal = config.first_subsampling_dl * config.conv_radius
    input_points, input_neighbors, input_pools, input_upsamples, input_batches_len = [], [], [], [], []
    for layer, (has_conv, strided, deform_conv, deform_pool) in enumerate(_pyramid_schedule(config)):
        if has_conv:
            r = r_normal * config.deform_radius / config.conv_radius if deform_conv else r_normal
            conv_i = batch_neighbors_kpconv(batched_points, batched_points, batched_lengths, batched_lengths, r,
                                            neighborhood_limits[layer])
        else:
            conv_i = torch.zeros((0, 1), dtype=torch.int64)
        if strided:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = batch_grid_subsampling_kpconv(batched_points, batched_lengths, sampleDl=dl)
            r = r_normal * config.deform_radius / config.conv_radius if deform_pool else r_normal
            pool_i = batch_neighbors_kpconv(pool_p, batched_points, pool_b, batched_lengths, r,
                                            neighborhood_limits[layer])
            up_i = batch_neighbors_kpconv(batched_points, pool_p, batched_lengths, pool_b, 2 * r,
                                          neighborhood_limits[layer])
        else:
            pool_i = torch.zeros((0, 1), dtype=torch.int64)
            pool_p = torch.zeros((0, 3), dtype=torch.float32)
            pool_b = torch.zeros((0,), dtype=torch.int64)
            up_i = torch.zeros((0, 1), dtype=torch.int64)
        input_points += [batched_points.float()]
        input_neighbors += [_long(conv_i)]
        input_pools += [_long(pool_i)]
        input_upsamples += [_long(up_i)]
        input_batches_len += [batched_lengths]
        batched_points, batched_lengths = pool_p, pool_b
        r_normal *= 2

    return {
        'points': input_points, 'neighbors': input_neighbors, 'pools': input_pools, 'upsamples': input_upsamples,
        'features': batched_features.float(), 'stack_lengths': input_batches_len,
        'rot': torch.from_numpy(rot), 'trans': torch.from_numpy(trans), 'correspondences': matching_inds,
        'src_pcd_raw': torch.from_numpy(src_pcd_raw).float(), 'tgt_pcd_raw': torch.from_numpy(tgt_pcd_raw).float(),
        'src_nghb': torch.from_numpy(src_nghb).float(), 'tgt_nghb': torch.from_numpy(tgt_nghb).float(),
        'sample': sample,
    }


def make_list_data(src_pcd, tgt_pcd):
    """A minimal dataset item in the reference's 12-tuple layout (features = ones, kitti.py:598-599)."""
    z3, e = np.zeros((1, 3), np.float32), np.zeros((0, 2), np.int64)
    return [(src_pcd, tgt_pcd, np.ones((len(src_pcd), 1), np.float32), np.ones((len(tgt_pcd), 1), np.float32),
             np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32), e, src_pcd, tgt_pcd, z3, z3, None)]


def calibrate_neighbors(dataset, config, collate_fn, keep_ratio=0.8, samples_threshold=2000):
    """dataloader.py:200-232 — 80th percentile of the per-layer neighbour-count histograms."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.deform_radius + 1) ** 3))
    neighb_hists = np.zeros((config.num_layers, hist_n), dtype=np.int32)
    for i in range(len(dataset)):
        batched_input = collate_fn([dataset[i]], config, neighborhood_limits=[hist_n] * 5)
        counts = [torch.sum(neighb_mat < neighb_mat.shape[0], dim=1).numpy() for neighb_mat in batched_input['neighbors']]
        hists = [np.bincount(c, minlength=hist_n)[:hist_n] for c in counts]
        neighb_hists += np.vstack(hists)
        if np.min(np.sum(neighb_hists, axis=1)) > samples_threshold:
            break
    cumsum = np.cumsum(neighb_hists.T, axis=0)
    percentiles = np.sum(cumsum < (keep_ratio * cumsum[hist_n - 1, :]), axis=0)
    return percentiles


# ----------------------------------------------------------------------------------------------------------------
# device-resident pyramid (no PCIe traffic; int32 indices of fixed width = limit)
# ----------------------------------------------------------------------------------------------------------------
def build_pyramid_device(points, lengths, config, neighborhood_limits, features=None):
    """points [N,3] f32 cuda, lengths [B] i32 cuda -> the reference's batch dict on the device.
    Index matrices are int32 [Nq, limit] (pad = Ns); the reference's width is min(max_count, limit), which differs only
    by extra all-pad columns when max_count < limit (see DESIGN.md for why that is result-neutral for KPConv)."""
    pts, lens = points.float().contiguous(), lengths.int().contiguous()
    r_normal = config.first_subsampling_dl * config.conv_radius
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[], pool_widths=[])
    empty_i = torch.zeros((0, 1), dtype=torch.int32, device=pts.device)
    sched = list(_pyramid_schedule(config))
    # One cell list per level, shared by every search whose supports are that level's points: conv(l), pool(l) and the
    # upsample search of level l-1 (radius 2 r_{l-1} = r_l). Built for the largest radius asked of it.
    grid = None
    for layer, (has_conv, strided, deform_conv, deform_pool) in enumerate(sched):
        lim = int(neighborhood_limits[layer])
        r_conv = r_normal * config.deform_radius / config.conv_radius if deform_conv else r_normal
        r_pool = r_normal * config.deform_radius / config.conv_radius if deform_pool else r_normal
        need = max(r_conv if has_conv else 0.0, r_pool if strided else 0.0)
        if grid is None or grid.radius < need:
            grid = ops.CellGrid(pts, lens, need) if need > 0 else None
        conv_i = grid.query(pts, lens, r_conv, lim) if has_conv else empty_i
        if strided:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = ops.grid_subsample(pts, lens, dl)
            pool_i, _, pool_w = grid.query(pool_p, pool_b, r_pool, lim, want_counts=True)
            # grid over the next level's points: serves this level's upsample search and the next level's conv/pool
            nxt = sched[layer + 1] if layer + 1 < len(sched) else None
            r_next = 2 * r_normal
            need_next = 2 * r_pool
            if nxt is not None:
                if nxt[0]:
                    need_next = max(need_next, r_next * config.deform_radius / config.conv_radius if nxt[2] else r_next)
                if nxt[1]:
                    need_next = max(need_next, r_next * config.deform_radius / config.conv_radius if nxt[3] else r_next)
            next_grid = ops.CellGrid(pool_p, pool_b, need_next)
            up_i = next_grid.query(pts, lens, 2 * r_pool, lim)
        else:
            pool_i, up_i, pool_w, next_grid = empty_i, empty_i, None, None
            pool_p = torch.zeros((0, 3), dtype=torch.float32, device=pts.device)
            pool_b = torch.zeros((0,), dtype=torch.int32, device=pts.device)
        out['points'].append(pts); out['neighbors'].append(conv_i); out['pools'].append(pool_i)
        out['upsamples'].append(up_i); out['stack_lengths'].append(lens); out['pool_widths'].append(pool_w)
        pts, lens, grid = pool_p, pool_b, next_grid
        r_normal *= 2
    out['features'] = features if features is not None else torch.ones((out['points'][0].shape[0], 1),
                                                                     dtype=torch.float32, device=points.device)
    return out


def calibrate_neighbors_device(pairs, config, keep_ratio=0.8, samples_threshold=2000):
    """dataloader.py:200-232 on the device: pairs = iterable of (points [N,3] cuda, lengths [2] cuda).
    Only the untruncated per-query counts of the conv searches are needed, so the searches run at width 1."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.deform_radius + 1) ** 3))
    hists = np.zeros((config.num_layers, hist_n), dtype=np.int64)
    for pts, lens in pairs:
        r_normal = config.first_subsampling_dl * config.conv_radius
        pts, lens = pts.float().contiguous(), lens.int().contiguous()
        for layer, (has_conv, strided, deform_conv, _) in enumerate(_pyramid_schedule(config)):
            if has_conv:
                r = r_normal * config.deform_radius / config.conv_radius if deform_conv else r_normal
                _, counts, _ = ops.radius_neighbors(pts, pts, lens, lens, r, 1, want_counts=True)
                c = torch.clamp(counts, max=hist_n).cpu().numpy()   # rows wider than hist_n are cut to hist_n columns
                hists[layer] += np.bincount(c, minlength=hist_n + 1)[:hist_n]   # same cut as np.bincount(...)[:hist_n]
            if strided:
                pts, lens = ops.grid_subsample(pts, lens, 2 * r_normal / config.conv_radius)
            r_normal *= 2
        if np.min(np.sum(hists, axis=1)) > samples_threshold:
            break
    cumsum = np.cumsum(hists.T, axis=0)
    return np.sum(cumsum < (keep_ratio * cumsum[hist_n - 1, :]), axis=0)
