"""TEST INFRASTRUCTURE ONLY — numpy restatement of open3d's PointCloud.voxel_down_sample, the first-level voxelisation the
reference applies to raw scans (/root/reference/Predator_APR/datasets/kitti.py:468-471, :588-589).

PARITY UNPINNED: open3d (pinned open3d==0.10.0.0 in Predator_APR/requirements.txt:5) is a third-party dependency that is
neither vendored under /root/reference nor installed here, and the reference ships no fixture for this step. The algorithm
below restates the published source of that version (open3d/geometry/PointCloud.cpp, PointCloud::VoxelDownSample, and
AccumulatedPoint in the same file):
    voxel_min_bound = GetMinBound() - voxel_size * 0.5
    for i in range(n):  voxel_index = floor((points[i] - voxel_min_bound) / voxel_size)   (double arithmetic, per axis)
                        voxelindex_to_accpoint[voxel_index].AddPoint(i)                   (double sum, in point order)
    output = [acc.point_ / double(acc.num_of_points_) for acc in the unordered_map]       (iteration order: unspecified)
The reference widens the float32 scan to float64 first (make_open3d_point_cloud -> Vector3dVector) and narrows the result
back to float32 when the collate stacks it (datasets/dataloader.py:125, :163). Rows are returned in ascending
(iz, iy, ix) order — the canonical order the CUDA path emits.
Only tests/ may import this module."""
import numpy as np


def voxel_down_sample_ref(points_f32, voxel_size):
    p = np.asarray(points_f32, np.float32)[:, :3].astype(np.float64)
    if len(p) == 0:
        return np.zeros((0, 3), np.float32)
    vs = float(voxel_size)
    origin = p.min(0) - vs * 0.5
    ijk = np.floor((p - origin) / vs).astype(np.int64)
    order = np.lexsort((ijk[:, 0], ijk[:, 1], ijk[:, 2]))          # stable: ascending point index inside a voxel
    ijk_s, p_s = ijk[order], p[order]
    head = np.ones(len(p), bool)
    head[1:] = np.any(ijk_s[1:] != ijk_s[:-1], axis=1)
    starts = np.flatnonzero(head)
    ends = np.append(starts[1:], len(p))
    out = np.empty((len(starts), 3), np.float64)
    for r, (a, b) in enumerate(zip(starts, ends)):                  # sequential double sum in point order, like AddPoint
        acc = np.zeros(3)
        for row in p_s[a:b]:
            acc += row
        out[r] = acc / float(b - a)
    return out.astype(np.float32)


def voxel_down_sample_batch_ref(raw, lens, voxel_size):
    outs, olens, a = [], [], 0
    for n in lens:
        o = voxel_down_sample_ref(raw[a:a + int(n)], voxel_size)
        outs.append(o); olens.append(len(o)); a += int(n)
    return np.concatenate(outs) if outs else np.zeros((0, 3), np.float32), np.asarray(olens, np.int32)
