"""Drop-in for the CPython module `grid_subsampling`
(/root/reference/Predator_APR/cpp_wrappers/cpp_subsampling/wrapper.cpp:28-33): `subsample` (:338-566) and
`subsample_batch` (:62-333). Host arrays in (numpy / CPU torch, any float dtype), fresh numpy arrays out; the work
runs on the GPU through libaprb200.so (aprb_grid_subsample_batch). No CPU fallback.

Same keyword-only options, defaults, return tuples and RuntimeError strings as the reference wrapper. Documented
deviations: rows come out in ascending voxel order per cloud (the reference's order is unordered_map iteration order);
the `classes` vote (grid_subsampling.cpp:96-101) breaks ties by the smallest label (the reference: map iteration order).
"""
import numpy as np
import torch

from ... import ops as _ops

_METHODS = ("barycenters", "voxelcenters")


def _to_np(obj, dtype, what):
    try:
        if isinstance(obj, torch.Tensor):
            obj = obj.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(obj), dtype=dtype)
    except Exception:
        raise RuntimeError(f"Error converting input {what} to numpy arrays of type "
                           f"{'float32' if dtype == np.float32 else 'int32'}") from None


def _run(points, batches, features, classes, sampleDl, method, max_p):
    if method not in _METHODS:                                           # wrapper.cpp:92-96 (validated, never used)
        raise RuntimeError('Error parsing method. Valid method names are "barycenters" and "voxelcenters" ')
    p = _to_np(points, np.float32, "points")
    b = _to_np(batches, np.int32, "batches")
    f = _to_np(features, np.float32, "features") if features is not None else None
    c = _to_np(classes, np.int32, "classes") if classes is not None else None
    if p.ndim != 2 or p.shape[1] != 3:                                   # wrapper.cpp:154-162
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    if b.ndim > 1:
        raise RuntimeError("Wrong dimensions : batches.shape is not (B,) ")
    if f is not None and (f.ndim != 2 or f.shape[0] != p.shape[0]):
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if c is not None and (c.ndim > 2 or c.shape[0] != p.shape[0]):       # wrapper.cpp:178-190
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    b = b.reshape(-1)
    if int(b.sum()) != p.shape[0] or (b < 0).any():
        raise RuntimeError("Wrong dimensions : batches do not sum to the number of points")
    if p.shape[0] == 0:
        raise RuntimeError("Error")                                      # wrapper.cpp:267-271: empty result
    dev = torch.device("cuda", torch.cuda.current_device())
    dp = torch.from_numpy(p).to(dev, non_blocking=True)
    db = torch.from_numpy(b).to(dev, non_blocking=True)
    df = torch.from_numpy(f).to(dev, non_blocking=True) if f is not None else None
    dc = torch.from_numpy(c).to(dev, non_blocking=True) if c is not None else None
    res = _ops.grid_subsample(dp, db, float(np.float32(sampleDl)), int(max_p), features=df, classes=dc)
    if res[0].shape[0] < 1:
        raise RuntimeError("Error")
    return tuple(_ops.to_host_numpy(t) for t in res)                     # classes come back [M, ldim]: wrapper.cpp:282-284


def subsample_batch(points, batches, *, features=None, classes=None, sampleDl=0.1, method="barycenters", max_p=0,
                    verbose=0):
    """(points f32 [M,3], lengths i32 [B][, features f32 [M,d]][, classes i32 [M,d]]) — wrapper.cpp:315-322."""
    return _run(points, batches, features, classes, sampleDl, method, max_p)


def subsample(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0):
    """Single cloud: points f32 [M,3] (or a tuple with features) — wrapper.cpp:338-566."""
    p = _to_np(points, np.float32, "points")
    res = _run(p, np.array([p.shape[0] if p.ndim == 2 else 0], np.int32), features, classes, sampleDl, method, 0)
    out = (res[0],) + tuple(res[2:])                                     # drop the lengths: wrapper.cpp:548-555
    return out[0] if len(out) == 1 else out
