"""Drop-in for the CPython module `radius_neighbors`
(/root/reference/Predator_APR/cpp_wrappers/cpp_neighbors/wrapper.cpp:25-29): `batch_query` (:58-238).
Host arrays in, a fresh int32 numpy array [Nq, max_count] out; the search runs on the GPU through libaprb200.so
(aprb_radius_neighbors_batch). No CPU fallback.

Rows are ascending in fp32 d2 with ties broken by ascending support index (the order of the reference's own
batch_ordered_neighbors, neighbors.cpp:176-181; nanoflann leaves ties to an unstable sort). Pads = Ns.
`max_neighbors` is an extension (not in the reference signature): the [:, :max_neighbors] cut of
datasets/dataloader.py:66-70 done on the device, so only the kept columns cross PCIe.
"""
import numpy as np
import torch

from ... import ops as _ops

_GUESS = 128   # first-try width; the true width (max_count) is known after the pass


def _to_np(obj, dtype, what):
    try:
        if isinstance(obj, torch.Tensor):
            obj = obj.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(obj), dtype=dtype)
    except Exception:
        raise RuntimeError(f"Error converting {what} to numpy arrays of type "
                           f"{'float32' if dtype == np.float32 else 'int32'}") from None


def batch_query(queries, supports, q_batches, s_batches, *, radius=0.1, max_neighbors=0):
    q = _to_np(queries, np.float32, "query points")
    s = _to_np(supports, np.float32, "support points")
    ql = _to_np(q_batches, np.int32, "query batches")
    sl = _to_np(s_batches, np.int32, "support batches")
    if q.ndim != 2 or q.shape[1] != 3:                                   # wrapper.cpp:127-171
        raise RuntimeError("Wrong dimensions : query.shape is not (N, 3)")
    if s.ndim != 2 or s.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : support.shape is not (N, 3)")
    if ql.ndim > 1:
        raise RuntimeError("Wrong dimensions : queries_batches.shape is not (B,) ")
    if sl.ndim > 1:
        raise RuntimeError("Wrong dimensions : supports_batches.shape is not (B,) ")
    ql, sl = ql.reshape(-1), sl.reshape(-1)
    if ql.shape[0] != sl.shape[0]:
        raise RuntimeError("Wrong number of batch elements: different for queries and supports ")
    if int(ql.sum()) != q.shape[0] or int(sl.sum()) != s.shape[0] or ql.shape[0] < 1:
        raise RuntimeError("Wrong dimensions : batches do not sum to the number of points")
    if q.shape[0] == 0:
        raise RuntimeError("Error")                                      # wrapper.cpp:201-205
    dev = torch.device("cuda", torch.cuda.current_device())
    dq = torch.from_numpy(q).to(dev, non_blocking=True)
    ds = dq if supports is queries else torch.from_numpy(s).to(dev, non_blocking=True)
    dql = torch.from_numpy(ql).to(dev, non_blocking=True)
    dsl = dql if s_batches is q_batches else torch.from_numpy(sl).to(dev, non_blocking=True)
    r = float(np.float32(radius))                                        # "f" parse narrows to fp32 (wrapper.cpp:72-75)
    width = int(max_neighbors) if max_neighbors > 0 else _GUESS
    idx, _, maxc = _ops.radius_neighbors(dq, ds, dql, dsl, r, width, want_counts=True)
    max_count = int(maxc.item())
    if max_count < 1:
        raise RuntimeError("Error")                                      # no query has any neighbour
    if max_neighbors <= 0 and max_count > width:                         # rare: neighbourhoods wider than the guess
        idx = _ops.radius_neighbors(dq, ds, dql, dsl, r, max_count)
        width = max_count
    w = min(max_count, width)
    return _ops.to_host_numpy(idx[:, :w])
