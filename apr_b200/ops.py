"""Device-resident operators: CUDA tensors in, CUDA tensors out, all work done by libaprb200.so.

These are the stream-ordered building blocks behind the reference-facing modules
(`apr_b200.cpp_wrappers.*`, `apr_b200.blocks`). torch only allocates memory and provides the current stream.
"""
import ctypes as C

import torch

from . import _native as N

_ws_cache = {}
TRACE = None   # when a list: every op appends its shape record (bench.py derives algorithmic bytes/flops from it)


def _workspace(nbytes, device):
    """Grow-only scratch buffer per (device, stream). Calls on one stream are ordered, so reuse is safe."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _dev_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise N.NativeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


def _dev_i32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise N.NativeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    return t.contiguous().int() if (t.dtype != torch.int32 or not t.is_contiguous()) else t


def _idx(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise N.NativeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype not in (torch.int32, torch.int64):
        t = t.long()
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t, (1 if t.dtype == torch.int64 else 0), (t.stride(0) if t.dim() == 2 and t.shape[0] > 1 else t.shape[-1])


# --------------------------------------------------------------------------------------------------------------
def to_host_numpy(t):
    """Device tensor -> a fresh numpy array the caller owns, copied by DMA straight into page-locked memory (torch's caching
    host allocator) instead of through a pageable bounce: the drop-in wrappers return ~30 MB of index matrices per pair and
    the pageable route ran at ~5 GB/s (tools/dropin_profile.py)."""
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out.numpy()


def grid_subsample(points, lens, dl, max_p=0, features=None, sync=True, key_bits=32, classes=None):
    """K1. points [N,3] f32 cuda, lens [B] i32 cuda. Returns (sub_points [M,3], sub_lens [B] i32[, sub_feats][, sub_classes]).
    classes [N] or [N,ldim] i32: per-voxel label vote (grid_subsampling.cpp:96-101; ties -> smallest label).
    With sync=False returns the full-capacity buffers plus the device scalars [M, status] (no host round trip).
    The 32-bit sort key is tried first; a grid that needs more bits is re-run with the 64-bit key."""
    N.require_cuda()
    points, lens = _dev_f32(points, "points"), _dev_i32(lens, "lens")
    n, b = points.shape[0], lens.shape[0]
    dev = points.device
    out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
    out_lens = torch.empty(b, dtype=torch.int32, device=dev)
    m_dev = torch.zeros(2, dtype=torch.int32, device=dev)      # [M, status]
    fdim, feats, out_f = 0, None, None
    if features is not None:
        feats = _dev_f32(features, "features")
        fdim = feats.shape[1]
        out_f = torch.empty((max(n, 1), fdim), dtype=torch.float32, device=dev)
    ldim, cls, out_c = 0, None, None
    if classes is not None:
        cls = _dev_i32(classes.reshape(classes.shape[0], -1), "classes")
        if cls.shape[0] != n:
            raise ValueError("classes must have one row per point")
        ldim = cls.shape[1]
        out_c = torch.empty((max(n, 1), ldim), dtype=torch.int32, device=dev)
    nbytes = N.lib().aprb_grid_subsample_ws_bytes(n, b, fdim)
    ws = _workspace(nbytes, dev)
    rc = N.lib().aprb_grid_subsample_batch_labels(N.ptr(points), N.ptr(lens), b, n, float(dl), int(max_p), N.ptr(feats), fdim,
                                                  N.ptr(cls), ldim, N.ptr(out), N.ptr(out_lens), N.ptr(m_dev[0:1]),
                                                  N.ptr(out_f), N.ptr(out_c), N.ptr(m_dev[1:2]), int(key_bits), N.ptr(ws),
                                                  ws.numel(), N.stream_ptr())
    N.check(rc, "aprb_grid_subsample_batch_labels")
    extra = tuple(t for t in (out_f, out_c) if t is not None)
    if not sync:
        return (out, out_lens, m_dev) + extra
    m, status = m_dev.tolist()
    if status == 2 and key_bits == 32:
        return grid_subsample(points, lens, dl, max_p, features, sync, key_bits=64, classes=classes)
    if TRACE is not None:
        TRACE.append(("sub", n, m))
    if status != 0:
        raise N.NativeError("grid_subsample: voxel grid does not fit the 64-bit sort key (cloud extent / sampleDl too large)")
    return (out[:m], out_lens) + tuple(t[:m] for t in extra)


def voxel_downsample_raw(raw, lens, voxel_size, sync=True, key_bits=32):
    """First-level voxelisation of raw scans with open3d's voxel_down_sample semantics (kitti.py:468-471). raw [N, >=3]
    f32 cuda rows starting with x, y, z (KITTI: [N,4] xyzr), lens [B] i32 cuda -> (points [M,3] f32, lens [B] i32);
    with sync=False the full-capacity buffers and the device scalars [M, status]."""
    N.require_cuda()
    if not (isinstance(raw, torch.Tensor) and raw.is_cuda and raw.dtype == torch.float32 and raw.dim() == 2 and raw.shape[1] >= 3):
        raise N.NativeError("voxel_downsample_raw: expected a CUDA float32 tensor [N, >= 3]")
    rr, ll = raw.contiguous(), _dev_i32(lens, "lens")
    n, stride, b = rr.shape[0], rr.shape[1], ll.shape[0]
    out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=rr.device)
    out_lens = torch.empty(b, dtype=torch.int32, device=rr.device)
    m_dev = torch.zeros(2, dtype=torch.int32, device=rr.device)
    ws = _workspace(N.lib().aprb_voxel_downsample_ws_bytes(n, b), rr.device)
    rc = N.lib().aprb_voxel_downsample_raw(N.ptr(rr), stride, N.ptr(ll), b, n, float(voxel_size), N.ptr(out), N.ptr(out_lens),
                                           N.ptr(m_dev[0:1]), N.ptr(m_dev[1:2]), int(key_bits), N.ptr(ws), ws.numel(), N.stream_ptr())
    N.check(rc, "aprb_voxel_downsample_raw")
    if not sync:
        return out, out_lens, m_dev
    m, status = m_dev.tolist()
    if status == 2 and key_bits == 32:
        return voxel_downsample_raw(raw, lens, voxel_size, sync, key_bits=64)
    if status != 0:
        raise N.NativeError("voxel_downsample_raw: voxel grid does not fit the 64-bit sort key")
    return out[:m], out_lens


def radius_neighbors(queries, supports, q_lens, s_lens, radius, width, want_counts=False):
    """K2+K3. Returns idx [Nq,width] i32 (pad = Ns), and optionally (counts [Nq] i32, max_count [1] i32) on device."""
    N.require_cuda()
    q, s = _dev_f32(queries, "queries"), _dev_f32(supports, "supports")
    ql, sl = _dev_i32(q_lens, "q_lens"), _dev_i32(s_lens, "s_lens")
    nq, ns, b = q.shape[0], s.shape[0], ql.shape[0]
    dev = q.device
    width = int(width)
    out = torch.empty((nq, width), dtype=torch.int32, device=dev)
    counts = torch.empty(max(nq, 1), dtype=torch.int32, device=dev) if want_counts else None
    maxc = torch.zeros(1, dtype=torch.int32, device=dev) if want_counts else None
    nbytes = N.lib().aprb_radius_neighbors_ws_bytes(nq, ns, b)
    ws = _workspace(nbytes, dev)
    rc = N.lib().aprb_radius_neighbors_batch(N.ptr(q), N.ptr(s), N.ptr(ql), N.ptr(sl), b, nq, ns, float(radius), width,
                                             N.ptr(out), width, N.ptr(counts), N.ptr(maxc), N.ptr(ws), ws.numel(),
                                             N.stream_ptr())
    N.check(rc, "aprb_radius_neighbors_batch")
    if TRACE is not None:
        TRACE.append(("nb", nq, ns, width))
    if want_counts:
        return out, counts[:nq], maxc
    return out


class CellGrid:
    """Cell list over a stacked support set (K2), reusable by several radius queries with radius <= the build radius."""

    def __init__(self, supports, s_lens, radius):
        N.require_cuda()
        self.s = _dev_f32(supports, "supports")
        self.sl = _dev_i32(s_lens, "s_lens")
        self.ns, self.b, self.radius = self.s.shape[0], self.sl.shape[0], float(radius)
        nbytes = N.lib().aprb_cell_grid_bytes(self.ns, self.b)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.s.device)
        N.check(N.lib().aprb_cell_grid_build(N.ptr(self.s), N.ptr(self.sl), self.b, self.ns, self.radius, N.ptr(self.buf),
                                             self.buf.numel(), N.stream_ptr()), "aprb_cell_grid_build")

    def query(self, queries, q_lens, radius, width, want_counts=False):
        """Same result as radius_neighbors(queries, supports, q_lens, s_lens, radius, width)."""
        if float(radius) > self.radius * (1 + 1e-6):
            raise N.NativeError("CellGrid.query: radius larger than the radius the grid was built for")
        q, ql = _dev_f32(queries, "queries"), _dev_i32(q_lens, "q_lens")
        nq, width = q.shape[0], int(width)
        out = torch.empty((nq, width), dtype=torch.int32, device=q.device)
        counts = torch.empty(max(nq, 1), dtype=torch.int32, device=q.device) if want_counts else None
        maxc = torch.zeros(1, dtype=torch.int32, device=q.device) if want_counts else None
        rc = N.lib().aprb_cell_grid_query(N.ptr(self.buf), self.buf.numel(), N.ptr(q), N.ptr(ql), self.b, nq, self.ns,
                                          float(radius), width, N.ptr(out), width, N.ptr(counts), N.ptr(maxc),
                                          N.stream_ptr())
        N.check(rc, "aprb_cell_grid_query")
        if TRACE is not None:
            TRACE.append(("nb", nq, self.ns, width))
        return (out, counts[:nq], maxc) if want_counts else out

    def query_nearest(self, queries, q_lens, radius):
        """[Nq, 1] int32: column 0 of query(...)'s matrix — the nearest support inside the ball (pad = Ns when it is empty),
        all the reference reads of an upsample matrix (closest_pool, blocks.py:71-83)."""
        if float(radius) > self.radius * (1 + 1e-6):
            raise N.NativeError("CellGrid.query_nearest: radius larger than the radius the grid was built for")
        q, ql = _dev_f32(queries, "queries"), _dev_i32(q_lens, "q_lens")
        nq = q.shape[0]
        out = torch.empty((nq, 1), dtype=torch.int32, device=q.device)
        rc = N.lib().aprb_cell_grid_query_nearest(N.ptr(self.buf), self.buf.numel(), N.ptr(q), N.ptr(ql), self.b, nq, self.ns,
                                                  float(radius), N.ptr(out), 1, N.stream_ptr())
        N.check(rc, "aprb_cell_grid_query_nearest")
        return out


def kpconv_prepare_weights(weights):
    """[K,Cin,Cout] f32 -> prepared TF32 K-major operand [Cout, K*Cin]."""
    N.require_cuda()
    w = _dev_f32(weights.detach(), "weights")
    k, cin, cout = w.shape
    out = torch.empty((cout, k * cin), dtype=torch.float32, device=w.device)
    N.check(N.lib().aprb_kpconv_prepare_weights(N.ptr(w), k, cin, cout, N.ptr(out), N.stream_ptr()),
            "aprb_kpconv_prepare_weights")
    return out


# Hand the GEMM epilogue's 32-row group statistics of a produced tensor to the InstanceNorm that consumes it (attribute
# `_aprb_gstat` on the returned tensor; instnorm_lrelu_seg picks it up) — the module-path twin of what aprb_kfe_forward
# does natively, so both paths run the same kernels in the same order.
FUSE_STATS = True


def _gstat_of(t):
    rec = getattr(t, "_aprb_gstat", None)
    if rec is None or rec[1] != t._version:
        return None
    return rec[0]


def group_stats(t):
    """The GEMM-epilogue group statistics attached to a produced tensor ([groups, 2, C] floats, flat), or None."""
    return _gstat_of(t)


def _group_stats_buffer(n, c, device):
    nbytes = N.lib().aprb_group_stats_bytes(int(n), int(c))
    return torch.empty(nbytes // 4, dtype=torch.float32, device=device)


def kpconv_prepare_weights_f16(weights):
    """[K,Cin,Cout] f32 -> prepared fp16 K-major operand [Cout, K*Cin] (kpconv mode 3)."""
    N.require_cuda()
    w = _dev_f32(weights.detach(), "weights")
    k, cin, cout = w.shape
    out = torch.empty((cout, k * cin), dtype=torch.float16, device=w.device)
    N.check(N.lib().aprb_kpconv_prepare_weights_f16(N.ptr(w), k, cin, cout, N.ptr(out), N.stream_ptr()),
            "aprb_kpconv_prepare_weights_f16")
    return out


def kpconv_prepare_weights_f16_ck(weights):
    """[K,Cin,Cout] f32 -> fp16 operand [Cout, Cin*16], column c*16 + k (kpconv mode 5: tcgen05 weighting kernel)."""
    N.require_cuda()
    w = _dev_f32(weights.detach(), "weights")
    k, cin, cout = w.shape
    out = torch.empty((cout, cin * 16), dtype=torch.float16, device=w.device)
    N.check(N.lib().aprb_kpconv_prepare_weights_f16_ck(N.ptr(w), k, cin, cout, N.ptr(out), N.stream_ptr()),
            "aprb_kpconv_prepare_weights_f16_ck")
    return out


def kpconv_tc_supported(h, k, cin, cout, ns=1):
    return bool(N.lib().aprb_kpconv_tc_supported(int(h), int(k), int(cin), int(cout), int(ns)))


def kpconv_f16_supported(k, cin, cout, h):
    return (k * cin) % 64 == 0 and cout % 16 == 0 and cin % 4 == 0 and h <= 128


def kpconv(q_pts, s_pts, neighb_inds, x, kernel_points, weights, extent, wprep=None, mode=0):
    """K5. Returns [Nq,Cout] f32. mode 4 takes x in fp16 (the native pipeline's activation storage)."""
    N.require_cuda()
    q, s = _dev_f32(q_pts, "q_pts"), _dev_f32(s_pts, "s_pts")
    if mode in (4, 5):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float16):
            raise N.NativeError("kpconv mode 4/5: x must be a CUDA float16 tensor")
        xx = x.contiguous()
    else:
        xx = _dev_f32(x, "x")
    kp, w = _dev_f32(kernel_points.detach(), "kernel_points"), _dev_f32(weights.detach(), "weights")
    idx, is64, ld = _idx(neighb_inds, "neighb_inds")
    nq, ns, h = q.shape[0], s.shape[0], idx.shape[1]
    k, cin, cout = w.shape
    if xx.shape[0] != ns or xx.shape[1] != cin:
        raise N.NativeError(f"kpconv: x has shape {tuple(xx.shape)}, expected ({ns}, {cin})")
    out = torch.empty((nq, cout), dtype=torch.float32, device=q.device)
    nbytes = N.lib().aprb_kpconv_ws_bytes(nq, ns, h, k, cin, cout)
    ws = _workspace(nbytes, q.device)
    gs = _group_stats_buffer(nq, cout, q.device) if (FUSE_STATS and wprep is not None and mode != 1 and nq > 0) else None
    written = C.c_int(0)
    rc = N.lib().aprb_kpconv_forward_stats(N.ptr(q), N.ptr(s), N.ptr(idx), is64, ld, N.ptr(xx), N.ptr(kp), N.ptr(w),
                                           N.ptr(wprep), float(extent), nq, ns, h, k, cin, cout, N.ptr(out), int(mode),
                                           N.ptr(gs), C.byref(written) if gs is not None else None,
                                           N.ptr(ws), ws.numel(), N.stream_ptr())
    N.check(rc, "aprb_kpconv_forward")
    if written.value:
        out._aprb_gstat = (gs, out._version)
    if TRACE is not None:
        TRACE.append(("kpconv", nq, ns, h, k, cin, cout))
    return out


def max_pool(x, inds, width_dev=None, seg_off=None):
    """K4. x [Ns,C] f32 (or f16: C % 128 == 0, int32 inds), inds [Nq,H] -> [Nq,C]; the shadow index Ns contributes an
    all-zero row. width_dev (device i32): columns that take part — one value, or one per segment with seg_off [S+1]
    (query-row offsets of the collated pairs of a super-batch)."""
    N.require_cuda()
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float16:
        xx, f16 = x.contiguous(), 1
    else:
        xx, f16 = _dev_f32(x, "x"), 0
    idx, is64, ld = _idx(inds, "inds")
    nq, h = idx.shape
    ns, c = xx.shape
    out = torch.empty((nq, c), dtype=xx.dtype, device=xx.device)
    if f16 or seg_off is not None:
        wd = _dev_i32(width_dev, "width_dev") if width_dev is not None else None
        so = _dev_i32(seg_off, "seg_off") if seg_off is not None else None
        nseg = so.shape[0] - 1 if so is not None else 1
        rc = N.lib().aprb_max_pool_seg(N.ptr(xx), f16, N.ptr(idx), is64, ld, nq, ns, h, c, N.ptr(so), nseg, N.ptr(wd),
                                       N.ptr(out), N.stream_ptr())
    else:
        rc = N.lib().aprb_max_pool(N.ptr(xx), N.ptr(idx), is64, ld, nq, ns, h, c, N.ptr(width_dev), N.ptr(out), N.stream_ptr())
    N.check(rc, "aprb_max_pool")
    return out


def pool_seg_widths(inds, ns, seg_off=None):
    """Per-segment width of the reference's pool matrix recovered from a fixed-width index matrix: 1 + last valid column
    over the segment's rows (device i32 [S])."""
    N.require_cuda()
    idx, is64, ld = _idx(inds, "inds")
    nq, h = idx.shape
    so = _dev_i32(seg_off, "seg_off") if seg_off is not None else None
    nseg = so.shape[0] - 1 if so is not None else 1
    out = torch.empty(nseg, dtype=torch.int32, device=idx.device)
    N.check(N.lib().aprb_pool_seg_widths(N.ptr(idx), is64, ld, nq, int(ns), h, N.ptr(so), nseg, N.ptr(out), N.stream_ptr()),
            "aprb_pool_seg_widths")
    return out


def closest_pool(x, inds):
    """K4. out[n] = (x ++ 0)[inds[n,0]]."""
    N.require_cuda()
    xx = _dev_f32(x, "x")
    if inds.dim() == 1:
        inds = inds.unsqueeze(1)
    idx, is64, ld = _idx(inds, "inds")
    nq = idx.shape[0]
    ns, c = xx.shape
    out = torch.empty((nq, c), dtype=torch.float32, device=xx.device)
    rc = N.lib().aprb_closest_pool(N.ptr(xx), N.ptr(idx), is64, ld, nq, ns, c, N.ptr(out), N.stream_ptr())
    N.check(rc, "aprb_closest_pool")
    return out


def instnorm_lrelu(x, slope=0.1, residual=None, norm_residual=False, eps=1e-5, out=None, round_tf32=False):
    """K6. y = act(standardise_cols(x) [+ residual | + standardise_cols(residual)]); slope=1.0 disables the activation."""
    N.require_cuda()
    xx = _dev_f32(x, "x")
    n, c = xx.shape
    res = _dev_f32(residual, "residual") if residual is not None else None
    y = out if out is not None else torch.empty_like(xx)
    ws = _workspace(N.lib().aprb_instnorm_ws_bytes(n, c), xx.device)
    rc = N.lib().aprb_instnorm_lrelu(N.ptr(xx), n, c, float(eps), float(slope), N.ptr(res), 1 if norm_residual else 0,
                                     1 if round_tf32 else 0, N.ptr(y), N.ptr(ws), ws.numel(), N.stream_ptr())
    N.check(rc, "aprb_instnorm_lrelu")
    return y


def segment_offsets(lens, clouds_per_segment):
    """Row offsets [S+1] i32 of the normalisation segments (one per collated pair) of a stacked batch."""
    N.require_cuda()
    ll = _dev_i32(lens, "lens")
    b = ll.shape[0]
    s = (b + clouds_per_segment - 1) // clouds_per_segment
    out = torch.empty(s + 1, dtype=torch.int32, device=ll.device)
    N.check(N.lib().aprb_segment_offsets(N.ptr(ll), b, int(clouds_per_segment), N.ptr(out), N.stream_ptr()),
            "aprb_segment_offsets")
    return out


def instnorm_lrelu_seg(x, seg_off=None, slope=0.1, residual=None, norm_residual=False, eps=1e-5, out=None, round_tf32=False):
    """K6, segmented: rows [seg_off[s], seg_off[s+1]) are standardised with their own column statistics (one segment per
    collated pair of a super-batch). seg_off=None is one segment over all rows."""
    N.require_cuda()
    xx = _dev_f32(x, "x")
    n, c = xx.shape
    res = _dev_f32(residual, "residual") if residual is not None else None
    so = _dev_i32(seg_off, "seg_off") if seg_off is not None else None
    nseg = so.shape[0] - 1 if so is not None else 1
    y = out if out is not None else torch.empty_like(xx)
    ws = _workspace(N.lib().aprb_instnorm_seg_ws_bytes(n, c, nseg), xx.device)
    # group statistics left on the tensors by the GEMM that produced them (only valid for that very tensor object)
    # (and only while it has not been modified in place since: the producer records the tensor's version counter)
    gx = _gstat_of(x) if xx is x else None
    gr = _gstat_of(residual) if (res is not None and res is residual and norm_residual) else None
    rc = N.lib().aprb_instnorm_lrelu_seg_pre(N.ptr(xx), n, c, N.ptr(so), nseg, float(eps), float(slope), N.ptr(res),
                                             1 if norm_residual else 0, 1 if round_tf32 else 0, N.ptr(y), N.ptr(gx), N.ptr(gr),
                                             N.ptr(ws), ws.numel(), N.stream_ptr())
    N.check(rc, "aprb_instnorm_lrelu_seg")
    return y


def linear_tf32_supported(n, cin, cout):
    return cin % 32 == 0 and cout % 16 == 0 and n > 0


def linear_tf32(x, weight):
    """y = x @ weight.T on tcgen05 (TF32 operands, fp32 accumulate). weight is nn.Linear's [Cout,Cin]."""
    N.require_cuda()
    xx, w = _dev_f32(x, "x"), _dev_f32(weight.detach(), "weight")
    n, cin = xx.shape
    cout = w.shape[0]
    y = torch.empty((n, cout), dtype=torch.float32, device=xx.device)
    ws = _workspace(N.lib().aprb_linear_tf32_ws_bytes(n, cin, cout), xx.device)
    gs = _group_stats_buffer(n, cout, xx.device) if (FUSE_STATS and n > 0) else None
    written = C.c_int(0)
    N.check(N.lib().aprb_linear_tf32_stats(N.ptr(xx), N.ptr(w), n, cin, cout, N.ptr(y), N.ptr(gs),
                                           C.byref(written) if gs is not None else None, N.ptr(ws), ws.numel(), N.stream_ptr()),
            "aprb_linear_tf32")
    if written.value:
        y._aprb_gstat = (gs, y._version)
    if TRACE is not None:
        TRACE.append(("linear", n, cin, cout))
    return y


def round_tf32(t):
    """Copy of t rounded to TF32 (nearest)."""
    N.require_cuda()
    tt = _dev_f32(t.detach(), "t")
    out = torch.empty_like(tt)
    N.check(N.lib().aprb_round_tf32(N.ptr(tt), N.ptr(out), tt.numel(), N.stream_ptr()), "aprb_round_tf32")
    return out


# ---- training path (stage A+B alone, data gradients) ------------------------------------------------------------------
def kpconv_weighted(q_pts, s_pts, neighb_inds, x, kernel_points, extent, round_tf32=False):
    """Stage A+B of KPConv: (wf [Nq, K*Cin], inv_nn [Nq]) — see aprb_kpconv_weighted."""
    N.require_cuda()
    q, s, xx = _dev_f32(q_pts, "q_pts"), _dev_f32(s_pts, "s_pts"), _dev_f32(x, "x")
    kp = _dev_f32(kernel_points.detach(), "kernel_points")
    idx, is64, ld = _idx(neighb_inds, "neighb_inds")
    nq, ns, h, k, cin = q.shape[0], s.shape[0], idx.shape[1], kp.shape[0], xx.shape[1]
    wf = torch.empty((nq, k * cin), dtype=torch.float32, device=q.device)
    inv_nn = torch.empty(nq, dtype=torch.float32, device=q.device)
    ws = _workspace(N.lib().aprb_kpconv_weighted_ws_bytes(ns), q.device)
    rc = N.lib().aprb_kpconv_weighted(N.ptr(q), N.ptr(s), N.ptr(idx), is64, ld, N.ptr(xx), N.ptr(kp), float(extent), nq, ns, h,
                                      k, cin, 1 if round_tf32 else 0, N.ptr(wf), N.ptr(inv_nn), N.ptr(ws), ws.numel(),
                                      N.stream_ptr())
    N.check(rc, "aprb_kpconv_weighted")
    return wf, inv_nn


def kpconv_weighted_f16(q_pts, s_pts, neighb_inds, x16, kernel_points, extent, layout_ck=False):
    """Stage A+B in the native pipeline's format: fp16 features -> (wf fp16, inv_nn [Nq] f32). wf is [Nq, K*Cin]
    (kernel-point-major, CUDA-core list kernel) or, with layout_ck, [Nq, Cin*16] (tcgen05 weighting kernel)."""
    N.require_cuda()
    q, s = _dev_f32(q_pts, "q_pts"), _dev_f32(s_pts, "s_pts")
    kp = _dev_f32(kernel_points.detach(), "kernel_points")
    idx = _dev_i32(neighb_inds, "neighb_inds")
    if not (x16.is_cuda and x16.dtype == torch.float16):
        raise N.NativeError("kpconv_weighted_f16: x must be a CUDA float16 tensor")
    xx = x16.contiguous()
    nq, ns, h, k, cin = q.shape[0], s.shape[0], idx.shape[1], kp.shape[0], xx.shape[1]
    wf = torch.empty((nq, (16 if layout_ck else k) * cin), dtype=torch.float16, device=q.device)
    inv_nn = torch.empty(nq, dtype=torch.float32, device=q.device)
    ws = _workspace(N.lib().aprb_kpconv_weighted_ws_bytes(ns), q.device)
    rc = N.lib().aprb_kpconv_weighted_f16(N.ptr(q), N.ptr(s), N.ptr(idx), idx.shape[1], N.ptr(xx), N.ptr(kp), float(extent), nq, ns,
                                          h, k, cin, 1 if layout_ck else 0, N.ptr(wf), N.ptr(inv_nn), N.ptr(ws), ws.numel(),
                                          N.stream_ptr())
    N.check(rc, "aprb_kpconv_weighted_f16")
    return wf, inv_nn


def kpconv_backward_data(q_pts, s_pts, neighb_inds, kernel_points, extent, dwf, cin):
    """dx [Ns,Cin] = scatter-add of w[n,k,h] * dwf[n,k,:] over the neighbour lists — see aprb_kpconv_backward_data."""
    N.require_cuda()
    q, s, g = _dev_f32(q_pts, "q_pts"), _dev_f32(s_pts, "s_pts"), _dev_f32(dwf, "dwf")
    kp = _dev_f32(kernel_points.detach(), "kernel_points")
    idx, is64, ld = _idx(neighb_inds, "neighb_inds")
    nq, ns, h, k = q.shape[0], s.shape[0], idx.shape[1], kp.shape[0]
    dx = torch.empty((ns, cin), dtype=torch.float32, device=q.device)
    rc = N.lib().aprb_kpconv_backward_data(N.ptr(q), N.ptr(s), N.ptr(idx), is64, ld, N.ptr(kp), float(extent), nq, ns, h, k,
                                           int(cin), N.ptr(g), N.ptr(dx), N.stream_ptr())
    N.check(rc, "aprb_kpconv_backward_data")
    return dx


def instnorm_lrelu_backward(y, dy, rstd, slope):
    """Gradient of y = LeakyReLU_slope(InstanceNorm(x)) w.r.t. x — see aprb_instnorm_lrelu_backward."""
    N.require_cuda()
    yy, gg, rr = _dev_f32(y, "y"), _dev_f32(dy, "dy"), _dev_f32(rstd.reshape(-1), "rstd")
    n, c = yy.shape
    dx = torch.empty_like(yy)
    ws = _workspace(N.lib().aprb_instnorm_backward_ws_bytes(c), yy.device)
    rc = N.lib().aprb_instnorm_lrelu_backward(N.ptr(yy), N.ptr(gg), N.ptr(rr), n, c, float(slope), N.ptr(dx), N.ptr(ws), ws.numel(),
                                              N.stream_ptr())
    N.check(rc, "aprb_instnorm_lrelu_backward")
    return dx


def max_pool_backward(x, inds, dy):
    """Gradient of max_pool w.r.t. x."""
    N.require_cuda()
    xx, g = _dev_f32(x, "x"), _dev_f32(dy, "dy")
    idx, is64, ld = _idx(inds, "inds")
    nq, h = idx.shape
    ns, c = xx.shape
    dx = torch.empty_like(xx)
    rc = N.lib().aprb_max_pool_backward(N.ptr(xx), N.ptr(idx), is64, ld, nq, ns, h, c, N.ptr(g), N.ptr(dx), N.stream_ptr())
    N.check(rc, "aprb_max_pool_backward")
    return dx
