"""Native whole-path runner: pyramid (collate) + KFE encoder blocks in ONE C-ABI call (`aprb_kfe_forward`).

The Python module path (`apr_b200.dataloader.build_pyramid_device` + `apr_b200.architectures.KPFCNNEncoder`) issues
~250 native calls per pair from Python; this wrapper hands the same schedule to `apr_b200/csrc/pipeline.cu`, which
sequences the same kernels from C++ (mirror of datasets/dataloader.py:72-198 and models/architectures.py:149-153).
Several pipelines can run concurrently from several host threads (one CUDA stream each; ctypes releases the GIL).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import blocks, ops


class _Block(C.Structure):
    _fields_ = [("type", C.c_int), ("strided", C.c_int), ("layer", C.c_int), ("in_dim", C.c_int), ("out_dim", C.c_int),
                ("radius", C.c_float), ("extent", C.c_float), ("kp", C.c_void_p), ("kp_W", C.c_void_p),
                ("kp_Wprep", C.c_void_p), ("unary1_W", C.c_void_p), ("unary2_W", C.c_void_p), ("shortcut_W", C.c_void_p),
                ("kp_Wprep16", C.c_void_p), ("kp_Wprep16ck", C.c_void_p)]


class _Config(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("K", C.c_int), ("first_subsampling_dl", C.c_float), ("conv_radius", C.c_float),
                ("limits", C.c_int * 8), ("build_upsamples", C.c_int), ("in_feats_dim", C.c_int),
                ("clouds_per_segment", C.c_int)]


class KFEPipeline:
    """encoder: an `apr_b200.architectures.KPFCNNEncoder` on a CUDA device. The native handle works on SNAPSHOTS of its
    parameters taken at construction (prepared TF32 / fp16 KPConv operands, rounded unary weights, fp16 copies); after
    load_state_dict or an optimizer step call `refresh_weights()` — `forward*` does so by itself when a parameter's
    (data_ptr, version) changed.

    clouds_per_segment=0: the stacked clouds of one call are ONE collate (the reference: one pair, dataloader.py:76).
    clouds_per_segment=2: the call carries P collated pairs stacked (B = 2P clouds); every kernel runs once over the
    super-batch and BatchNormBlock keeps per-pair statistics, so the result equals P separate calls."""

    def __init__(self, encoder, config, neighborhood_limits, build_upsamples=True, stream=None, clouds_per_segment=0):
        N.require_cuda()
        self.lib = N.lib()
        self.config = config
        self.device = next(encoder.parameters()).device
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.encoder = encoder
        self._limits = [int(v) for v in list(neighborhood_limits)[:8]]
        # True: full [N_l, limit] upsample matrices like the reference's collate; "nearest": only their column 0 ([N_l, 1]) —
        # all the reference reads of them (closest_pool: inds[:, 0], blocks.py:71-83); False: none
        self._build_upsamples = 2 if build_upsamples in ("nearest", 2) else (1 if build_upsamples else 0)
        self._cps = int(clouds_per_segment)
        self.handle = None
        self.arena = None
        self._tap = None
        self._host_out = None
        self._build()

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.encoder.parameters())

    def refresh_weights(self):
        """Rebuild the native handle from the encoder's current parameters (waits for the calls in flight)."""
        self.stream.synchronize()
        if self.handle:
            self.lib.aprb_kfe_destroy(self.handle)
            self.handle = None
        self._build()
        if self._tap is not None:
            N.check(self.lib.aprb_kfe_set_tap(self.handle, self._tap.data_ptr(), self._tap.numel()), "aprb_kfe_set_tap")

    def _fresh(self):
        if self._param_key() != self._weights_key:
            self.refresh_weights()

    def _build(self):
        with torch.cuda.device(self.device):
            self._build_on_device()

    def _build_on_device(self):
        encoder, config = self.encoder, self.config
        neighborhood_limits, build_upsamples, clouds_per_segment = self._limits, self._build_upsamples, self._cps
        self._weights_key = self._param_key()
        self._keep = []                      # tensors whose device memory the native handle points into
        blks = []
        with torch.cuda.stream(self.stream):
            for m in encoder.encoder_blocks:
                b = _Block()
                conv = m.KPConv
                b.type = 0 if isinstance(m, blocks.SimpleBlock) else 1
                b.strided = 1 if 'strided' in m.block_name else 0
                b.layer, b.in_dim, b.out_dim = m.layer_ind, m.in_dim, m.out_dim
                b.radius, b.extent = float(conv.radius), float(conv.KP_extent)
                kp = conv.kernel_points.detach().float().contiguous()
                w = conv.weights.detach().float().contiguous()
                self._keep += [kp, w]
                b.kp, b.kp_W = kp.data_ptr(), w.data_ptr()
                if (conv.K * conv.in_channels) % 32 == 0 and conv.out_channels % 16 == 0:
                    wp = ops.kpconv_prepare_weights(w)
                    self._keep.append(wp)
                    b.kp_Wprep = wp.data_ptr()
                if blocks.KPCONV_F16 and ops.kpconv_f16_supported(conv.K, conv.in_channels, conv.out_channels, 0):
                    wp16 = ops.kpconv_prepare_weights_f16(w)          # KPConv on fp16 operands (its input is a norm output)
                    self._keep.append(wp16)
                    b.kp_Wprep16 = wp16.data_ptr()
                    if ops.kpconv_tc_supported(1, conv.K, conv.in_channels, conv.out_channels):
                        wck = ops.kpconv_prepare_weights_f16_ck(w)   # tcgen05 weighting kernel: channel-major weighted tile
                        self._keep.append(wck)
                        b.kp_Wprep16ck = wck.data_ptr()
                for name in ("unary1", "unary2", "unary_shortcut"):
                    sub = getattr(m, name, None)
                    if isinstance(sub, blocks.UnaryBlock):
                        wt = ops.round_tf32(sub.mlp.weight)
                        self._keep.append(wt)
                        setattr(b, {"unary1": "unary1_W", "unary2": "unary2_W", "unary_shortcut": "shortcut_W"}[name],
                                wt.data_ptr())
                blks.append(b)
        cfg = _Config()
        cfg.num_layers, cfg.K = config.num_layers, config.num_kernel_points
        cfg.first_subsampling_dl, cfg.conv_radius = config.first_subsampling_dl, config.conv_radius
        for i, v in enumerate(list(neighborhood_limits)[:8]):
            cfg.limits[i] = int(v)
        cfg.build_upsamples = int(build_upsamples)
        cfg.in_feats_dim = config.in_feats_dim
        cfg.clouds_per_segment = int(clouds_per_segment)   # 2 = a super-batch of collated pairs (per-pair InstanceNorm)
        arr = (_Block * len(blks))(*blks)
        h = C.c_void_p()
        N.check(self.lib.aprb_kfe_create(C.byref(cfg), arr, len(blks), C.byref(h)), "aprb_kfe_create")
        self.handle = h
        self._out_cols = blks[-1].out_dim if blks[-1].type == 1 else blks[-1].out_dim // 2
        self._nblocks = len(blks)
        self.stream.synchronize()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.aprb_kfe_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    LEVEL_RATIO = 0.6   # arena sized for level l <= N * 0.6^l points (measured 0.40-0.42); grown to the bound on demand

    def _ensure_arena(self, n, b, full=False):
        if full:
            need = self.lib.aprb_kfe_arena_bytes(self.handle, int(n), int(b))
        else:
            need = self.lib.aprb_kfe_arena_bytes_est(self.handle, int(n), int(b), C.c_float(self.LEVEL_RATIO))
        if self.arena is None or self.arena.numel() < need:
            # Kernels of earlier calls (this stream and the handle's copy stream) may still be reading the old arena, and
            # the caching allocator would hand its block to the next allocation at once: drain them before the release,
            # and allocate the new arena on the stream that uses it.
            if self.arena is not None:
                self.stream.synchronize()
                for t in (0, 1):
                    self.lib.aprb_kfe_wait_host(self.handle, t)
            self.arena = None
            with torch.cuda.stream(self.stream):
                self.arena = torch.empty(int(need * 1.05) + 4096, dtype=torch.uint8, device=self.device)
        return self.arena

    def _call(self, fn, n, b):
        """fn(arena) -> status; retried once with the unconditional arena bound when a level outgrew the estimate
        (the native driver checks every carve and returns APRB_ERR_WORKSPACE before writing past the arena)."""
        # the C-ABI launches on self.stream: the calling host thread must have that stream's device current (a worker
        # thread of a multi-GPU process starts on device 0)
        with torch.cuda.device(self.device):
            rc = fn(self._ensure_arena(n, b))
            if rc == -2:
                torch.cuda.current_stream(self.device).synchronize(); self.stream.synchronize()
                rc = fn(self._ensure_arena(n, b, full=True))
        return rc

    def _view(self, ptr, rows, cols, dtype):
        off = ptr - self.arena.data_ptr()
        nbytes = rows * cols * 4
        return self.arena[off:off + nbytes].view(dtype).view(rows, cols)

    def forward(self, points, lengths):
        """points [N,3] f32 cuda, lengths [B] i32 cuda -> encoder output [N_last, C] (a view into the arena, valid until
        the next forward). Kernels are queued on self.stream; the call returns without waiting for them."""
        self._fresh()
        pts, lens = points.float().contiguous(), lengths.int().contiguous()
        out, rows, cols = C.c_void_p(), C.c_int(), C.c_int()
        rc = self._call(lambda arena: self.lib.aprb_kfe_forward(
            self.handle, pts.data_ptr(), lens.data_ptr(), None, pts.shape[0], lens.shape[0], arena.data_ptr(), arena.numel(),
            C.byref(out), C.byref(rows), C.byref(cols), C.c_void_p(self.stream.cuda_stream)), pts.shape[0], lens.shape[0])
        N.check(rc, "aprb_kfe_forward")
        self._last_inputs = (pts, lens)
        return self._view(out.value, rows.value, cols.value, torch.float32)

    def forward_host(self, points, lengths, out=None):
        """points [N,3] f32 / lengths [B] i32 HOST tensors (pinned recommended) -> encoder output as a pinned host
        tensor [N_last, C]; does H2D, the whole path and D2H, and synchronises the stream."""
        pts = points if isinstance(points, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(points, np.float32))
        lens = lengths if isinstance(lengths, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(lengths, np.int32))
        self._fresh()
        pts, lens = pts.float().contiguous(), lens.int().contiguous()
        n, b = pts.shape[0], lens.shape[0]
        if out is None:
            cap = min(n, max(n // 4, 4096))         # the last level keeps ~6 % of the level-0 rows
            if self._host_out is None or self._host_out.shape[0] < cap:
                self._host_out = torch.empty((cap, self._out_cols), dtype=torch.float32).pin_memory()
            out = self._host_out
        rows, cols = C.c_int(), C.c_int()
        rc = self._call(lambda arena: self.lib.aprb_kfe_forward_host(
            self.handle, pts.data_ptr(), lens.data_ptr(), n, b, arena.data_ptr(), arena.numel(), out.data_ptr(), out.shape[0],
            C.byref(rows), C.byref(cols), C.c_void_p(self.stream.cuda_stream)), n, b)
        N.check(rc, "aprb_kfe_forward_host")
        return out[:rows.value]

    def forward_host_async(self, points, lengths, out):
        """Like forward_host, but returns as soon as the call is queued: (view of `out`, ticket). The encoder output
        lands in `out` (pinned host tensor [>= N_last, C]) through a side stream while the next call — which must use a
        different `out` — already runs; `wait_host(ticket)` blocks until this call's output is complete. At most two
        calls in flight per pipeline. `out` may be float16: the final activation is rounded to a 10-bit mantissa, so the
        fp16 copy converts back to the same fp32 values (|v| >= 2^-14) and moves half the bytes over PCIe."""
        self._fresh()
        pts, lens = points.float().contiguous(), lengths.int().contiguous()
        n, b = pts.shape[0], lens.shape[0]
        if out.dtype not in (torch.float32, torch.float16):
            raise TypeError("forward_host_async: out must be float32 or float16")
        N.check(self.lib.aprb_kfe_set_host_output_f16(self.handle, 1 if out.dtype == torch.float16 else 0),
                "aprb_kfe_set_host_output_f16")
        rows, cols, ticket = C.c_int(), C.c_int(), C.c_int()
        rc = self._call(lambda arena: self.lib.aprb_kfe_forward_host_async(
            self.handle, pts.data_ptr(), lens.data_ptr(), n, b, arena.data_ptr(), arena.numel(), out.data_ptr(), out.shape[0],
            C.byref(rows), C.byref(cols), C.byref(ticket), C.c_void_p(self.stream.cuda_stream)), n, b)
        N.check(rc, "aprb_kfe_forward_host_async")
        self._last_host_inputs = (pts, lens)                         # keep the host buffers alive until the copy ran
        return out[:rows.value], ticket.value

    def forward_from_raw(self, raw_host, lens_host, out_host):
        """The path from RAW scans: raw_host [N, 4] f32 pinned (x, y, z, reflectance rows of the stacked
        scans, datasets/kitti.py:191-194), lens_host [B] i32 -> H2D, first-level voxelisation at first_subsampling_dl with
        open3d's voxel_down_sample semantics (ops.voxel_downsample_raw), pyramid + encoder, D2H of the fp32 output into
        out_host (pinned [>= rows, C]). Synchronises the stream; returns the filled view of out_host."""
        self._fresh()
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            raw = raw_host.to(self.device, non_blocking=True)
            lens = lens_host.to(self.device, non_blocking=True)
            p0, l0 = ops.voxel_downsample_raw(raw, lens, self.config.first_subsampling_dl)   # reads M back: one sync
            y = self.forward(p0, l0)
            view = out_host[:y.shape[0]]
            view.copy_(y, non_blocking=True)
        self.stream.synchronize()
        return view

    def wait_host(self, ticket):
        with torch.cuda.device(self.device):
            N.check(self.lib.aprb_kfe_wait_host(self.handle, int(ticket)), "aprb_kfe_wait_host")

    def block_output(self, i):
        """Output of encoder block i of the last forward (fp16 or fp32 view into the arena, valid until the next one)."""
        p, r, c, f = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        N.check(self.lib.aprb_kfe_get_block_output(self.handle, int(i), C.byref(p), C.byref(r), C.byref(c), C.byref(f)),
                "aprb_kfe_get_block_output")
        dt, es = (torch.float16, 2) if f.value else (torch.float32, 4)
        off = p.value - self.arena.data_ptr()
        return self.arena[off:off + r.value * c.value * es].view(dt).view(r.value, c.value)

    def set_tap(self, nbytes):
        """Test hook: keep a device copy of every encoder block's output of the following forwards (nbytes of device
        memory; 0 switches it off). Read them with `taps()`."""
        self.stream.synchronize()
        self._tap = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device) if nbytes else None
        N.check(self.lib.aprb_kfe_set_tap(self.handle, self._tap.data_ptr() if nbytes else None, int(nbytes)), "aprb_kfe_set_tap")

    def taps(self):
        """{(block, kind): tensor} of the last forward (views into the tap buffer, fp16 or fp32); kind 'out' = block
        output, 'kp_out' = raw KPConv output, 'kp_in' = KPConv input."""
        out = {}
        for i in range(self.lib.aprb_kfe_tap_count(self.handle)):
            p, r, c, f, tag = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
            N.check(self.lib.aprb_kfe_get_tap(self.handle, i, C.byref(p), C.byref(r), C.byref(c), C.byref(f), C.byref(tag)),
                    "aprb_kfe_get_tap")
            dt, es = (torch.float16, 2) if f.value else (torch.float32, 4)
            off = p.value - self._tap.data_ptr()
            t = self._tap[off:off + r.value * c.value * es].view(dt).view(r.value, c.value)
            out[(tag.value // 4, ("out", "kp_out", "kp_in")[tag.value % 4])] = t
        return out

    def pyramid(self):
        """The batch dict of the last forward (views into the arena): points, neighbors, pools, upsamples, stack_lengths."""
        out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[])
        names = {0: ("points", torch.float32), 1: ("neighbors", torch.int32), 2: ("pools", torch.int32),
                 3: ("upsamples", torch.int32), 4: ("stack_lengths", torch.int32)}
        for lvl in range(self.config.num_layers):
            for what, (key, dt) in names.items():
                p, r, c = C.c_void_p(), C.c_int(), C.c_int()
                N.check(self.lib.aprb_kfe_get(self.handle, what, lvl, C.byref(p), C.byref(r), C.byref(c)), "aprb_kfe_get")
                if not p.value or r.value == 0:
                    t = torch.zeros((0, max(c.value, 1)), dtype=dt, device=self.device)
                elif lvl == 0 and what in (0, 4):          # level-0 points / lengths are the caller's own tensors
                    t = self._last_inputs[0] if what == 0 else self._last_inputs[1].view(-1, 1)
                else:
                    t = self._view(p.value, r.value, c.value, dt)
                out[key].append(t.view(-1) if what == 4 else t)
        return out


class KPFCNNPipeline:
    """BASELINE config 3: the full KPFCNN forward (models/architectures.py:137-212) for P collated pairs stacked in one
    call, stream-ordered on one CUDA stream: the KFE encoder natively (`aprb_kfe_forward`), then the bottleneck Conv1d, the
    GCN (per pair: self / cross attention on the coarsest level, N_3 ~ 2k points — library ops, models/gcn.py), the
    cross-saliency softmax and the nearest-upsample + unary decoder, super-batched with per-pair InstanceNorm segments
    (`ops.closest_pool`, `ops.instnorm_lrelu_seg`, Linear through cuBLAS or the tcgen05 GEMM by shape).
    net: apr_b200.architectures.KPFCNN on the device. Returns (feats_f [N0, D], scores_overlap [N0], scores_saliency [N0])."""

    def __init__(self, net, config, neighborhood_limits, stream=None, clouds_per_segment=2):
        self.net, self.config = net, config
        self.enc = KFEPipeline(net, config, neighborhood_limits, build_upsamples="nearest", stream=stream,
                               clouds_per_segment=clouds_per_segment)
        self.stream, self.device = self.enc.stream, self.enc.device
        self.cps = max(int(clouds_per_segment), 0)

    @torch.no_grad()
    def forward(self, points, lengths):
        import torch.nn.functional as F
        from .gcn import _conv1d
        net, cfg = self.net, self.config
        y = self.enc.forward(points, lengths)                         # encoder output [N_last, C] fp32 (arena view)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            pyr = self.enc.pyramid()
            nlev = cfg.num_layers
            cps = self.cps if self.cps > 0 else int(lengths.shape[0])
            segs = [ops.segment_offsets(pyr['stack_lengths'][l], cps) for l in range(nlev)]
            # skip features = the INPUT of every strided block (architectures.py:149-152) = the previous block's output
            skips = [self.enc.block_output(bi - 1).float() for bi in net.encoder_skips if 0 < bi < len(net.encoder_blocks)]
            # one tiny read-back: rows per cloud and level (the GCN runs per pair; so does InstanceNorm on odd widths)
            lens_h = torch.stack([pyr['stack_lengths'][l] for l in range(nlev)]).cpu()
            lens_c = lens_h[nlev - 1].view(-1, 2) if cps == 2 else lens_h[nlev - 1].view(1, -1)
            pair_rows = [lens_h[l].view(-1, cps).sum(1).tolist() for l in range(nlev)]
            pts_c = pyr['points'][nlev - 1]
            feats_all = _conv1d(net.bottle, y)                          # bottleneck, all pairs at once
            if cps == 2:
                # GCN + cross saliency for all pairs at once on clouds padded to the longest one (gcn.gcn_padded)
                from .gcn import gcn_padded
                lc = lens_h[nlev - 1]                                    # rows per cloud on the coarsest level (host)
                nb, nmax = int(lc.shape[0]), int(lc.max())
                lens_d = pyr['stack_lengths'][nlev - 1].long()
                offs = torch.cumsum(lens_d, 0) - lens_d
                jj = torch.arange(nmax, device=self.device)
                mask = jj[None, :] < lens_d[:, None]                     # [2P, Nmax]
                pad = offs[:, None] + torch.minimum(jj[None, :], (lens_d - 1).clamp_min(0)[:, None])
                fp = gcn_padded(net.gnn, pts_c[pad], feats_all[pad], mask)
                g = _conv1d(net.proj_gnn, fp)                            # [2P, Nmax, G]
                scores = _conv1d(net.proj_score, g)                      # [2P, Nmax, 1]
                fn = F.normalize(g, p=2, dim=2)
                inner = fn[0::2] @ fn[1::2].transpose(1, 2)              # [P, Nmax_src, Nmax_tgt]
                temperature = torch.exp(net.epsilon) + 0.03
                m0, m1 = mask[0::2], mask[1::2]
                s1 = torch.softmax((inner / temperature).masked_fill(~m1[:, None, :], float("-inf")), dim=2) @ scores[1::2]
                s2 = torch.softmax((inner.transpose(1, 2) / temperature).masked_fill(~m0[:, None, :], float("-inf")), dim=2) @ scores[0::2]
                sal = torch.stack([s1, s2], dim=1).reshape(nb, nmax, 1)
                body = g if net.condition else feats_all[pad]
                xp = torch.cat([scores, sal, body] if net.add_cross_overlap else [scores, body], dim=2)
                x = xp[mask]                                              # back to the stacked rows (cloud-major order)
            else:
                rows, a = [], 0
                for pr in range(lens_c.shape[0]):                       # one collate: GCN + cross saliency on its two clouds
                    n_src, n_tgt = int(lens_c[pr, 0]), int(lens_c[pr, 1])
                    f = feats_all[a:a + n_src + n_tgt]
                    p = pts_c[a:a + n_src + n_tgt]
                    f0, f1 = net.gnn(p[:n_src], p[n_src:], f[:n_src], f[n_src:])
                    g = _conv1d(net.proj_gnn, torch.cat([f0, f1], dim=0))
                    scores = _conv1d(net.proj_score, g)
                    fn = F.normalize(g, p=2, dim=1)
                    inner = fn[:n_src] @ fn[n_src:].t()
                    temperature = torch.exp(net.epsilon) + 0.03
                    s1 = torch.softmax(inner / temperature, dim=1) @ scores[n_src:]
                    s2 = torch.softmax(inner.t() / temperature, dim=1) @ scores[:n_src]
                    body = g if net.condition else f
                    rows.append(torch.cat([scores, torch.cat((s1, s2), dim=0), body] if net.add_cross_overlap else [scores, body], dim=1))
                    a += n_src + n_tgt
                x = torch.cat(rows, dim=0)
            layer = nlev - 1
            for i, blk in enumerate(net.decoder_blocks):                # decoder, super-batched
                if i in net.decoder_concats:
                    x = torch.cat([x, skips.pop()], dim=1)
                if isinstance(blk, blocks.NearestUpsampleBlock):
                    x = ops.closest_pool(x, pyr['upsamples'][blk.layer_ind - 1])
                    layer -= 1
                elif isinstance(blk, blocks.LastUnaryBlock):
                    x = blocks._linear(blk.mlp, x)
                else:                                                    # UnaryBlock: Linear -> per-pair InstanceNorm -> LeakyReLU
                    t = blocks._linear(blk.mlp, x)
                    slope = 1.0 if blk.no_relu else 0.1
                    if t.shape[1] % 4 == 0:
                        x = ops.instnorm_lrelu_seg(t, segs[layer], slope=slope)
                    else:                                                # odd widths (g + 2 = 258): the generic kernel, pair by pair
                        x, r0 = torch.empty_like(t), 0
                        for nr in pair_rows[layer]:
                            ops.instnorm_lrelu(t[r0:r0 + nr], slope=slope, out=x[r0:r0 + nr])
                            r0 += nr
            d = net.final_feats_dim
            overlap = net.regular_score(torch.sigmoid(x[:, d]).clamp(0, 1))
            sal = net.regular_score(torch.sigmoid(x[:, d + 1]).clamp(0, 1))
            return F.normalize(x[:, :d], p=2, dim=1), overlap, sal
