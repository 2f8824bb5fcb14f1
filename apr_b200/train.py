"""Training path of the KFE network (BASELINE config 5): autograd for the KPConv stack, the
reference's NPR generative head and loss terms, and a bucketed gradient all-reduce for one-process-per-GPU data parallel.

What runs where:
  * KPConv forward/backward  -> our kernels for the gather/weighting (aprb_kpconv_weighted) and its transpose
    (aprb_kpconv_backward_data, scatter-add over the neighbour lists); the dense contractions out = wf @ W,
    dwf = g @ W^T and dW = wf^T @ g on the hand-written tcgen05 TF32 GEMM (aprb_linear_tf32) where the shape allows;
  * strided shortcut          -> aprb_max_pool / aprb_max_pool_backward;
  * InstanceNorm, LeakyReLU, Linear, nearest upsample, the bottleneck GCN, the NPR MLP -> stock torch (autograd-native).
The reference trains on ONE GPU (lib/trainer.py:316-322 accumulates `iter_size` pairs); averaging per-rank gradients
with an all-reduce is the same mean over pairs, spread over the ranks.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import blocks, ops
from .gcn import _conv1d


# The three dense contractions of a KPConv training step — out = wf @ W, dwf = g @ W^T, dW = wf^T @ g — on the hand-written
# tcgen05 GEMM (TF32 operands rounded to nearest, fp32 accumulation in TMEM; aprb_linear_tf32, split-K when the output grid
# cannot fill the SMs) where the shapes allow (K % 32 == 0, N % 16 == 0), else torch.matmul (fp32). False = fp32 everywhere.
TENSOR_GEMM = True


def _mm_nt(a, bt):
    """a [M,K] @ bt [N,K]^T -> [M,N]"""
    if TENSOR_GEMM and a.is_cuda and a.shape[0] > 0 and ops.linear_tf32_supported(a.shape[0], a.shape[1], bt.shape[0]):
        return ops.linear_tf32(ops.round_tf32(a), ops.round_tf32(bt))
    return a @ bt.t()


def _mm_tn(a, b):
    """a [R,M]^T @ b [R,N] -> [M,N] (reduction over the rows): both operands transposed into K-major form, the reduction
    dimension zero-padded to a multiple of 32"""
    r = a.shape[0]
    if TENSOR_GEMM and a.is_cuda and r > 0 and b.shape[1] % 16 == 0:
        rp = (r + 31) // 32 * 32
        at = torch.zeros((a.shape[1], rp), dtype=torch.float32, device=a.device)
        bt = torch.zeros((b.shape[1], rp), dtype=torch.float32, device=a.device)
        at[:, :r] = a.t(); bt[:, :r] = b.t()
        return ops.linear_tf32(ops.round_tf32(at), ops.round_tf32(bt))
    return a.t() @ b


class _KPConvFn(torch.autograd.Function):
    """KPConv.forward (models/blocks.py:229-374) with its gradients w.r.t. the features and the weights."""

    @staticmethod
    def forward(ctx, x, weights, q, s, idx, kp, extent):
        wf, inv_nn = ops.kpconv_weighted(q, s, idx, x, kp, extent)
        k, cin, cout = weights.shape
        w2d = weights.reshape(k * cin, cout)
        ctx.save_for_backward(wf, inv_nn, w2d, q, s, idx, kp)
        ctx.extent, ctx.cin, ctx.wshape = extent, cin, weights.shape
        return _mm_nt(wf, w2d.t().contiguous()) * inv_nn.unsqueeze(1)

    @staticmethod
    def backward(ctx, dout):
        wf, inv_nn, w2d, q, s, idx, kp = ctx.saved_tensors
        g = dout.contiguous() * inv_nn.unsqueeze(1)
        dw = _mm_tn(wf, g).reshape(ctx.wshape) if ctx.needs_input_grad[1] else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.kpconv_backward_data(q, s, idx, kp, ctx.extent, _mm_nt(g, w2d.contiguous()), ctx.cin)
        return dx, dw, None, None, None, None, None


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(x, idx)
        return ops.max_pool(x, idx)

    @staticmethod
    def backward(ctx, dy):
        x, idx = ctx.saved_tensors
        return ops.max_pool_backward(x, idx, dy.contiguous()), None


def kpconv(x, conv, q, s, idx):
    return _KPConvFn.apply(x, conv.weights, q, s, idx, conv.kernel_points, float(conv.KP_extent))


class _NormActFn(torch.autograd.Function):
    """y = LeakyReLU_slope(InstanceNorm(x)) over all rows of the pair (BatchNormBlock, blocks.py:459-468, and the
    activation that follows it): forward by the fused inference kernel (aprb_instnorm_lrelu), backward by
    aprb_instnorm_lrelu_backward from y and the per-column rstd — x itself is not kept."""

    @staticmethod
    def forward(ctx, x, slope):
        x = x.contiguous()
        var = torch.var(x, dim=0, unbiased=False)
        y = ops.instnorm_lrelu(x, slope=slope)
        ctx.save_for_backward(y, torch.rsqrt(var + 1e-5))
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        y, rstd = ctx.saved_tensors
        return ops.instnorm_lrelu_backward(y, dy.contiguous(), rstd, ctx.slope), None


NATIVE_NORM = True     # InstanceNorm (+ LeakyReLU) forward/backward through the native kernels (False: stock torch autograd)


def _norm(x, slope=1.0, eps=1e-5):
    """LeakyReLU_slope(BatchNormBlock(x)); BatchNormBlock = InstanceNorm over all rows of the pair (blocks.py:459-468)."""
    if NATIVE_NORM and x.is_cuda and x.shape[0] > 1:
        return _NormActFn.apply(x, float(slope))
    var, mean = torch.var_mean(x, dim=0, unbiased=False, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps)
    return y if slope == 1.0 else F.leaky_relu(y, slope)


def _unary(blk, x):
    y = F.linear(x, blk.mlp.weight)
    if isinstance(blk, blocks.LastUnaryBlock):
        return y
    if blk.use_bn:
        return _norm(y, 1.0 if blk.no_relu else 0.1)
    y = y + blk.batch_norm.bias
    return y if blk.no_relu else F.leaky_relu(y, 0.1)


def _closest(x, inds):
    return torch.cat((x, torch.zeros_like(x[:1])), 0)[inds[:, 0].long()]


def block_forward(blk, x, batch):
    """Differentiable forward of one block module of apr_b200.blocks (same parameters, same semantics)."""
    if isinstance(blk, blocks.SimpleBlock):
        q, s, idx = blocks._select(blk.block_name, blk.layer_ind, batch)
        return _norm(kpconv(x, blk.KPConv, q, s, idx), 0.1)
    if isinstance(blk, blocks.ResnetBottleneckBlock):
        q, s, idx = blocks._select(blk.block_name, blk.layer_ind, batch)
        y = _unary(blk.unary1, x) if isinstance(blk.unary1, blocks.UnaryBlock) else x
        y = _norm(kpconv(y, blk.KPConv, q, s, idx), 0.1)
        y = _unary(blk.unary2, y)
        sc = _MaxPoolFn.apply(x, idx) if 'strided' in blk.block_name else x
        if isinstance(blk.unary_shortcut, blocks.UnaryBlock):
            sc = _unary(blk.unary_shortcut, sc)
        return F.leaky_relu(y + sc, 0.1)
    if isinstance(blk, (blocks.UnaryBlock, blocks.LastUnaryBlock)):
        return _unary(blk, x)
    if isinstance(blk, blocks.NearestUpsampleBlock):
        return _closest(x, batch['upsamples'][blk.layer_ind - 1])
    raise NotImplementedError(type(blk).__name__)


def kpfcnn_forward_train(net, batch):
    """KPFCNN.forward (architectures.py:137-212) with gradients; `net` is an apr_b200.architectures.KPFCNN."""
    x = batch['features']
    skips = []
    for i, blk in enumerate(net.encoder_blocks):
        if i in net.encoder_skips:
            skips.append(x)
        x = block_forward(blk, x, batch)
    n_src = int(batch['stack_lengths'][-1][0])
    pts_c = batch['points'][-1]
    feats = _conv1d(net.bottle, x)
    unconditioned = feats
    f0, f1 = net.gnn(pts_c[:n_src], pts_c[n_src:], feats[:n_src], feats[n_src:])
    feats = _conv1d(net.proj_gnn, torch.cat([f0, f1], dim=0))
    scores = _conv1d(net.proj_score, feats)
    fn = F.normalize(feats, p=2, dim=1)
    inner = fn[:n_src] @ fn[n_src:].t()
    temperature = torch.exp(net.epsilon) + 0.03
    s1 = torch.softmax(inner / temperature, dim=1) @ scores[n_src:]
    s2 = torch.softmax(inner.t() / temperature, dim=1) @ scores[:n_src]
    body = feats if net.condition else unconditioned
    x = torch.cat([scores, torch.cat((s1, s2), 0), body] if net.add_cross_overlap else [scores, body], dim=1)
    for i, blk in enumerate(net.decoder_blocks):
        if i in net.decoder_concats:
            x = torch.cat([x, skips.pop()], dim=1)
        x = block_forward(blk, x, batch)
    d = net.final_feats_dim
    return F.normalize(x[:, :d], p=2, dim=1), torch.sigmoid(x[:, d]).clamp(0, 1), torch.sigmoid(x[:, d + 1]).clamp(0, 1)


class NPRHead(nn.Module):
    """GenerativeMLP_98 (models/mlp.py:104-157): three Linear -> ReLU -> BatchNorm1d stages, final_feats_dim -> 512 ->
    256 -> 3 * point_generation_ratio offsets per point (trainer.py:158-166)."""

    def __init__(self, in_channel=32, out_points=4, bn_momentum=0.02, channels=(512, 256)):
        super().__init__()
        dims = [in_channel, *channels, out_points * 3]
        self.list_modules = nn.ModuleList(
            nn.Sequential(nn.Linear(a, b), nn.ReLU(), nn.BatchNorm1d(b, momentum=bn_momentum)) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for m in self.list_modules:
            x = m(x)
        return x


def _nearest_sq(a, b, chunk=4096):
    """sum_i min_j |a_i - b_j|^2 (differentiable w.r.t. a and b through the selected pairs)."""
    with torch.no_grad():
        arg = torch.cat([torch.cdist(a[i:i + chunk], b).argmin(dim=1) for i in range(0, a.shape[0], chunk)])
    return ((a - b[arg]) ** 2).sum()


def chamfer(a, b):
    """trainer.py:131-140 — forward/n1 + backward/n2 of chamferdist.ChamferDistance (third-party, chamferdist==1.0.0 in
    the reference's requirements; sum-of-nearest-squared-distances semantics, parity unpinned)."""
    return _nearest_sq(a, b) / a.shape[0] + _nearest_sq(b, a) / b.shape[0]


def npr_loss(head, feats, pts, apc, ratio=4, reg_strength=0.01, loss_ratio=0.001):
    """Generative loss of one frame (trainer.py:158-172): chamfer(points + generated offsets, aggregated cloud) +
    reg_strength * mean |offset|^2, scaled by loss_ratio."""
    gen = head(feats)
    reg = (gen.reshape(-1, 3) ** 2).sum(-1).mean()
    mod = (gen + pts.repeat(1, ratio)).reshape(-1, 3)
    return (chamfer(mod, apc) + reg * reg_strength) * loss_ratio


def surrogate_desc_loss(feats_src, feats_tgt, corr, overlap, saliency, n_src):
    """Stand-in for lib/loss.py (circle / overlap / saliency losses; out of scope): cosine attraction on the
    ground-truth correspondences, repulsion on shuffled pairs, BCE of both score heads against 'has a correspondence'."""
    lab = torch.zeros_like(overlap)
    loss = overlap.sum() * 0
    if corr.shape[0] > 0:
        lab[corr[:, 0]] = 1.0
        lab[n_src + corr[:, 1]] = 1.0
        fs, ft = feats_src[corr[:, 0]], feats_tgt[corr[:, 1]]
        loss = (1 - (fs * ft).sum(1)).mean() + F.relu((fs * ft.roll(1, 0)).sum(1) - 0.1).mean()
    return loss + F.binary_cross_entropy(overlap, lab) + F.binary_cross_entropy(saliency, lab)


class GradBucketReducer:
    """Bucketed, hook-driven gradient averaging over a process group: parameters are packed (in reverse registration
    order, the order their gradients become ready) into ~bucket_mb flat buckets; when the last gradient of a bucket has
    been accumulated its all-reduce is launched asynchronously, overlapping the rest of backward. finish() waits for all
    buckets and writes the averaged gradients back. Works with NCCL (GPU) and gloo (CPU tests)."""

    def __init__(self, params, group=None, bucket_mb=25.0):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p); size += p.numel() * 4
            if size >= bucket_mb * (1 << 20):
                self.buckets.append(cur); cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device) for b in self.buckets]
        self.where = {p: (bi, off) for bi, b in enumerate(self.buckets)
                      for p, off in zip(b, [sum(q.numel() for q in b[:j]) for j in range(len(b))])}
        self.pending = [len(b) for b in self.buckets]
        self.work = [None] * len(self.buckets)
        self.fired = set()
        self.launched = 0
        self.sync = True            # False inside no_sync(): gradients accumulate locally, nothing is reduced
        self.fired = set()
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def no_sync(self):
        """Context manager for gradient accumulation (the reference's iter_size > 1, lib/trainer.py:316-322): backward
        passes inside it only accumulate into p.grad; the LAST micro-batch runs outside it and reduces the sums."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self.sync = self.sync, False
            try:
                yield
            finally:
                self.sync = prev
        return ctx()

    def _hook(self, p):
        if not self.sync or p in self.fired:
            return
        bi, off = self.where[p]
        self.fired.add(p)
        self.flat[bi][off:off + p.numel()].copy_(p.grad.reshape(-1))   # post-accumulate: p.grad holds the accumulated sum
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        if self.world > 1:
            self.work[bi] = self.dist.all_reduce(self.flat[bi], group=self.group, async_op=True)
        self.launched += 1

    def finish(self):
        """Call after backward(): reduces buckets whose parameters got no gradient this step too, waits, averages."""
        for bi, b in enumerate(self.buckets):       # fixed bucket order on every rank
            if self.pending[bi] > 0:
                for p in b:                      # parameters whose hook did not fire this step: their current gradient, or zeros
                    if p not in self.fired:
                        o = self.where[p][1]
                        if p.grad is None:
                            self.flat[bi][o:o + p.numel()].zero_()
                        else:
                            self.flat[bi][o:o + p.numel()].copy_(p.grad.reshape(-1))
                self._launch(bi)
        for bi, b in enumerate(self.buckets):
            if self.work[bi] is not None:
                self.work[bi].wait()
            self.flat[bi].div_(self.world)
            for p in b:
                o = self.where[p][1]
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(self.flat[bi][o:o + p.numel()].view_as(p))
        self.pending = [len(b) for b in self.buckets]
        self.work = [None] * len(self.buckets)
        self.fired = set()
