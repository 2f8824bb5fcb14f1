"""KPConv network blocks — host-side mirror of /root/reference/Predator_APR/models/blocks.py with the same class
names, constructor signatures, attribute names and state_dict keys (`KPConv.weights` [K,Cin,Cout],
`KPConv.kernel_points` [K,3] non-trainable, `mlp.weight`), so `KPFCNN` checkpoints load unchanged.

Forward passes run on the GPU through libaprb200.so: KPConv -> aprb_kpconv_forward, BatchNormBlock (InstanceNorm)
+ LeakyReLU (+ residual) -> aprb_instnorm_lrelu, max_pool/closest_pool -> aprb_max_pool/aprb_closest_pool.
Inference only (no autograd through the native calls). CPU tensors raise: there is no CPU fallback.
Deformable / modulated KPConv, 'constant'/'gaussian' influence and 'closest' aggregation are not exercised by any
shipped config (configs/train/kitti.yaml:22-28) and raise NotImplementedError.
"""
import math

import torch
import torch.nn as nn
from torch.nn.init import kaiming_uniform_
from torch.nn.parameter import Parameter

from . import ops
from .kernel_points import load_kernels

# nn.Linear of the unary blocks: 'tf32' = our tcgen05 TF32 GEMM when the shape allows, 'fp32' = cuBLAS fp32 (library)
LINEAR_MODE = 'fp32'
KPCONV_F16 = True      # with LINEAR_MODE == 'tf32': KPConv contraction on fp16 operands (ops.kpconv mode 3)
# KPConv contraction: 0 = auto (tcgen05 TF32 when supported), 1 = fp32 CUDA cores, 2 = force tcgen05
KPCONV_MODE = 0


def gather(x, idx, method=2):
    """x[idx] (blocks.py:27-58; the reference's three formulations are equivalent in the forward pass)."""
    return x[idx]


def closest_pool(x, inds):
    """blocks.py:71-83 — features of the nearest support (column 0; relies on distance-sorted neighbours)."""
    return ops.closest_pool(x, inds)


def max_pool(x, inds, width=None):
    """blocks.py:86-102 — max over the neighbourhood, zero shadow row included. `width` (device int32 [1], optional) is
    the reference's matrix width min(max_count, limit) when `inds` is a fixed-width device matrix (dataloader.
    build_pyramid_device's batch['pool_widths']): columns beyond it are all-pad artefacts and do not take part."""
    return ops.max_pool(x, inds, width_dev=width)


def global_average(x, batch_lengths):
    """blocks.py:105-127"""
    out, i0 = [], 0
    for length in batch_lengths:
        length = int(length)
        out.append(torch.mean(x[i0:i0 + length], dim=0))
        i0 += length
    return torch.stack(out)


class KPConv(nn.Module):
    """blocks.py:135-379 (rigid, KP_influence='linear', aggregation_mode='sum')."""

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False):
        super(KPConv, self).__init__()
        if deformable or modulated:
            raise NotImplementedError("deformable / modulated KPConv is outside the accelerated path")
        if KP_influence != 'linear' or aggregation_mode != 'sum':
            raise NotImplementedError("only KP_influence='linear', aggregation_mode='sum' (the shipped configs)")
        self.K = kernel_size
        self.p_dim = p_dim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.radius = radius
        self.KP_extent = KP_extent
        self.fixed_kernel_points = fixed_kernel_points
        self.KP_influence = KP_influence
        self.aggregation_mode = aggregation_mode
        self.deformable = deformable
        self.modulated = modulated
        self.min_d2 = None
        self.deformed_KP = None
        self.offset_features = None
        self.weights = Parameter(torch.zeros((self.K, in_channels, out_channels), dtype=torch.float32),
                                 requires_grad=True)
        self.offset_dim = None
        self.offset_conv = None
        self.offset_bias = None
        self.reset_parameters()
        self.kernel_points = self.init_KP()
        self._wprep = None          # prepared TF32 operand, rebuilt when the weights change
        self._wprep_key = None

    def reset_parameters(self):
        kaiming_uniform_(self.weights, a=math.sqrt(5))          # blocks.py:208-212

    def init_KP(self):
        kp = load_kernels(self.radius, self.K, dimension=self.p_dim, fixed=self.fixed_kernel_points)
        return Parameter(torch.tensor(kp, dtype=torch.float32), requires_grad=False)

    def _prepared(self):
        key = (self.weights.data_ptr(), self.weights._version, self.weights.device)
        if self._wprep is None or self._wprep_key != key:
            self._wprep = ops.kpconv_prepare_weights(self.weights)
            self._wprep_key = key
        return self._wprep

    def _prepared_f16(self):
        key = (self.weights.data_ptr(), self.weights._version, self.weights.device)
        if getattr(self, '_wprep16', None) is None or self._wprep16_key != key:
            self._wprep16 = ops.kpconv_prepare_weights_f16(self.weights)
            self._wprep16_key = key
        return self._wprep16

    def forward(self, q_pts, s_pts, neighb_inds, x):
        # fp16 operands (same mantissa as TF32, half the bytes of the weighted tile) in the regime where every KPConv
        # input with Cin > 1 is an InstanceNorm output: the TF32 module path (LINEAR_MODE == 'tf32')
        if (KPCONV_F16 and KPCONV_MODE == 0 and LINEAR_MODE == 'tf32'
                and ops.kpconv_f16_supported(self.K, self.in_channels, self.out_channels, neighb_inds.shape[1])):
            return ops.kpconv(q_pts, s_pts, neighb_inds, x, self.kernel_points, self.weights, self.KP_extent,
                              wprep=self._prepared_f16(), mode=3)
        wprep = None
        if KPCONV_MODE != 1 and (self.K * self.in_channels) % 32 == 0 and self.out_channels % 16 == 0:
            wprep = self._prepared()
        return ops.kpconv(q_pts, s_pts, neighb_inds, x, self.kernel_points, self.weights, self.KP_extent,
                          wprep=wprep, mode=KPCONV_MODE if wprep is not None else 1)

    def __repr__(self):
        return 'KPConv(radius: {:.2f}, extent: {:.2f}, in_feat: {:d}, out_feat: {:d})'.format(
            self.radius, self.KP_extent, self.in_channels, self.out_channels)


def block_decider(block_name, radius, in_dim, out_dim, layer_ind, config):
    """blocks.py:387-433"""
    if block_name == 'unary':
        return UnaryBlock(in_dim, out_dim, config.use_batch_norm, config.batch_norm_momentum)
    if block_name == 'last_unary':
        if config.switch_to_decoder and config.symmetric:
            return LastUnaryBlock(in_dim, config.point_generation_ratio * 3, config.use_batch_norm,
                                  config.batch_norm_momentum)
        return LastUnaryBlock(in_dim, config.final_feats_dim + 2, config.use_batch_norm, config.batch_norm_momentum)
    if block_name in ['simple', 'simple_deformable', 'simple_invariant', 'simple_equivariant', 'simple_strided',
                      'simple_deformable_strided', 'simple_invariant_strided', 'simple_equivariant_strided']:
        return SimpleBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name in ['resnetb', 'resnetb_invariant', 'resnetb_equivariant', 'resnetb_deformable', 'resnetb_strided',
                      'resnetb_deformable_strided', 'resnetb_equivariant_strided', 'resnetb_invariant_strided']:
        return ResnetBottleneckBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name == 'max_pool' or block_name == 'max_pool_wide':
        return MaxPoolBlock(layer_ind)
    if block_name == 'global_average':
        return GlobalAverageBlock()
    if block_name == 'nearest_upsample':
        return NearestUpsampleBlock(layer_ind)
    raise ValueError('Unknown block name in the architecture definition : ' + block_name)


class BatchNormBlock(nn.Module):
    """blocks.py:436-473 — despite the name, nn.InstanceNorm1d over ALL rows of the stacked pair (no affine, no
    running stats, eps 1e-5), or a learned bias when use_bn is False. `fused(x, slope, residual, norm_residual)` is
    the entry the enclosing blocks use to fold the following LeakyReLU / residual add into the same kernel."""

    def __init__(self, in_dim, use_bn, bn_momentum):
        super(BatchNormBlock, self).__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.in_dim = in_dim
        if self.use_bn:
            self.batch_norm = nn.InstanceNorm1d(in_dim, momentum=bn_momentum)   # stateless; kept for repr/state parity
        else:
            self.bias = Parameter(torch.zeros(in_dim, dtype=torch.float32), requires_grad=True)

    def reset_parameters(self):
        nn.init.zeros_(self.bias)

    def fused(self, x, slope=1.0, residual=None, norm_residual=False):
        if self.use_bn:
            if x.shape[1] % 4 == 0:      # two-launch kernels; the module path is one collate = one segment over all rows
                return ops.instnorm_lrelu_seg(x, None, slope=slope, residual=residual,
                                              norm_residual=norm_residual, round_tf32=(LINEAR_MODE == 'tf32'))
            return ops.instnorm_lrelu(x, slope=slope, residual=residual, norm_residual=norm_residual,
                                      round_tf32=(LINEAR_MODE == 'tf32'))
        y = x + self.bias
        if residual is not None:
            y = y + residual
        return y if slope == 1.0 else nn.functional.leaky_relu(y, slope)

    def forward(self, x):
        return self.fused(x)

    def __repr__(self):
        return 'BatchNormBlock(in_feat: {:d}, momentum: {:.3f}, only_bias: {:s})'.format(
            self.in_dim, self.bn_momentum, str(not self.use_bn))


def _linear(mlp, x):
    if LINEAR_MODE == 'tf32' and ops.linear_tf32_supported(x.shape[0], mlp.in_features, mlp.out_features):
        key = (mlp.weight.data_ptr(), mlp.weight._version)
        cache = getattr(mlp, '_aprb_w_tf32', None)
        if cache is None or cache[0] != key:                 # TF32-rounded copy, rebuilt when the weight changes
            cache = (key, ops.round_tf32(mlp.weight))
            mlp._aprb_w_tf32 = cache
        return ops.linear_tf32(x, cache[1])
    return nn.functional.linear(x, mlp.weight)


class UnaryBlock(nn.Module):
    """blocks.py:476-510 — Linear(no bias) -> InstanceNorm -> LeakyReLU(0.1) (unless no_relu)."""

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super(UnaryBlock, self).__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.no_relu = no_relu
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=False)
        self.batch_norm = BatchNormBlock(out_dim, self.use_bn, self.bn_momentum)
        if not no_relu:
            self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, batch=None):
        x = _linear(self.mlp, x)
        return self.batch_norm.fused(x, slope=1.0 if self.no_relu else 0.1)

    def __repr__(self):
        return 'UnaryBlock(in_feat: {:d}, out_feat: {:d}, BN: {:s}, ReLU: {:s})'.format(
            self.in_dim, self.out_dim, str(self.use_bn), str(not self.no_relu))


class LastUnaryBlock(nn.Module):
    """blocks.py:513-536"""

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super(LastUnaryBlock, self).__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=False)

    def forward(self, x, batch=None):
        return _linear(self.mlp, x)

    def __repr__(self):
        return 'LastUnaryBlock(in_feat: {:d}, out_feat: {:d})'.format(self.in_dim, self.out_dim)


def _select(block_name, layer_ind, batch):
    """(q_pts, s_pts, neighb_inds) by block type — blocks.py:583-590 / :655-662."""
    if 'strided' in block_name:
        return batch['points'][layer_ind + 1], batch['points'][layer_ind], batch['pools'][layer_ind]
    return batch['points'][layer_ind], batch['points'][layer_ind], batch['neighbors'][layer_ind]


class SimpleBlock(nn.Module):
    """blocks.py:539-593 — KPConv(in -> out/2) -> InstanceNorm -> LeakyReLU(0.1)."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super(SimpleBlock, self).__init__()
        current_extent = radius * config.KP_extent / config.conv_radius
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.layer_ind = layer_ind
        self.block_name = block_name
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.KPConv = KPConv(config.num_kernel_points, config.in_points_dim, in_dim, out_dim // 2, current_extent,
                             radius, fixed_kernel_points=config.fixed_kernel_points,
                             KP_influence=config.KP_influence, aggregation_mode=config.aggregation_mode,
                             deformable='deform' in block_name, modulated=config.modulated)
        self.batch_norm = BatchNormBlock(out_dim // 2, self.use_bn, self.bn_momentum)
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, batch):
        q_pts, s_pts, neighb_inds = _select(self.block_name, self.layer_ind, batch)
        x = self.KPConv(q_pts, s_pts, neighb_inds, x)
        return self.batch_norm.fused(x, slope=0.1)


class ResnetBottleneckBlock(nn.Module):
    """blocks.py:596-681 — unary1 -> KPConv -> IN+LReLU -> unary2(no relu) (+ shortcut) -> LReLU."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super(ResnetBottleneckBlock, self).__init__()
        current_extent = radius * config.KP_extent / config.conv_radius
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.block_name = block_name
        self.layer_ind = layer_ind
        self.in_dim = in_dim
        self.out_dim = out_dim
        if in_dim != out_dim // 4:
            self.unary1 = UnaryBlock(in_dim, out_dim // 4, self.use_bn, self.bn_momentum)
        else:
            self.unary1 = nn.Identity()
        self.KPConv = KPConv(config.num_kernel_points, config.in_points_dim, out_dim // 4, out_dim // 4,
                             current_extent, radius, fixed_kernel_points=config.fixed_kernel_points,
                             KP_influence=config.KP_influence, aggregation_mode=config.aggregation_mode,
                             deformable='deform' in block_name, modulated=config.modulated)
        self.batch_norm_conv = BatchNormBlock(out_dim // 4, self.use_bn, self.bn_momentum)
        self.unary2 = UnaryBlock(out_dim // 4, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
        if in_dim != out_dim:
            self.unary_shortcut = UnaryBlock(in_dim, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
        else:
            self.unary_shortcut = nn.Identity()
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, features, batch):
        q_pts, s_pts, neighb_inds = _select(self.block_name, self.layer_ind, batch)
        x = self.unary1(features)
        x = self.KPConv(q_pts, s_pts, neighb_inds, x)
        x = self.batch_norm_conv.fused(x, slope=0.1)
        if 'strided' in self.block_name:
            widths = batch.get('pool_widths') if isinstance(batch, dict) else None
            shortcut = max_pool(features, neighb_inds, widths[self.layer_ind] if widths else None)
        else:
            shortcut = features
        x2 = _linear(self.unary2.mlp, x)
        if isinstance(self.unary_shortcut, nn.Identity):
            # LeakyReLU(IN(x2) + shortcut) in one pass
            return self.unary2.batch_norm.fused(x2, slope=0.1, residual=shortcut)
        sc = _linear(self.unary_shortcut.mlp, shortcut)
        if self.use_bn:
            # LeakyReLU(IN(x2) + IN(sc)) in one pass (both branches standardised with their own statistics)
            return self.unary2.batch_norm.fused(x2, slope=0.1, residual=sc, norm_residual=True)
        sc = self.unary_shortcut.batch_norm.fused(sc)
        return self.unary2.batch_norm.fused(x2, slope=0.1, residual=sc)


class GlobalAverageBlock(nn.Module):
    """blocks.py:684-694"""

    def __init__(self):
        super(GlobalAverageBlock, self).__init__()

    def forward(self, x, batch):
        return global_average(x, batch['stack_lengths'][-1])


class NearestUpsampleBlock(nn.Module):
    """blocks.py:697-712"""

    def __init__(self, layer_ind):
        super(NearestUpsampleBlock, self).__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        return closest_pool(x, batch['upsamples'][self.layer_ind - 1])

    def __repr__(self):
        return 'NearestUpsampleBlock(layer: {:d} -> {:d})'.format(self.layer_ind, self.layer_ind - 1)


class MaxPoolBlock(nn.Module):
    """blocks.py:715-726"""

    def __init__(self, layer_ind):
        super(MaxPoolBlock, self).__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        widths = batch.get('pool_widths') if isinstance(batch, dict) else None
        return max_pool(x, batch['pools'][self.layer_ind + 1], widths[self.layer_ind + 1] if widths else None)
