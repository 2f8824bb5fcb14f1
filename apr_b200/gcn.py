"""Bottleneck graph network of KPFCNN — mirror of /root/reference/Predator_APR/models/gcn.py (class GCN :173-210 and
its layers) with the same module tree, so the reference's `gnn.*` state_dict keys load unchanged.

This part runs on the coarsest level only (N_3 ~ 1-2k points per cloud): it is small dense work, done with stock torch
ops on the device. What changes against the reference is the formulation, not the result:
  * rows are points ([N, C] row-major, the layout of the KPConv stack) instead of [1, C, N];
  * the DGCNN edge feature conv1x1(cat(x_i, x_j - x_i)) (gcn.py:9-35, :62-68) is evaluated as
    (W_a - W_b) x_i + W_b x_j — two [N,C] GEMMs and a k-row gather instead of the reference's [1,C,N,N] repeat and a
    GEMM over N*k edge features (k = 10 times fewer flops, no O(N^2 C) temporary);
  * attention heads use the reference's channel split c = d * num_heads + h (gcn.py:113-114).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def knn_indices(coords, k):
    """[N,3] -> [N,k] indices of the k nearest OTHER points, by the reference's distance expression
    -2ab + |a|^2 + |b|^2 clamped at 1e-12 (lib/utils.py:78-97), ascending, first hit (the point itself) dropped
    (gcn.py:22-23)."""
    sq = (coords * coords).sum(dim=1)
    dist = (-2.0 * coords @ coords.t() + sq[:, None] + sq[None, :]).clamp_min(1e-12)
    return dist.topk(k + 1, dim=1, largest=False, sorted=True)[1][:, 1:]


def _edge_conv(x, idx, weight, slope=0.2, eps=1e-5):
    """max_j LeakyReLU(InstanceNorm2d(conv1x1(cat(x_i, x_j - x_i)))) over the k graph neighbours j of i.
    x [N,C], idx [N,k], weight [Cout, 2C(,1,1)] -> [N,Cout]. Statistics per channel over all N*k edges (gcn.py:64)."""
    w = weight.reshape(weight.shape[0], -1)
    c = x.shape[1]
    wa, wb = w[:, :c], w[:, c:]
    centre = x @ (wa - wb).t()                    # [N,Cout]
    neigh = x @ wb.t()                            # [N,Cout]
    e = centre.unsqueeze(1) + neigh[idx]          # [N,k,Cout]
    mean = e.mean(dim=(0, 1), keepdim=True)
    var = e.var(dim=(0, 1), unbiased=False, keepdim=True)
    e = F.leaky_relu((e - mean) * torch.rsqrt(var + eps), slope)
    return e.max(dim=1)[0]


class SelfAttention(nn.Module):
    """gcn.py:39-79 — two DGCNN edge convolutions on the coordinate k-NN graph, then a pointwise fusion layer."""

    def __init__(self, feature_dim, k=10):
        super().__init__()
        self.conv1 = nn.Conv2d(feature_dim * 2, feature_dim, kernel_size=1, bias=False)
        self.in1 = nn.InstanceNorm2d(feature_dim)
        self.conv2 = nn.Conv2d(feature_dim * 2, feature_dim * 2, kernel_size=1, bias=False)
        self.in2 = nn.InstanceNorm2d(feature_dim * 2)
        self.conv3 = nn.Conv2d(feature_dim * 4, feature_dim, kernel_size=1, bias=False)
        self.in3 = nn.InstanceNorm2d(feature_dim)
        self.k = k

    def forward(self, coords, feats):
        """coords [N,3], feats [N,C] -> [N,C]"""
        idx = knn_indices(coords, self.k)
        x1 = _edge_conv(feats, idx, self.conv1.weight)
        x2 = _edge_conv(x1, idx, self.conv2.weight)
        x3 = torch.cat((feats, x1, x2), dim=1) @ self.conv3.weight.reshape(self.conv3.weight.shape[0], -1).t()
        mean = x3.mean(dim=0, keepdim=True)
        var = x3.var(dim=0, unbiased=False, keepdim=True)
        return F.leaky_relu((x3 - mean) * torch.rsqrt(var + 1e-5), 0.2)


def _conv1d(layer, x):
    """nn.Conv1d(kernel_size=1) applied to row-major points: [N,Cin] -> [N,Cout]."""
    return F.linear(x, layer.weight.squeeze(-1), layer.bias)


class MultiHeadedAttention(nn.Module):
    """gcn.py:101-116"""

    def __init__(self, num_heads, d_model):
        super().__init__()
        assert d_model % num_heads == 0
        self.dim = d_model // num_heads
        self.num_heads = num_heads
        self.merge = nn.Conv1d(d_model, d_model, kernel_size=1)
        self.proj = nn.ModuleList([nn.Conv1d(d_model, d_model, kernel_size=1) for _ in range(3)])

    def forward(self, query, key, value):
        """[Nq,C], [Nk,C], [Nk,C] -> [Nq,C]; channel c belongs to head c % num_heads (view(b, dim, heads, n))."""
        q, k, v = [_conv1d(l, t).view(t.shape[0], self.dim, self.num_heads).permute(2, 0, 1)
                   for l, t in zip(self.proj, (query, key, value))]          # [heads, N, dim]
        prob = torch.softmax(q @ k.transpose(1, 2) / self.dim ** .5, dim=-1)  # [heads, Nq, Nk]
        out = (prob @ v).permute(1, 2, 0).reshape(query.shape[0], self.dim * self.num_heads)
        return _conv1d(self.merge, out)


class AttentionalPropagation(nn.Module):
    """gcn.py:119-129 — message = attention(x, source); MLP(cat(x, message)) with InstanceNorm1d + ReLU in between."""

    def __init__(self, feature_dim, num_heads):
        super().__init__()
        self.attn = MultiHeadedAttention(num_heads, feature_dim)
        self.mlp = nn.Sequential(nn.Conv1d(feature_dim * 2, feature_dim * 2, kernel_size=1),
                                 nn.InstanceNorm1d(feature_dim * 2), nn.ReLU(),
                                 nn.Conv1d(feature_dim * 2, feature_dim, kernel_size=1))
        nn.init.constant_(self.mlp[-1].bias, 0.0)

    def forward(self, x, source):
        h = _conv1d(self.mlp[0], torch.cat([x, self.attn(x, source, source)], dim=1))
        mean = h.mean(dim=0, keepdim=True)
        var = h.var(dim=0, unbiased=False, keepdim=True)
        return _conv1d(self.mlp[3], F.relu((h - mean) * torch.rsqrt(var + 1e-5)))


class GCN(nn.Module):
    """gcn.py:173-210 — alternates 'self' (per cloud) and 'cross' (between the clouds of the pair) layers.
    The 'cross_cat' variant (gcn.py:132-170) is not used by any shipped config and raises."""

    def __init__(self, num_head, feature_dim, k, layer_names):
        super().__init__()
        layers = []
        for name in layer_names:
            if name == 'cross':
                layers.append(AttentionalPropagation(feature_dim, num_head))
            elif name == 'self':
                layers.append(SelfAttention(feature_dim, k))
            else:
                raise NotImplementedError(f"GCN layer '{name}' is not exercised by the shipped configs")
        self.layers = nn.ModuleList(layers)
        self.names = list(layer_names)

    def forward(self, coords0, coords1, desc0, desc1):
        """coords [N,3], desc [N,C] per cloud -> updated (desc0, desc1)."""
        for layer, name in zip(self.layers, self.names):
            if name == 'cross':
                desc0 = desc0 + layer(desc0, desc1)
                desc1 = desc1 + layer(desc1, desc0)           # sees the already-updated desc0 (gcn.py:199-200)
            else:
                desc0 = layer(coords0, desc0)
                desc1 = layer(coords1, desc1)
        return desc0, desc1


# ---- padded-batch form (BASELINE config 3 with P collated pairs per call) ------------------------------------------------
# The same layers over B = 2P clouds at once: clouds are padded to the longest one ([B, Nmax, C] + validity mask), every
# reduction that the per-cloud form takes over "all points of the cloud" (k-NN candidates, InstanceNorm statistics,
# attention keys) is masked. One pass of ~150 library ops per call instead of one per pair: with several calls in flight
# from several host threads the per-pair Python loop was the bottleneck of the whole KPFCNN forward (GIL-bound dispatch).
def _masked_stats(t, mask, dims):
    """mean / biased variance of t over `dims`, counting only the positions where mask (broadcastable to t) is set"""
    m = mask.to(t.dtype)
    cnt = m.expand_as(t).sum(dim=dims, keepdim=True).clamp_min(1.0)
    mean = (t * m).sum(dim=dims, keepdim=True) / cnt
    var = (((t - mean) * m) ** 2).sum(dim=dims, keepdim=True) / cnt
    return mean, var


def _edge_conv_padded(x, idx, weight, mask, slope=0.2, eps=1e-5):
    """_edge_conv on padded clouds: x [B,N,C], idx [B,N,k] (local indices), mask [B,N] -> [B,N,Cout]"""
    w = weight.reshape(weight.shape[0], -1)
    c = x.shape[2]
    wa, wb = w[:, :c], w[:, c:]
    centre = x @ (wa - wb).t()
    neigh = x @ wb.t()
    b, n, k = idx.shape
    gathered = torch.gather(neigh, 1, idx.reshape(b, n * k, 1).expand(-1, -1, neigh.shape[2])).view(b, n, k, -1)
    e = centre.unsqueeze(2) + gathered                                   # [B,N,k,Cout]
    mean, var = _masked_stats(e, mask[:, :, None, None], (1, 2))
    e = F.leaky_relu((e - mean) * torch.rsqrt(var + eps), slope)
    return e.max(dim=2)[0]


def self_attention_padded(layer, coords, feats, mask):
    """SelfAttention.forward on padded clouds: coords [B,N,3], feats [B,N,C], mask [B,N]"""
    sq = (coords * coords).sum(dim=2)
    dist = (-2.0 * coords @ coords.transpose(1, 2) + sq[:, :, None] + sq[:, None, :]).clamp_min(1e-12)
    dist = dist.masked_fill(~mask[:, None, :], float("inf"))             # padded points are nobody's neighbour
    idx = dist.topk(layer.k + 1, dim=2, largest=False, sorted=True)[1][:, :, 1:]
    x1 = _edge_conv_padded(feats, idx, layer.conv1.weight, mask)
    x2 = _edge_conv_padded(x1, idx, layer.conv2.weight, mask)
    x3 = torch.cat((feats, x1, x2), dim=2) @ layer.conv3.weight.reshape(layer.conv3.weight.shape[0], -1).t()
    mean, var = _masked_stats(x3, mask[:, :, None], (1,))
    return F.leaky_relu((x3 - mean) * torch.rsqrt(var + 1e-5), 0.2)


def cross_attention_padded(layer, x, source, mask_x, mask_s):
    """AttentionalPropagation.forward on padded clouds: x [P,N,C] attends to source [P,M,C]"""
    att = layer.attn
    def proj(l, t):
        return _conv1d(l, t).view(t.shape[0], t.shape[1], att.dim, att.num_heads).permute(0, 3, 1, 2)   # [P,heads,N,dim]
    q, k, v = proj(att.proj[0], x), proj(att.proj[1], source), proj(att.proj[2], source)
    logits = q @ k.transpose(2, 3) / att.dim ** .5
    prob = torch.softmax(logits.masked_fill(~mask_s[:, None, None, :], float("-inf")), dim=-1)
    msg = (prob @ v).permute(0, 2, 3, 1).reshape(x.shape[0], x.shape[1], att.dim * att.num_heads)
    msg = _conv1d(att.merge, msg)
    h = _conv1d(layer.mlp[0], torch.cat([x, msg], dim=2))
    mean, var = _masked_stats(h, mask_x[:, :, None], (1,))
    return _conv1d(layer.mlp[3], F.relu((h - mean) * torch.rsqrt(var + 1e-5)))


def gcn_padded(gcn, coords, feats, mask):
    """GCN.forward for P pairs at once: coords [2P,N,3], feats [2P,N,C], mask [2P,N]; clouds 2p / 2p+1 are pair p's
    source / target. Returns the updated [2P,N,C] (padding rows hold garbage)."""
    for layer, name in zip(gcn.layers, gcn.names):
        if name == 'cross':
            d0, d1 = feats[0::2], feats[1::2]
            m0, m1 = mask[0::2], mask[1::2]
            d0 = d0 + cross_attention_padded(layer, d0, d1, m0, m1)
            d1 = d1 + cross_attention_padded(layer, d1, d0, m1, m0)           # sees the already-updated d0 (gcn.py:199-200)
            feats = torch.stack([d0, d1], dim=1).reshape(feats.shape)
        else:
            feats = self_attention_padded(layer, coords, feats, mask)
    return feats
