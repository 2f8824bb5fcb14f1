"""TEST INFRASTRUCTURE ONLY — fp32 CPU restatement (plain torch ops) of the KPConv block stack.

Functional re-statement, written from scratch, of /root/reference/Predator_APR/models/blocks.py for the branch the
shipped configs exercise (rigid KPConv, linear influence, sum aggregation, InstanceNorm "batch norm"):
  kpconv_ref        <- KPConv.forward            blocks.py:229-374
  max_pool_ref      <- max_pool                  blocks.py:86-102
  closest_pool_ref  <- closest_pool              blocks.py:71-83
  instnorm_ref      <- BatchNormBlock.forward    blocks.py:459-468 (nn.InstanceNorm1d on [1,C,N], eps 1e-5, biased var)
  unary_ref         <- UnaryBlock.forward        blocks.py:496-510
  simple_ref        <- SimpleBlock.forward       blocks.py:581-593
  resnetb_ref       <- ResnetBottleneckBlock.forward  blocks.py:653-681
  encoder_ref       <- KPFCNN.forward encoder loop    models/architectures.py:149-153
Parameters come from a state_dict with the reference's keys (`encoder_blocks.<i>.KPConv.weights`, ...).
Pinned against the real reference modules (imported from /root/reference in this container) by
oracle/make_golden.py, which also freezes the golden vectors under tests/golden/.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import torch
import torch.nn.functional as F

# The oracle is the fp32 statement of the reference: pin the CPU (oneDNN) backends to IEEE fp32 — a host whose oneDNN
# defaults to a reduced-precision fp32 math mode would otherwise move the yardstick itself by ~1e-4. CUDA settings are
# not touched.
for _mod in ("torch.backends.mkldnn", "torch.backends.mkldnn.matmul", "torch.backends.mkldnn.conv"):
    try:
        _m = eval(_mod)
        if hasattr(_m, "fp32_precision"):
            _m.fp32_precision = "ieee"
    except Exception:
        pass


def _id(t):
    return t


def kpconv_ref(q_pts, s_pts, inds, x, kernel_points, weights, extent, quant=_id):
    """quant (default: identity = the fp32 reference) emulates a reduced-precision OPERAND format: it is applied to the
    weights and to the weighted tile, i.e. to the two operands of the contraction, nothing else. Tests use it to tell
    the drift that a 10-bit-mantissa operand format causes by itself from kernel error."""
    inds = inds.long()
    s = torch.cat((s_pts, torch.zeros_like(s_pts[:1, :]) + 1e6), 0)           # :269 shadow point
    nb = s[inds, :] - q_pts.unsqueeze(1)                                      # :272-275
    diff = nb.unsqueeze(2) - kernel_points                                    # :285-286 [N,H,K,3]
    sq = torch.sum(diff ** 2, dim=3)                                          # :289
    w = torch.clamp(1 - torch.sqrt(sq) / extent, min=0.0).transpose(1, 2)     # :328-329 [N,K,H]
    xz = torch.cat((x, torch.zeros_like(x[:1, :])), 0)                        # :348 zero feature row
    nx = xz[inds]                                                             # :351 [N,H,Cin]
    wf = quant(torch.matmul(w, nx))                                           # :354 [N,K,Cin]
    out = torch.matmul(wf.permute(1, 0, 2), quant(weights)).sum(dim=0)        # :361-366
    nn_ = torch.sum(torch.gt(torch.sum(nx, dim=-1), 0.0), dim=-1)             # :369-370
    nn_ = torch.max(nn_, torch.ones_like(nn_))                                # :371
    return out / nn_.unsqueeze(1)                                             # :372


def max_pool_ref(x, inds):
    xz = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    return torch.max(xz[inds.long()], 1)[0]


def closest_pool_ref(x, inds):
    xz = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    return xz[inds[:, 0].long()]


def instnorm_ref(x, eps=1e-5):
    mean = x.mean(dim=0, keepdim=True)
    var = x.var(dim=0, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def unary_ref(x, weight, relu=True, quant=_id):
    y = instnorm_ref(F.linear(quant(x), quant(weight)))
    return F.leaky_relu(y, 0.1) if relu else y


def _select(name, layer, batch):
    if 'strided' in name:
        return batch['points'][layer + 1], batch['points'][layer], batch['pools'][layer]
    return batch['points'][layer], batch['points'][layer], batch['neighbors'][layer]


def simple_ref(x, batch, sd, prefix, name, layer, extent, quant=_id, quant_in=True):
    q, s, inds = _select(name, layer, batch)
    # (the first block's KPConv runs in fp32 on the product path too: Cin = 1, CUDA cores)
    y = kpconv_ref(q, s, inds, x, sd[prefix + 'KPConv.kernel_points'], sd[prefix + 'KPConv.weights'], extent,
                   quant if quant_in else _id)
    return quant(F.leaky_relu(instnorm_ref(y), 0.1))


def resnetb_ref(feats, batch, sd, prefix, name, layer, extent, quant=_id):
    """quant != identity: every stored activation (= every InstanceNorm + LeakyReLU output) and every contraction operand
    is passed through it — the product path's storage/operand format; statistics and accumulation stay fp32."""
    q, s, inds = _select(name, layer, batch)
    x = feats
    if prefix + 'unary1.mlp.weight' in sd:                                    # :665 (Identity if in == out/4)
        x = quant(unary_ref(x, sd[prefix + 'unary1.mlp.weight'], quant=quant))
    x = kpconv_ref(q, s, inds, x, sd[prefix + 'KPConv.kernel_points'], sd[prefix + 'KPConv.weights'], extent, quant)
    x = quant(F.leaky_relu(instnorm_ref(x), 0.1))                             # :669
    x = unary_ref(x, sd[prefix + 'unary2.mlp.weight'], relu=False, quant=quant)   # :672
    sc = max_pool_ref(feats, inds) if 'strided' in name else feats            # :675-678
    if prefix + 'unary_shortcut.mlp.weight' in sd:
        sc = unary_ref(sc, sd[prefix + 'unary_shortcut.mlp.weight'], relu=False, quant=quant)
    return quant(F.leaky_relu(x + sc, 0.1))                                   # :681


def encoder_ref(batch, sd, config, return_all=False, quant=_id):
    """Runs the encoder blocks (architectures.py:37-71 bookkeeping, :149-153 loop). batch holds CPU tensors."""
    x = batch['features'].clone()
    r = config.first_subsampling_dl * config.conv_radius
    layer, outs, bi = 0, [], 0
    for block in config.architecture:
        if 'upsample' in block:
            break
        extent = r * config.KP_extent / config.conv_radius                   # blocks.py:552 / :609
        prefix = f'encoder_blocks.{bi}.'
        if 'simple' in block:
            x = simple_ref(x, batch, sd, prefix, block, layer, extent, quant, quant_in=False)
        else:
            x = resnetb_ref(x, batch, sd, prefix, block, layer, extent, quant)
        outs.append(x)
        bi += 1
        if 'pool' in block or 'strided' in block:
            layer += 1
            r *= 2
    return outs if return_all else x


# ---- bottleneck GNN + decoder half of KPFCNN (architectures.py:155-212, models/gcn.py) ------------------------------
# Restated in the reference's own formulation ([1,C,N] tensors, explicit edge features), which is deliberately NOT the
# product's factored formulation (apr_b200/gcn.py), so that the comparison checks the algebra as well.
def _sqdist_ref(a, b):
    d = -2 * torch.matmul(a, b.t())                                           # lib/utils.py:89-96
    d = d + torch.sum(a ** 2, dim=-1)[:, None] + torch.sum(b ** 2, dim=-1)[None, :]
    return torch.clamp(d, min=1e-12)


def _graph_feature_ref(coords, feats, k):
    """coords [N,3], feats [C,N] -> [2C,N,k] = cat(x_i, x_j - x_i) over the k nearest other points (gcn.py:9-35)."""
    idx = _sqdist_ref(coords, coords).topk(k=k + 1, dim=-1, largest=False, sorted=True)[1][:, 1:]   # [N,k]
    neigh = feats[:, idx]                                                     # [C,N,k]
    centre = feats.unsqueeze(-1).expand(-1, -1, k)
    return torch.cat((centre, neigh - centre), dim=0)


def _in2d_lrelu_ref(x, slope=0.2, eps=1e-5):
    """InstanceNorm2d (no affine) on [C,N,k] + LeakyReLU."""
    m = x.mean(dim=(1, 2), keepdim=True)
    v = x.var(dim=(1, 2), unbiased=False, keepdim=True)
    return F.leaky_relu((x - m) / torch.sqrt(v + eps), slope)


def self_attention_ref(coords, feats, sd, prefix, k):
    """feats [C,N] -> [C,N] (gcn.py:52-79)"""
    w1 = sd[prefix + 'conv1.weight'].flatten(1); w2 = sd[prefix + 'conv2.weight'].flatten(1); w3 = sd[prefix + 'conv3.weight'].flatten(1)
    x1 = _in2d_lrelu_ref(torch.einsum('oc,cnk->onk', w1, _graph_feature_ref(coords, feats, k))).max(dim=-1)[0]
    x2 = _in2d_lrelu_ref(torch.einsum('oc,cnk->onk', w2, _graph_feature_ref(coords, x1, k))).max(dim=-1)[0]
    x3 = torch.matmul(w3, torch.cat((feats, x1, x2), dim=0)).unsqueeze(-1)
    return _in2d_lrelu_ref(x3).squeeze(-1)


def cross_attention_ref(x, src, sd, prefix, heads):
    """x [C,N], src [C,M] -> update for x [C,N] (gcn.py:92-129)"""
    def conv(name, t):
        return torch.matmul(sd[prefix + name + '.weight'].squeeze(-1), t) + sd[prefix + name + '.bias'][:, None]
    c = x.shape[0]
    dim = c // heads
    q = conv('attn.proj.0', x).view(dim, heads, -1)
    kk = conv('attn.proj.1', src).view(dim, heads, -1)
    v = conv('attn.proj.2', src).view(dim, heads, -1)
    prob = torch.softmax(torch.einsum('dhn,dhm->hnm', q, kk) / dim ** .5, dim=-1)
    msg = conv('attn.merge', torch.einsum('hnm,dhm->dhn', prob, v).reshape(c, -1))
    h = conv('mlp.0', torch.cat([x, msg], dim=0))
    h = (h - h.mean(dim=1, keepdim=True)) / torch.sqrt(h.var(dim=1, unbiased=False, keepdim=True) + 1e-5)   # InstanceNorm1d
    return conv('mlp.3', F.relu(h))


def kpfcnn_ref(batch, sd, config, enc_outs=None):
    """Full KPFCNN.forward (architectures.py:137-212) on CPU tensors -> (feats_f, scores_overlap, scores_saliency).
    enc_outs (optional): the outputs of the encoder blocks to use instead of running the encoder — tests feed the bottleneck
    / GNN / decoder half with the product's own encoder activations to check that half on identical inputs."""
    outs = enc_outs if enc_outs is not None else encoder_ref(batch, sd, config, return_all=True)
    arch = list(config.architecture)
    start = next(i for i, b in enumerate(arch) if 'upsample' in b)
    skips, x_in = [], batch['features']
    for i, b in enumerate(arch[:start]):                                      # :149-152: the INPUT of every strided block
        if any(t in b for t in ('pool', 'strided', 'global')):
            skips.append(x_in)
        x_in = outs[i]
    x = outs[-1]
    n_src = int(batch['stack_lengths'][-1][0])
    pts = batch['points'][-1]
    feats = torch.matmul(sd['bottle.weight'].squeeze(-1), x.t()) + sd['bottle.bias'][:, None]      # [C,N] :157-158
    uncond = feats.t()
    d0, d1 = feats[:, :n_src], feats[:, n_src:]
    for li, name in enumerate(config.nets):                                   # gcn.py:194-209
        p = f'gnn.layers.{li}.'
        if name == 'self':
            d0 = self_attention_ref(pts[:n_src], d0, sd, p, config.dgcnn_k)
            d1 = self_attention_ref(pts[n_src:], d1, sd, p, config.dgcnn_k)
        else:
            d0 = d0 + cross_attention_ref(d0, d1, sd, p, config.num_head)
            d1 = d1 + cross_attention_ref(d1, d0, sd, p, config.num_head)
    feats = torch.cat([d0, d1], dim=1)
    feats = torch.matmul(sd['proj_gnn.weight'].squeeze(-1), feats) + sd['proj_gnn.bias'][:, None]
    scores = (torch.matmul(sd['proj_score.weight'].squeeze(-1), feats) + sd['proj_score.bias'][:, None]).t()   # [N,1]
    raw = feats.t()
    fn = F.normalize(raw, p=2, dim=1)
    inner = fn[:n_src] @ fn[n_src:].t()
    temp = torch.exp(sd['epsilon']) + 0.03
    s1 = torch.softmax(inner / temp, dim=1) @ scores[n_src:]
    s2 = torch.softmax(inner.t() / temp, dim=1) @ scores[:n_src]
    sal = torch.cat((s1, s2), dim=0)
    body = raw if config.condition_feature else uncond
    x = torch.cat([scores, sal, body] if config.add_cross_score else [scores, body], dim=1)
    layer = sum(1 for b in arch[:start] if 'pool' in b or 'strided' in b)
    for i, b in enumerate(arch[start:]):                                      # :195-198
        if i > 0 and 'upsample' in arch[start + i - 1]:
            x = torch.cat([x, skips.pop()], dim=1)
        if 'upsample' in b:
            x = closest_pool_ref(x, batch['upsamples'][layer - 1])
            layer -= 1
        elif b == 'last_unary':
            x = F.linear(x, sd[f'decoder_blocks.{i}.mlp.weight'])
        else:
            x = unary_ref(x, sd[f'decoder_blocks.{i}.mlp.weight'])
    d = config.final_feats_dim
    clean = lambda s: torch.nan_to_num(torch.clamp(torch.sigmoid(s), 0, 1), nan=0.0, posinf=0.0, neginf=0.0)
    return F.normalize(x[:, :d], p=2, dim=1), clean(x[:, d]), clean(x[:, d + 1])
