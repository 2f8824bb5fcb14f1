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


def kpconv_ref(q_pts, s_pts, inds, x, kernel_points, weights, extent):
    inds = inds.long()
    s = torch.cat((s_pts, torch.zeros_like(s_pts[:1, :]) + 1e6), 0)           # :269 shadow point
    nb = s[inds, :] - q_pts.unsqueeze(1)                                      # :272-275
    diff = nb.unsqueeze(2) - kernel_points                                    # :285-286 [N,H,K,3]
    sq = torch.sum(diff ** 2, dim=3)                                          # :289
    w = torch.clamp(1 - torch.sqrt(sq) / extent, min=0.0).transpose(1, 2)     # :328-329 [N,K,H]
    xz = torch.cat((x, torch.zeros_like(x[:1, :])), 0)                        # :348 zero feature row
    nx = xz[inds]                                                             # :351 [N,H,Cin]
    wf = torch.matmul(w, nx)                                                  # :354 [N,K,Cin]
    out = torch.matmul(wf.permute(1, 0, 2), weights).sum(dim=0)               # :361-366
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


def unary_ref(x, weight, relu=True):
    y = instnorm_ref(F.linear(x, weight))
    return F.leaky_relu(y, 0.1) if relu else y


def _select(name, layer, batch):
    if 'strided' in name:
        return batch['points'][layer + 1], batch['points'][layer], batch['pools'][layer]
    return batch['points'][layer], batch['points'][layer], batch['neighbors'][layer]


def simple_ref(x, batch, sd, prefix, name, layer, extent):
    q, s, inds = _select(name, layer, batch)
    y = kpconv_ref(q, s, inds, x, sd[prefix + 'KPConv.kernel_points'], sd[prefix + 'KPConv.weights'], extent)
    return F.leaky_relu(instnorm_ref(y), 0.1)


def resnetb_ref(feats, batch, sd, prefix, name, layer, extent):
    q, s, inds = _select(name, layer, batch)
    x = feats
    if prefix + 'unary1.mlp.weight' in sd:                                    # :665 (Identity if in == out/4)
        x = unary_ref(x, sd[prefix + 'unary1.mlp.weight'])
    x = kpconv_ref(q, s, inds, x, sd[prefix + 'KPConv.kernel_points'], sd[prefix + 'KPConv.weights'], extent)
    x = F.leaky_relu(instnorm_ref(x), 0.1)                                    # :669
    x = unary_ref(x, sd[prefix + 'unary2.mlp.weight'], relu=False)            # :672
    sc = max_pool_ref(feats, inds) if 'strided' in name else feats            # :675-678
    if prefix + 'unary_shortcut.mlp.weight' in sd:
        sc = unary_ref(sc, sd[prefix + 'unary_shortcut.mlp.weight'], relu=False)
    return F.leaky_relu(x + sc, 0.1)                                          # :681


def encoder_ref(batch, sd, config, return_all=False):
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
            x = simple_ref(x, batch, sd, prefix, block, layer, extent)
        else:
            x = resnetb_ref(x, batch, sd, prefix, block, layer, extent)
        outs.append(x)
        bi += 1
        if 'pool' in block or 'strided' in block:
            layer += 1
            r *= 2
    return outs if return_all else x
