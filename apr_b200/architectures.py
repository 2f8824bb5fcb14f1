"""KFE encoder — the encoder half of KPFCNN (/root/reference/Predator_APR/models/architectures.py:11-71, :137-153)
with the same `encoder_blocks` ModuleList, so `KPFCNN` state_dict keys `encoder_blocks.*` load unchanged.
The bottleneck / GNN / decoder half (:73-129, :155-212) is the 'next' row of the scope table."""
import numpy as np
import torch
import torch.nn as nn

from .blocks import block_decider


class KPFCNNEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        layer = 0
        r = config.first_subsampling_dl * config.conv_radius
        in_dim = config.in_feats_dim
        out_dim = config.first_feats_dim
        self.K = config.num_kernel_points
        self.encoder_blocks = nn.ModuleList()
        self.encoder_skip_dims = []
        self.encoder_skips = []
        for block_i, block in enumerate(config.architecture):
            if ('equivariant' in block) and (not out_dim % 3 == 0):
                raise ValueError('Equivariant block but features dimension is not a factor of 3')
            if np.any([tmp in block for tmp in ['pool', 'strided', 'upsample', 'global']]):
                self.encoder_skips.append(block_i)
                self.encoder_skip_dims.append(in_dim)
            if 'upsample' in block:
                break
            self.encoder_blocks.append(block_decider(block, r, in_dim, out_dim, layer, config))
            in_dim = out_dim // 2 if 'simple' in block else out_dim
            if 'pool' in block or 'strided' in block:
                layer += 1
                r *= 2
                out_dim *= 2
        self.out_dim = in_dim

    @torch.no_grad()
    def forward(self, batch, return_skips=False):
        x = batch['features'].clone().detach()
        skip_x = []
        for block_i, block_op in enumerate(self.encoder_blocks):
            if block_i in self.encoder_skips:
                skip_x.append(x)
            x = block_op(x, batch)
        return (x, skip_x) if return_skips else x


class KPFCNN(KPFCNNEncoder):
    """The full Predator_APR KFE network (architectures.py:9-212): joint encoder -> bottleneck Conv1d -> GCN ->
    overlap / cross-saliency scores -> nearest-upsample + unary decoder. Same attribute names as the reference
    (`encoder_blocks`, `bottle`, `gnn`, `proj_gnn`, `proj_score`, `decoder_blocks`, `epsilon`), so its checkpoints load
    with load_state_dict(strict=True). One collated pair per call (stack_lengths[l] = [len_src, len_tgt])."""

    def __init__(self, config):
        super().__init__(config)
        from .gcn import GCN
        self.final_feats_dim = config.final_feats_dim
        self.epsilon = torch.nn.Parameter(torch.tensor(-5.0))
        self.condition = config.condition_feature
        self.add_cross_overlap = config.add_cross_score
        g = config.gnn_feats_dim
        self.bottle = nn.Conv1d(self.out_dim, g, kernel_size=1, bias=True)                 # :76-77
        self.gnn = GCN(config.num_head, g, config.dgcnn_k, config.nets)                    # :80
        self.proj_gnn = nn.Conv1d(g, g, kernel_size=1, bias=True)
        self.proj_score = nn.Conv1d(g, 1, kernel_size=1, bias=True)
        # decoder bookkeeping (:88-129): walk the architecture from the first upsample block
        arch = list(config.architecture)
        start = next(i for i, b in enumerate(arch) if 'upsample' in b)
        n_down = sum(1 for b in arch[:start] if 'pool' in b or 'strided' in b)
        layer = n_down
        r = config.first_subsampling_dl * config.conv_radius * 2 ** n_down
        in_dim, out_dim = self.out_dim, g + (2 if self.add_cross_overlap else 1)
        self.decoder_blocks = nn.ModuleList()
        self.decoder_concats = []
        for i, block in enumerate(arch[start:]):
            if i > 0 and 'upsample' in arch[start + i - 1]:
                in_dim += self.encoder_skip_dims[layer]
                self.decoder_concats.append(i)
            self.decoder_blocks.append(block_decider(block, r, in_dim, out_dim, layer, config))
            in_dim = out_dim
            if 'upsample' in block:
                layer -= 1
                r *= 0.5
                out_dim = out_dim // 2

    @staticmethod
    def regular_score(score):
        return torch.nan_to_num(score, nan=0.0, posinf=0.0, neginf=0.0)                    # :131-134

    @torch.no_grad()
    def forward(self, batch):
        """batch: the collate dict (points, neighbors, pools, upsamples, stack_lengths, features) on the device.
        Returns (feats_f [N0, final_feats_dim] L2-normalised, scores_overlap [N0], scores_saliency [N0])."""
        import torch.nn.functional as F
        from .gcn import _conv1d
        x, skips = KPFCNNEncoder.forward(self, batch, return_skips=True)                  # 1. joint encoder
        n_src = int(batch['stack_lengths'][-1][0])
        pts_c = batch['points'][-1]
        feats = _conv1d(self.bottle, x)                                                    # 2. bottleneck
        unconditioned = feats
        f0, f1 = self.gnn(pts_c[:n_src], pts_c[n_src:], feats[:n_src], feats[n_src:])      # 3. GNN
        feats = _conv1d(self.proj_gnn, torch.cat([f0, f1], dim=0))
        scores = _conv1d(self.proj_score, feats)                                           # [N,1]
        fn = F.normalize(feats, p=2, dim=1)
        inner = fn[:n_src] @ fn[n_src:].t()                                                # 4. cross saliency
        temperature = torch.exp(self.epsilon) + 0.03
        s1 = torch.softmax(inner / temperature, dim=1) @ scores[n_src:]
        s2 = torch.softmax(inner.t() / temperature, dim=1) @ scores[:n_src]
        saliency = torch.cat((s1, s2), dim=0)
        body = feats if self.condition else unconditioned
        x = torch.cat([scores, saliency, body] if self.add_cross_overlap else [scores, body], dim=1)
        for i, blk in enumerate(self.decoder_blocks):                                      # decoder
            if i in self.decoder_concats:
                x = torch.cat([x, skips.pop()], dim=1)
            x = blk(x, batch)
        d = self.final_feats_dim
        overlap = self.regular_score(torch.sigmoid(x[:, d]).clamp(0, 1))
        sal = self.regular_score(torch.sigmoid(x[:, d + 1]).clamp(0, 1))
        return F.normalize(x[:, :d], p=2, dim=1), overlap, sal
