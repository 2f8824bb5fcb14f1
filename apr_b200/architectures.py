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
