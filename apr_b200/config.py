"""Model hyper-parameters that fix every shape on the hot path.

Mirrors the `model:` section of Predator_APR/configs/train/kitti.yaml:11-32 (flattened into one namespace like
lib/utils.py:46-65 does) and the architecture lists of configs/models.py:22-60. `AttrDict` stands in for easydict
(main.py:31), which is not installed here.
"""


class AttrDict(dict):
    """dict with attribute access (config.x == config['x'])."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


_KPFCNN_ARCH = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb',
                'resnetb_strided', 'resnetb', 'resnetb', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary',
                'nearest_upsample', 'last_unary']

architectures = {'indoor': list(_KPFCNN_ARCH), 'kitti': list(_KPFCNN_ARCH), 'nuscenes': list(_KPFCNN_ARCH)}


def kitti_config(**overrides):
    """configs/train/kitti.yaml:11-38 (model + overlap_attention_module) with architecture = architectures['kitti']."""
    cfg = AttrDict(
        dataset='kitti', num_layers=4, in_points_dim=3, first_feats_dim=256, final_feats_dim=32,
        first_subsampling_dl=0.3, in_feats_dim=1, conv_radius=4.25, deform_radius=5.0, num_kernel_points=15,
        KP_extent=2.0, KP_influence='linear', aggregation_mode='sum', fixed_kernel_points='center',
        use_batch_norm=True, batch_norm_momentum=0.02, deformable=False, modulated=False, add_cross_score=True,
        condition_feature=True, model='KPFCNN', gnn_feats_dim=256, dgcnn_k=10, num_head=4,
        nets=['self', 'cross', 'self'], switch_to_decoder=False, symmetric=False, point_generation_ratio=4,
        architecture=list(architectures['kitti']))
    cfg.update(overrides)
    return cfg


def nuscenes_config(**overrides):
    """configs/train/nuscenes.yaml model section: identical to KITTI's on every hot-path key."""
    cfg = kitti_config(dataset='nuscenes', architecture=list(architectures['nuscenes']))
    cfg.update(overrides)
    return cfg
