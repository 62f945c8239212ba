"""Model-argument dictionaries (the ``hypes['model']['args']`` of the reference YAMLs) for the
workloads this repo benchmarks and tests.  Geometry follows
opv2v/opencood/hypes_yaml/opcamera/corpbevt.yaml:49-114 unless overridden.
"""
import copy


def corpbevt_args(image_h=512, image_w=512, bev=256, id_pick=(1, 2, 3), q_win=((16, 16), (16, 16), (32, 32)),
                  feat_win=((8, 8), (8, 8), (16, 16)), max_cav=5, fusion_window=8, num_layers=34,
                  fusion_depth=3, self_attn_window=None, use_mask=True, target='dynamic'):
    final = bev // 8
    return {
        'target': target,
        'max_cav': max_cav,
        'encoder': {'num_layers': num_layers, 'pretrained': False, 'image_width': image_w, 'image_height': image_h,
                    'id_pick': list(id_pick)},
        'compression': 0,
        'decoder': {'input_dim': 128, 'num_layer': 3, 'num_ch_dec': [32, 64, 128]},
        'fax': {
            'dim': [128, 128, 128], 'middle': [2, 2, 2],
            'bev_embedding': {'sigma': 1.0, 'bev_height': bev, 'bev_width': bev, 'h_meters': 100, 'w_meters': 100,
                              'offset': 0.0, 'upsample_scales': [2, 4, 8]},
            'cross_view': {'image_height': image_h, 'image_width': image_w, 'no_image_features': False, 'skip': True,
                           'heads': [4, 4, 4], 'dim_head': [32, 32, 32], 'qkv_bias': True},
            'cross_view_swap': {'rel_pos_emb': False, 'q_win_size': [list(w) for w in q_win],
                                'feat_win_size': [list(w) for w in feat_win],
                                'bev_embedding_flag': [True, False, False]},
            'self_attn': {'dim_head': 32, 'dropout': 0.1, 'window_size': self_attn_window or final},
        },
        'sttf': {'resolution': 0.390625, 'downsample_rate': 8, 'use_roi_mask': use_mask},
        'fax_fusion': {'input_dim': 128, 'mlp_dim': 256, 'agent_size': max_cav, 'window_size': fusion_window,
                       'dim_head': 32, 'drop_out': 0.1, 'depth': fusion_depth, 'mask': use_mask},
        'seg_head_dim': 32,
        'output_class': 2,
    }


def native_opv2v():
    """corpbevt.yaml as shipped: 4 cams x 512x512, BEV 256 (BASELINE configs C3/C4)."""
    return corpbevt_args()


def north_star():
    """BASELINE.json metric shape: 5 agents x 6 cams x 224x480 -> 256x256 BEV.  ResNet34 levels
    1..3 (strides 4/8/16: 56x120, 28x60, 14x30) so the 8x8 / 4x4 / 1 window grids line up
    (SURVEY 8d, C2 geometry)."""
    return corpbevt_args(image_h=224, image_w=480, id_pick=(0, 1, 2),
                         feat_win=((7, 15), (7, 15), (14, 30)))


def small(image=64, bev=64, n_levels=(1, 2, 3)):
    """Tiny CoBEVT for CPU-oracle parity tests: 64x64 images, BEV 64 (x: 32/16/8),
    features 8/4/2 px, one window per stage except stage 0 (2x2 windows)."""
    return corpbevt_args(image_h=image, image_w=image, bev=bev, id_pick=n_levels,
                         q_win=((16, 16), (16, 16), (8, 8)), feat_win=((4, 4), (4, 4), (2, 2)),
                         fusion_window=4, num_layers=18, fusion_depth=1)


def fax_fused_args(cfg):
    """The SinBEVT-only model (fax.yaml / FaxFusedTransformer) from a CoBEVT argument dict."""
    c = copy.deepcopy(cfg)
    return {k: c[k] for k in ('target', 'encoder', 'decoder', 'fax', 'seg_head_dim', 'output_class')}


def nuscenes_cvt(image_h=224, image_w=480, bev=200, q_win=((10, 10), (10, 10), (25, 25)),
                 feat_win=((6, 12), (6, 12), (14, 30))):
    """nuscenes/config/model/cvt_pyramid_axial.yaml:9-58 as constructor kwargs.  Defaults = the reference's
    native 200x200 BEV (56x120 / 28x60 features are zero-padded to 60x120 / 30x60: quirk Q2 is live);
    BASELINE config C2 is `nuscenes_cvt(bev=256, q_win=((16,16),(16,16),(32,32)), feat_win=((7,15),(7,15),(14,30)))`."""
    return {
        'encoder': {
            'dim': [32, 64, 128], 'middle': [2, 2, 2], 'scale': 1.0,
            'cross_view': {'heads': [1, 2, 4], 'dim_head': [32, 32, 32], 'qkv_bias': True, 'skip': True,
                           'no_image_features': False, 'image_height': image_h, 'image_width': image_w},
            'cross_view_swap': {'rel_pos_emb': False, 'q_win_size': [list(w) for w in q_win],
                                'feat_win_size': [list(w) for w in feat_win],
                                'bev_embedding_flag': [True, False, False]},
            'bev_embedding': {'sigma': 1.0, 'bev_height': bev, 'bev_width': bev, 'h_meters': 100.0,
                              'w_meters': 100.0, 'offset': 0.0, 'upsample_scales': [2, 4, 8]},
            'self_attn': {'dim_head': 32, 'dropout': 0.1, 'window_size': 25},
        },
        'decoder_blocks': [128, 128, 64], 'decoder_dim': 128, 'dim_last': 64,
        'outputs': {'bev': [0, 1], 'center': [1, 2]},
    }


def nuscenes_c2():
    return nuscenes_cvt(bev=256, q_win=((16, 16), (16, 16), (32, 32)), feat_win=((7, 15), (7, 15), (14, 30)))
