# FocalFormer3D-Waymo15-L: the Waymo LiDAR model with 14x14 ROI grids and class-aware regression heads
# (projects/configs/focalformer3d/FocalFormer3D_Waymo15_L.py; model / test_cfg parts only).
# FocalFormer3D-L for Waymo (LiDAR-only, 64-beam, 0.1 m voxels -> 192x192 BEV) in the reference's mmcv config format.
# Hyper-parameters follow the shipped projects/configs/focalformer3d/FocalFormer3D_Waymo_L.py (model/test_cfg only).
# Deltas vs the nuScenes model: HardVFE (learned 5->64 voxel features, max_num_points 5), 3 classes, 3 HIP stages x 200
# proposals, no velocity head (code_size 8), two FocalEncoder layers, 'Waymo' NMS-exempt classes {1, 2}.
plugin = True
plugin_dir = 'projects/mmdet3d_plugin/'

point_cloud_range = [-76.8, -76.8, -2, 76.8, 76.8, 4]
class_names = ['Car', 'Pedestrian', 'Cyclist']
voxel_size = [0.1, 0.1, 0.15]
out_size_factor = 8
hidden = 128
hip_extra_stages = 2

_decoder = dict(
    type='DeformableDetrTransformerDecoder', num_layers=3, return_intermediate=False,
    transformerlayers=dict(
        type='DetrTransformerDecoderLayer',
        attn_cfgs=[
            dict(type='MultiheadAttention', embed_dims=hidden, num_heads=8, dropout=0.1),
            dict(type='MultiScaleDeformableAttention', embed_dims=hidden, num_levels=3, num_points=4, num_heads=8),
        ],
        feedforward_channels=1024, ffn_dropout=0.1,
        ffn_cfgs=dict(type='FFN', embed_dims=hidden, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True)),
        operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))

model = dict(
    type='FocalFormer3D',
    input_img=False,
    freeze_img=True,
    freeze_pts=True,
    pts_voxel_layer=dict(max_num_points=5, voxel_size=voxel_size, max_voxels=150000,
                         point_cloud_range=point_cloud_range),
    pts_voxel_encoder=dict(type='HardVFE', in_channels=5, feat_channels=[64], with_distance=False,
                           with_cluster_center=False, with_voxel_center=False, voxel_size=voxel_size,
                           norm_cfg=dict(type='BN1d', eps=0.001, momentum=0.01),
                           point_cloud_range=point_cloud_range),
    pts_middle_encoder=dict(
        type='SparseEncoder', in_channels=64, sparse_shape=[41, 1536, 1536], output_channels=128,
        order=('conv', 'norm', 'act'),
        encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
        encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)),
        block_type='basicblock'),
    pts_backbone=dict(type='SECOND', in_channels=256, out_channels=[128, 256], layer_nums=[5, 5],
                      layer_strides=[1, 2], norm_cfg=dict(type='BN', eps=0.001, momentum=0.01),
                      conv_cfg=dict(type='Conv2d', bias=False)),
    pts_neck=dict(type='SECONDFPN', in_channels=[128, 256], out_channels=[256, 256], upsample_strides=[1, 2],
                  norm_cfg=dict(type='BN', eps=0.001, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False),
                  use_conv_for_no_stride=True),
    imgpts_neck=dict(type='FocalEncoder', num_layers=hip_extra_stages, in_channels_img=256, in_channels_pts=512,
                     hidden_channel=hidden, bn_momentum=0.1, max_points_height=10, bias='auto',
                     iterbev='bevfusionmb2', input_img=False, iterbev_wo_img=True,
                     multistage_heatmap=hip_extra_stages, extra_feat=True),
    pts_bbox_head=dict(
        type='FocalDecoder', reuse_first_heatmap=True, extra_feat=True, roi_feats=14, roi_dropout_rate=0.1,
        roi_based_reg=True, roi_expand_ratio=1.2, heatmap_box=False, thin_heatmap_box=False, multiscale=True,
        multistage_heatmap=hip_extra_stages, mask_heatmap_mode='poscls', input_img=False, iterbev_wo_img=True,
        bevpos=True, num_proposals=200, hidden_channel=hidden, num_classes=len(class_names),
        num_decoder_layers=2, num_heads=8, initialize_by_heatmap=True, nms_kernel_size=3, bn_momentum=0.1,
        activation='relu', classaware_reg=True,
        common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2)),
        bbox_coder=dict(type='TransFusionBBoxCoder', pc_range=point_cloud_range[:2], voxel_size=voxel_size[:2],
                        out_size_factor=out_size_factor, post_center_range=[-80, -80, -10.0, 80, 80, 10.0],
                        score_threshold=0.0, code_size=8),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2, alpha=0.25, reduction='mean', loss_weight=1.0),
        loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=2.0),
        loss_heatmap=dict(type='GaussianFocalLoss', reduction='mean', loss_weight=1.0),
        decoder_cfg=_decoder),
    test_cfg=dict(pts=dict(dataset='Waymo', grid_size=[1536, 1536, 40], out_size_factor=out_size_factor,
                           pc_range=point_cloud_range[0:2], voxel_size=voxel_size[:2], nms_type=None)))
