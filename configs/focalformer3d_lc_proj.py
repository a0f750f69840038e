# FocalFormer3D-LC-Proj: the LiDAR + camera model WITHOUT Lift-Splat-Shoot -- image features are lifted to BEV by the
# layer-0 I2P projection block (pillar sample points projected into the cameras + single-head attention), following the
# shipped projects/configs/focalformer3d/FocalFormer3D_LC_Proj.py (model / test_cfg parts only).
# STATUS: oracle + golden only (oracle/bev.py I2P, tests/test_fusion_cpu.py); the CUDA product refuses this config.
plugin = True
plugin_dir = 'projects/mmdet3d_plugin/'

point_cloud_range = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
voxel_size = [0.075, 0.075, 0.2]
out_size_factor = 8
class_names = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier',
               'motorcycle', 'bicycle', 'pedestrian', 'traffic_cone']
hidden = 128
multistage_heatmap = 2
img_scale = (800, 448)

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
    freeze_img=True,
    freeze_pts=True,
    input_img=True,
    img_backbone=dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                      norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch'),
    img_neck=dict(type='FPN', in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5),
    pts_voxel_layer=dict(max_num_points=10, voxel_size=voxel_size, max_voxels=(120000, 160000),
                         point_cloud_range=point_cloud_range),
    pts_voxel_encoder=dict(type='HardSimpleVFE', num_features=5),
    pts_middle_encoder=dict(
        type='SparseEncoder', in_channels=5, sparse_shape=[41, 1440, 1440], output_channels=128,
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
    imgpts_neck=dict(type='FocalEncoder', num_layers=multistage_heatmap, in_channels_img=256, in_channels_pts=512,
                     hidden_channel=hidden, bn_momentum=0.1, max_points_height=10, bias='auto', iterbev='bevfusion',
                     iter_bev_cam=True, multistage_heatmap=multistage_heatmap, extra_feat=True),
    pts_bbox_head=dict(
        type='FocalDecoder', reuse_first_heatmap=False, extra_feat=True, roi_feats=7, roi_dropout_rate=0.1,
        roi_based_reg=True, roi_expand_ratio=1.2, heatmap_box=False, thin_heatmap_box=False, multiscale=True,
        multistage_heatmap=multistage_heatmap, mask_heatmap_mode='poscls', input_img=True, iterbev_wo_img=True,
        bevpos=True, num_proposals=300, hidden_channel=hidden, num_classes=len(class_names),
        num_decoder_layers=2, num_heads=8, initialize_by_heatmap=True, nms_kernel_size=3, bn_momentum=0.1,
        activation='relu',
        common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)),
        bbox_coder=dict(type='TransFusionBBoxCoder', pc_range=point_cloud_range[:2], voxel_size=voxel_size[:2],
                        out_size_factor=out_size_factor,
                        post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
                        score_threshold=0.0, code_size=10),
        # use_sigmoid=True keeps num_classes at 10 (focal_decoder.py:164-166 adds a background class otherwise)
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2, alpha=0.25, reduction='mean', loss_weight=1.0),
        loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=0.25),
        loss_heatmap=dict(type='GaussianFocalLoss', reduction='mean', loss_weight=1.0),
        decoder_cfg=_decoder),
    test_cfg=dict(pts=dict(dataset='nuScenes', grid_size=[1440, 1440, 40], out_size_factor=out_size_factor,
                           pc_range=point_cloud_range[0:2], voxel_size=voxel_size[:2], nms_type=None)))
