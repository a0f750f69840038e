# DeformFormer3D-C-R50 (camera-only nuScenes, 6 x 448 x 800 images) in the reference's mmcv config format.
# Hyper-parameters follow projects/configs/focalformer3d/DeformFormer3D_C_R50.py (model / test_cfg only): ResNet-50 +
# FPN image features, Lift-Splat-Shoot camera->BEV (0.6 m cells, 41 depth bins), no LiDAR tower (input_pts=False),
# single-stage averaged heatmap head, 200 proposals, one decoder stage.
plugin = True
plugin_dir = 'projects/mmdet3d_plugin/'

point_cloud_range = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
voxel_size = [0.075, 0.075, 0.2]
out_size_factor = 8
img_scale = (800, 448)
class_names = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier',
               'motorcycle', 'bicycle', 'pedestrian', 'traffic_cone']
hidden = 128

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
    freeze_img=False,
    freeze_pts=True,
    input_img=True,
    input_pts=False,
    img_backbone=dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                      norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch'),
    img_neck=dict(type='FPN', in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5),
    imgpts_neck=dict(type='FocalEncoder', num_layers=None, input_pts=False, cam_lss=True, pc_range=point_cloud_range,
                     img_scale=(img_scale[1], img_scale[0]), in_channels_img=256, in_channels_pts=512,
                     hidden_channel=hidden, bn_momentum=0.1, max_points_height=10, bias='auto', iterbev='bevfusion',
                     iter_bev_cam=True, multistage_heatmap=None, extra_feat=False),
    pts_bbox_head=dict(
        type='FocalDecoder', reuse_first_heatmap=False, extra_feat=False, roi_feats=0, roi_based_reg=False,
        multiscale=True, multistage_heatmap=None, mask_heatmap_mode='poscls', input_img=False, iterbev_wo_img=True,
        bevpos=True, num_proposals=200, hidden_channel=hidden, num_classes=len(class_names), num_decoder_layers=1,
        num_heads=8, initialize_by_heatmap=True, nms_kernel_size=3, bn_momentum=0.1, activation='relu',
        common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)),
        bbox_coder=dict(type='TransFusionBBoxCoder', pc_range=point_cloud_range[:2], voxel_size=voxel_size[:2],
                        out_size_factor=out_size_factor,
                        post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
                        score_threshold=0.0, code_size=10),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2, alpha=0.25, reduction='mean', loss_weight=1.0),
        loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=0.25),
        loss_heatmap=dict(type='GaussianFocalLoss', reduction='mean', loss_weight=1.0),
        decoder_cfg=_decoder),
    test_cfg=dict(pts=dict(dataset='nuScenes', grid_size=[1440, 1440, 40], out_size_factor=out_size_factor,
                           pc_range=point_cloud_range[0:2], voxel_size=voxel_size[:2], nms_type=None)))
