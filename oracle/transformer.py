"""Oracle: the deformable-DETR decoder stack the head builds.  TEST INFRASTRUCTURE (see oracle/__init__.py).

[upstream] mmcv-full 1.3.18 ``cnn/bricks/transformer.py`` (BaseTransformerLayer, MultiheadAttention, FFN),
``ops/multi_scale_deform_attn.py`` (MultiScaleDeformableAttention + multi_scale_deformable_attn_pytorch) and
mmdet 2.14.0 ``models/utils/transformer.py`` (DeformableDetrTransformerDecoder, DetrTransformerDecoderLayer),
as configured at ``projects/configs/focalformer3d/FocalFormer3D_L.py:285-313`` and called at
``projects/mmdet3d_plugin/models/dense_heads/focal_decoder.py:927-933``.  Eval mode: every dropout is identity.
"""
import math
import torch
from torch import nn
import torch.nn.functional as F


def ms_deform_attn_core(value, spatial_shapes, sampling_locations, attention_weights):
    """mmcv multi_scale_deformable_attn_pytorch: bilinear, zero padding, align_corners=False.

    value [B, sum(HW), heads, d]; spatial_shapes list of (H, W); sampling_locations [B, Nq, heads, L, P, 2]
    in [0, 1] (x, y); attention_weights [B, Nq, heads, L, P].  Returns [B, Nq, heads*d].
    """
    B, _, heads, d = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    vlist = value.split([h * w for h, w in spatial_shapes], dim=1)
    grids = 2 * sampling_locations - 1
    samp = []
    for lvl, (h, w) in enumerate(spatial_shapes):
        v = vlist[lvl].flatten(2).transpose(1, 2).reshape(B * heads, d, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        samp.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(B * heads, 1, Nq, L * P)
    out = (torch.stack(samp, dim=-2).flatten(-2) * aw).sum(-1).view(B, heads * d, Nq)
    return out.transpose(1, 2).contiguous()


def ms_deform_attn_loop(value, spatial_shapes, sampling_locations, attention_weights):
    """Independent formulation (self-check) of ms_deform_attn_core: explicit 4-corner bilinear gather
    (the arithmetic of mmcv's ms_deform_attn_cuda im2col kernel)."""
    B, _, heads, d = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    out = value.new_zeros(B, Nq, heads, d)
    start = 0
    for lvl, (h, w) in enumerate(spatial_shapes):
        v = value[:, start:start + h * w]                             # [B, hw, heads, d]
        start += h * w
        loc = sampling_locations[:, :, :, lvl]                        # [B, Nq, heads, P, 2]
        x = loc[..., 0] * w - 0.5
        y = loc[..., 1] * h - 0.5
        x0, y0 = torch.floor(x), torch.floor(y)
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi = x0 + dx, y0 + dy
                wgt = (1 - (x - xi).abs()) * (1 - (y - yi).abs())
                ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
                idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long()       # [B, Nq, heads, P]
                g = torch.gather(v.permute(0, 2, 1, 3), 2,
                                 idx.permute(0, 2, 1, 3).reshape(B, heads, Nq * P, 1).expand(-1, -1, -1, d))
                g = g.view(B, heads, Nq, P, d).permute(0, 2, 1, 3, 4)
                a = attention_weights[:, :, :, lvl] * wgt * ok
                out += (g * a[..., None]).sum(3)
    return out.reshape(B, Nq, heads * d)


class MultiScaleDeformableAttention(nn.Module):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64, dropout=0.1,
                 batch_first=False, **kw):
        super().__init__()
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.batch_first = batch_first
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kw):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, nq, _ = query.shape
        nv = value.shape[1]
        value = self.value_proj(value).view(bs, nv, self.num_heads, -1)
        off = self.sampling_offsets(query).view(bs, nq, self.num_heads, self.num_levels, self.num_points, 2)
        aw = self.attention_weights(query).view(bs, nq, self.num_heads, self.num_levels * self.num_points)
        aw = aw.softmax(-1).view(bs, nq, self.num_heads, self.num_levels, self.num_points)
        assert reference_points.shape[-1] == 2
        shapes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
        normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).to(off.dtype)
        loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
        out = ms_deform_attn_core(value, shapes, loc, aw)
        out = self.output_proj(out)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return out + identity


class MultiheadAttention(nn.Module):
    """mmcv 1.3.18 wrapper over nn.MultiheadAttention (seq-first)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0, dropout=None, batch_first=False, **kw):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, 0.0)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kw):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=key_padding_mask)[0]
        return identity + out


class FFN(nn.Module):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, **kw):
        super().__init__()
        assert num_fcs == 2
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(0.0)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.0))

    def forward(self, x, identity=None):
        out = self.layers(x)
        return (x if identity is None else identity) + out


class DetrTransformerDecoderLayer(nn.Module):
    """BaseTransformerLayer with operation_order (self_attn, norm, cross_attn, norm, ffn, norm), post-norm."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_cfgs=None, operation_order=None, **kw):
        super().__init__()
        assert tuple(operation_order) == ("self_attn", "norm", "cross_attn", "norm", "ffn", "norm")
        a0, a1 = attn_cfgs
        assert a0["type"] == "MultiheadAttention" and a1["type"] == "MultiScaleDeformableAttention"
        self.embed_dims = a0["embed_dims"]
        self.attentions = nn.ModuleList([
            MultiheadAttention(a0["embed_dims"], a0["num_heads"]),
            MultiScaleDeformableAttention(**{k: v for k, v in a1.items() if k != "type"})])
        self.ffns = nn.ModuleList([FFN(self.embed_dims, feedforward_channels, (ffn_cfgs or {}).get("num_fcs", 2))])
        self.norms = nn.ModuleList([nn.LayerNorm(self.embed_dims) for _ in range(3)])

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                key_padding_mask=None, **kwargs):
        query = self.attentions[0](query, query, query, None, query_pos=query_pos, key_pos=query_pos,
                                   attn_mask=attn_masks)
        query = self.norms[0](query)
        query = self.attentions[1](query, key, value, None, query_pos=query_pos, key_pos=key_pos,
                                   key_padding_mask=key_padding_mask, **kwargs)
        query = self.norms[1](query)
        query = self.ffns[0](query, None)
        query = self.norms[2](query)
        return query


class DeformableDetrTransformerDecoder(nn.Module):
    def __init__(self, num_layers, transformerlayers, return_intermediate=False, **kw):
        super().__init__()
        assert not return_intermediate
        tl = {k: v for k, v in transformerlayers.items() if k != "type"}
        self.layers = nn.ModuleList([DetrTransformerDecoderLayer(**tl) for _ in range(num_layers)])

    def forward(self, query, *args, reference_points=None, valid_ratios=None, **kwargs):
        output = query
        for layer in self.layers:
            assert reference_points.shape[-1] == 2
            ref_in = reference_points[:, :, None] * valid_ratios[:, None]
            output = layer(output, *args, reference_points=ref_in, **kwargs)
        return output, reference_points
