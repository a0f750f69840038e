"""Synthetic inputs and seeded synthetic weights (no datasets / checkpoints are reachable offline).

``param_spec`` enumerates the reference checkpoint's keys and shapes for a model config (SURVEY.md
Appendix B; key names are the attribute paths of ``FocalFormer3D`` and its sub-modules, e.g.
``pts_bbox_head.heatmap_head.0.conv.weight`` from ``focal_decoder.py:202-221``).  The same spec validates
real checkpoints at load time and drives ``make_state_dict`` for benchmarks and parity tests.
"""
import math
import zlib
import numpy as np
import torch


def _bn(spec, name, c):
    spec[name + ".weight"] = ((c,), "bn_w")
    spec[name + ".bias"] = ((c,), "bn_b")
    spec[name + ".running_mean"] = ((c,), "bn_m")
    spec[name + ".running_var"] = ((c,), "bn_v")
    spec[name + ".num_batches_tracked"] = ((), "count")


def _inverted_residual(spec, name, cin, cout, expand):
    hid = cin * expand
    i = 0
    if expand != 1:
        spec[f"{name}.conv.0.0.weight"] = ((hid, cin, 1, 1), ("w", cin))
        _bn(spec, f"{name}.conv.0.1", hid)
        i = 1
    spec[f"{name}.conv.{i}.0.weight"] = ((hid, 1, 3, 3), ("w", 9))
    _bn(spec, f"{name}.conv.{i}.1", hid)
    spec[f"{name}.conv.{i + 1}.weight"] = ((cout, hid, 1, 1), ("w", hid))
    _bn(spec, f"{name}.conv.{i + 2}", cout)


def _resnet50_spec(s, p):
    """[upstream] mmdet ResNet-50 (torchvision key names)."""
    s[f"{p}.conv1.weight"] = ((64, 3, 7, 7), ("w_img", 3 * 49))
    _bn(s, f"{p}.bn1", 64)
    inpl = 64
    for li, (planes, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for b in range(blocks):
            q = f"{p}.layer{li}.{b}"
            s[f"{q}.conv1.weight"] = ((planes, inpl, 1, 1), ("w", inpl))
            _bn(s, f"{q}.bn1", planes)
            s[f"{q}.conv2.weight"] = ((planes, planes, 3, 3), ("w", planes * 9))
            _bn(s, f"{q}.bn2", planes)
            s[f"{q}.conv3.weight"] = ((planes * 4, planes, 1, 1), ("w_res", planes))
            _bn(s, f"{q}.bn3", planes * 4)
            if b == 0:
                s[f"{q}.downsample.0.weight"] = ((planes * 4, inpl, 1, 1), ("w", inpl))
                _bn(s, f"{q}.downsample.1", planes * 4)
            inpl = planes * 4


def camera_param_spec(model_cfg, s):
    """Image branch + Lift-Splat-Shoot tensors (camera configs: focalformer3d.py:133-153, lss.py:149-210)."""
    if model_cfg.get("img_backbone"):
        assert model_cfg["img_backbone"]["depth"] == 50
        _resnet50_spec(s, "img_backbone")
    nk = model_cfg.get("img_neck")
    if nk:
        oc = nk["out_channels"]
        for i, ci in enumerate(nk["in_channels"]):
            s[f"img_neck.lateral_convs.{i}.conv.weight"] = ((oc, ci, 1, 1), ("w", ci))
            s[f"img_neck.lateral_convs.{i}.conv.bias"] = ((oc,), "b")
            s[f"img_neck.fpn_convs.{i}.conv.weight"] = ((oc, oc, 3, 3), ("w", oc * 9))
            s[f"img_neck.fpn_convs.{i}.conv.bias"] = ((oc,), "b")
    ne = model_cfg["imgpts_neck"]
    if ne.get("cam_lss"):
        H, W = ne["img_scale"]
        fH, fW = H // 4, W // 4
        D, camC, hc = 41, 64, ne["hidden_channel"]
        r = ne["pc_range"]
        cz = int(camC * ((r[5] - r[2]) // 0.6))
        q = "imgpts_neck.cam_lss"
        s[f"{q}.frustum"] = ((D, fH, fW, 3), ("frustum", (H, W, fH, fW)))
        s[f"{q}.camencode.depthnet.weight"] = ((D + camC, 256, 1, 1), ("w", 256))
        s[f"{q}.camencode.depthnet.bias"] = ((D + camC,), "b")
        for i, (ci, co) in enumerate(((cz, cz), (cz, 512), (512, 512), (512, hc))):
            s[f"{q}.bevencode.{3 * i}.weight"] = ((co, ci, 3, 3), ("w_bev" if i == 0 else "w", ci * 9))
            _bn(s, f"{q}.bevencode.{3 * i + 1}", co)


def param_spec(model_cfg):
    """name -> (shape, init kind) for every tensor of the reference state dict."""
    s = {}
    camera_param_spec(model_cfg, s)
    if not model_cfg.get("input_pts", True):          # camera-only (DeformFormer3D_C_R50): no LiDAR tower
        return _head_spec(model_cfg, s)
    ve = model_cfg["pts_voxel_encoder"]
    me = model_cfg["pts_middle_encoder"]
    if ve["type"] == "HardVFE":
        fc = ve["feat_channels"][0]
        s["pts_voxel_encoder.vfe_layers.0.linear.weight"] = ((fc, ve["in_channels"]), ("w_in", ve["in_channels"]))
        _bn(s, "pts_voxel_encoder.vfe_layers.0.norm", fc)
    # --- SparseEncoder (mmdet3d v0.17.1; spconv-v1 weight layout [kD,kH,kW,Cin,Cout])
    p = "pts_middle_encoder"
    base = me.get("base_channels", 16)
    taps = (3.0, 10.0, 16.0, 21.0)        # measured mean active taps per SubM level on LiDAR-like clouds
    # raw point statistics reach conv_input only with the mean VFE; HardVFE hands it O(1) learned features
    s[f"{p}.conv_input.0.weight"] = ((3, 3, 3, me["in_channels"], base),
                                     ("spw_in" if ve["type"] in ("HardSimpleVFE", "DynamicSimpleVFE") else "spw", me["in_channels"] * taps[0]))
    _bn(s, f"{p}.conv_input.1", base)
    cin = base
    enc = me["encoder_channels"]
    for i, blocks in enumerate(enc):
        for j, cout in enumerate(blocks):
            q = f"{p}.encoder_layers.encoder_layer{i + 1}.{j}"
            if j == len(blocks) - 1 and i != len(enc) - 1:
                s[f"{q}.0.weight"] = ((3, 3, 3, cin, cout), ("spw", cin * taps[min(i, 3)] * 0.6))
                _bn(s, f"{q}.1", cout)
            else:
                for n in (1, 2):
                    s[f"{q}.conv{n}.weight"] = ((3, 3, 3, cout, cout), ("spw", cout * taps[min(i, 3)] * n))
                    _bn(s, f"{q}.bn{n}", cout)
            cin = cout
    oc = me.get("output_channels", 128)
    s[f"{p}.conv_out.0.weight"] = ((3, 1, 1, cin, oc), ("w", cin * 2))
    _bn(s, f"{p}.conv_out.1", oc)
    # --- SECOND
    bb = model_cfg["pts_backbone"]
    inf = [bb["in_channels"], *bb["out_channels"][:-1]]
    for i, n in enumerate(bb["layer_nums"]):
        c = bb["out_channels"][i]
        s[f"pts_backbone.blocks.{i}.0.weight"] = ((c, inf[i], 3, 3), ("w", inf[i] * 9))
        _bn(s, f"pts_backbone.blocks.{i}.1", c)
        for l in range(n):
            s[f"pts_backbone.blocks.{i}.{3 * (l + 1)}.weight"] = ((c, c, 3, 3), ("w", c * 9))
            _bn(s, f"pts_backbone.blocks.{i}.{3 * (l + 1) + 1}", c)
    # --- SECONDFPN
    nk = model_cfg["pts_neck"]
    for i, c in enumerate(nk["out_channels"]):
        st = nk["upsample_strides"][i]
        ci = nk["in_channels"][i]
        if st > 1 or not nk.get("use_conv_for_no_stride", False):
            s[f"pts_neck.deblocks.{i}.0.weight"] = ((ci, c, st, st), ("w", ci))       # ConvTranspose2d [Cin,Cout,k,k]
        else:
            s[f"pts_neck.deblocks.{i}.0.weight"] = ((c, ci, 1, 1), ("w", ci))
        _bn(s, f"pts_neck.deblocks.{i}.1", c)
    return _head_spec(model_cfg, s)


def _head_spec(model_cfg, s):
    # --- FocalEncoder
    ne = model_cfg["imgpts_neck"]
    hc = ne["hidden_channel"]
    if not ne.get("input_pts", True):
        return _decoder_spec(model_cfg, s)
    s["imgpts_neck.shared_conv_pts.weight"] = ((hc, ne["in_channels_pts"], 3, 3), ("w", ne["in_channels_pts"] * 9))
    s["imgpts_neck.shared_conv_pts.bias"] = ((hc,), "b")
    fused = bool(model_cfg.get("input_img", False))               # LiDAR + camera (FocalFormer3D_LC): 'bevfusion' blocks
    if fused and not ne.get("cam_lss"):
        # 'proj' variant (FocalFormer3D_LC_Proj): image-plane 3x3 conv, then the layer-0 I2P projection block
        # (focal_encoder.py:134-141,28-31; encoder_utils.py:184-193: one single-head nn.MultiheadAttention)
        ci = ne["in_channels_img"]
        s["imgpts_neck.shared_conv_img.weight"] = ((hc, ci, 3, 3), ("w", ci * 9))
        s["imgpts_neck.shared_conv_img.bias"] = ((hc,), "b")
        if ne.get("num_layers") and not ne.get("iterbev_wo_img", False):
            q0 = "imgpts_neck.fusion_blocks.0.I2P_block.learnedAlign"
            s[f"{q0}.in_proj_weight"] = ((3 * hc, hc), ("w_att", hc))
            s[f"{q0}.in_proj_bias"] = ((3 * hc,), "b")
            s[f"{q0}.out_proj.weight"] = ((hc, hc), ("w", hc))
            s[f"{q0}.out_proj.bias"] = ((hc,), "b")
    for i in range(ne["num_layers"] or 0):
        q = f"imgpts_neck.fusion_blocks.{i}"
        if ne.get("iterbev", "bevfusion") == "bevfusionmb2":
            _inverted_residual(s, f"{q}.P_IML", hc, hc, 2)
            _inverted_residual(s, f"{q}.P_out_proj", 2 * hc, hc, 1)
            _inverted_residual(s, f"{q}.P_integration", 2 * hc, hc, 1)
        else:
            # focal_encoder.py:40-43: LocalContextAttentionBlock(k=9) + two 1x1 ConvBN (encoder_utils.py:109-163)
            for proj in ("query_project.0", "query_project.1", "key_project.0", "key_project.1", "value_project"):
                # query / key gains make q.k / sqrt(C) spread over a few units: a peaked, not a flat, 81-way soft-max
                s[f"{q}.P_IML.{proj}.conv.weight"] = ((hc, hc, 1, 1), ("w" if proj == "value_project" else "w_att", hc))
                _bn(s, f"{q}.P_IML.{proj}.bn", hc)
            for nm in ("P_out_proj", "P_integration"):
                s[f"{q}.{nm}.conv.weight"] = ((hc, 2 * hc, 1, 1), ("w", 2 * hc))
                _bn(s, f"{q}.{nm}.bn", hc)
        if fused and not ne.get("iterbev_wo_img", False):
            for n in (1, 2):                                      # iterimg_conv = resnet.BasicBlock (focal_encoder.py:50-52)
                s[f"{q}.iterimg_conv.0.conv{n}.weight"] = ((hc, hc, 3, 3), ("w" if n == 1 else "w_res", hc * 9))
                _bn(s, f"{q}.iterimg_conv.0.bn{n}", hc)
    if ne.get("extra_feat"):
        s["imgpts_neck.extra_output.conv.weight"] = ((hc, hc, 3, 3), ("w", hc * 9))
        _bn(s, "imgpts_neck.extra_output.bn", hc)
    return _decoder_spec(model_cfg, s)


def _decoder_spec(model_cfg, s):
    # --- FocalDecoder
    hd = model_cfg["pts_bbox_head"]
    h = "pts_bbox_head"
    hc = hd["hidden_channel"]
    nc = hd["num_classes"]
    if hd.get("multiscale"):
        for n in ("dconv", "dconv2"):
            s[f"{h}.{n}.conv.weight"] = ((hc, hc, 3, 3), ("w", hc * 9))
            _bn(s, f"{h}.{n}.bn", hc)
    if hd.get("roi_feats"):
        pre = hd["roi_feats"] ** 2 * hc * (3 if hd.get("multiscale") else 1)
        hr = hd.get("hidden_channel_roi", 512)
        step = 4 if hd.get("roi_dropout_rate", 0.0) > 1e-4 else 3
        for i, chl in enumerate((hr, hr, hc)):
            s[f"{h}.roi_mlp.{i * step}.weight"] = ((chl, pre), ("w", pre))
            _bn(s, f"{h}.roi_mlp.{i * step + 1}", chl)
            pre = chl

    def heat(name):
        s[f"{name}.0.conv.weight"] = ((hc, hc, 3, 3), ("w", hc * 9))
        _bn(s, f"{name}.0.bn", hc)
        s[f"{name}.1.weight"] = ((nc, hc, 3, 3), ("w", hc * 9))
        s[f"{name}.1.bias"] = ((nc,), "heat_b")
    heat(f"{h}.heatmap_head")
    stages = (hd.get("multistage_heatmap") or 0) + (1 if hd.get("reuse_first_heatmap") else 0)
    if hd.get("input_img", True) or hd.get("iterbev_wo_img"):
        if stages:
            for i in range(stages):
                if not (i == 0 and hd.get("reuse_first_heatmap")):
                    heat(f"{h}.heatmap_head_img.{i}")
        else:
            heat(f"{h}.heatmap_head_img")
    s[f"{h}.class_encoding.weight"] = ((hc, nc, 1), ("w", nc))
    s[f"{h}.class_encoding.bias"] = ((hc,), "b")
    dc = hd["decoder_cfg"]
    tl = dc["transformerlayers"]
    ff = tl["feedforward_channels"]
    msda = tl["attn_cfgs"][1]
    nh, nl, npt = msda["num_heads"], msda["num_levels"], msda["num_points"]
    for i in range(hd["num_decoder_layers"]):
        s[f"{h}.pos_embed_learned.{i}.layers.0.weight"] = ((hc, 256), ("w", 256))
        s[f"{h}.pos_embed_learned.{i}.layers.0.bias"] = ((hc,), "b")
        s[f"{h}.pos_embed_learned.{i}.layers.1.weight"] = ((hc, hc), ("w", hc))
        s[f"{h}.pos_embed_learned.{i}.layers.1.bias"] = ((hc,), "b")
        for j in range(dc["num_layers"]):
            q = f"{h}.decoder.{i}.layers.{j}"
            s[f"{q}.attentions.0.attn.in_proj_weight"] = ((3 * hc, hc), ("w", hc))
            s[f"{q}.attentions.0.attn.in_proj_bias"] = ((3 * hc,), "b")
            s[f"{q}.attentions.0.attn.out_proj.weight"] = ((hc, hc), ("w", hc))
            s[f"{q}.attentions.0.attn.out_proj.bias"] = ((hc,), "b")
            s[f"{q}.attentions.1.sampling_offsets.weight"] = ((nh * nl * npt * 2, hc), ("w_small", hc))
            s[f"{q}.attentions.1.sampling_offsets.bias"] = ((nh * nl * npt * 2,), "offs_b")
            s[f"{q}.attentions.1.attention_weights.weight"] = ((nh * nl * npt, hc), ("w", hc))
            s[f"{q}.attentions.1.attention_weights.bias"] = ((nh * nl * npt,), "b")
            s[f"{q}.attentions.1.value_proj.weight"] = ((hc, hc), ("w", hc))
            s[f"{q}.attentions.1.value_proj.bias"] = ((hc,), "b")
            s[f"{q}.attentions.1.output_proj.weight"] = ((hc, hc), ("w", hc))
            s[f"{q}.attentions.1.output_proj.bias"] = ((hc,), "b")
            s[f"{q}.ffns.0.layers.0.0.weight"] = ((ff, hc), ("w", hc))
            s[f"{q}.ffns.0.layers.0.0.bias"] = ((ff,), "b")
            s[f"{q}.ffns.0.layers.1.weight"] = ((hc, ff), ("w", ff))
            s[f"{q}.ffns.0.layers.1.bias"] = ((hc,), "b")
            for n in range(3):
                s[f"{q}.norms.{n}.weight"] = ((hc,), "ln_w")
                s[f"{q}.norms.{n}.bias"] = ((hc,), "b")
        heads = dict(hd["common_heads"])
        if hd.get("classaware_reg"):                      # focal_decoder.py:317-319: one regression set per class
            heads = {k_: (v_[0] * nc, v_[1]) for k_, v_ in heads.items()}
        heads["heatmap"] = (nc, hd.get("num_heatmap_convs", 2))
        for name, (k, nconv) in heads.items():
            assert nconv == 2
            s[f"{h}.prediction_heads.{i}.{name}.0.conv.weight"] = ((64, hc, 1), ("w", hc))
            _bn(s, f"{h}.prediction_heads.{i}.{name}.0.bn", 64)
            s[f"{h}.prediction_heads.{i}.{name}.1.weight"] = ((k, 64, 1), ("w_small", 64))
            s[f"{h}.prediction_heads.{i}.{name}.1.bias"] = ((k,), "heat_b" if name == "heatmap" else "b")
    return s


def _gen(name, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def make_state_dict(model_cfg, seed=0):
    """Deterministic synthetic weights (per-key seeded, fp32, CPU).  Gains are chosen so activations stay
    O(1) through the ~40 stacked layers and the heatmaps have well separated peaks."""
    sd = {}
    for name, (shape, kind) in param_spec(model_cfg).items():
        g = _gen(name, seed)
        if kind == "count":
            sd[name] = torch.zeros((), dtype=torch.long)
            continue
        if isinstance(kind, tuple) and kind[0] == "frustum":
            H, W, fH, fW = kind[1]                     # lss.py:215-226 (not random: it is geometry)
            ds = torch.arange(4.0, 45.0, 1.0).view(-1, 1, 1).expand(-1, fH, fW)
            D = ds.shape[0]
            xs = torch.linspace(0, W - 1, fW).view(1, 1, fW).expand(D, fH, fW)
            ys = torch.linspace(0, H - 1, fH).view(1, fH, 1).expand(D, fH, fW)
            sd[name] = torch.stack((xs, ys, ds), -1).contiguous()
            continue
        if isinstance(kind, tuple):
            k, fan = kind
            gain = {"w": 1.0, "w_small": 0.5, "spw": 1.2, "spw_in": 1.2, "w_in": 1.0, "w_img": 1.0, "w_res": 0.3,
                    "w_bev": 4.0, "w_att": 6.0}[k]
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan))
            if k == "w_in":              # HardVFE linear [C, F] on raw point features
                t[:, :3] /= 20.0
                if shape[1] > 3:
                    t[:, 3] /= 128.0
            if k == "spw_in":            # raw inputs: metres (|x|~30), intensity 0..255, dt
                t[..., :3, :] /= 20.0
                if shape[3] > 3:
                    t[..., 3, :] /= 128.0
        elif kind == "bn_w" or kind == "ln_w":
            t = 0.8 + 0.4 * torch.rand(shape, generator=g)
        elif kind == "bn_v":
            t = 0.7 + 0.6 * torch.rand(shape, generator=g)
        elif kind in ("bn_b", "bn_m", "b"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "heat_b":
            t = torch.full(shape, -2.19) + 0.05 * torch.randn(shape, generator=g)
        elif kind == "offs_b":
            t = 1.5 * torch.randn(shape, generator=g)
        else:
            raise KeyError(kind)
        sd[name] = t.float()
    return sd


def synth_points(n_points, pc_range, seed=0, n_sweeps=10, n_beams=32, n_features=5):
    """Seeded nuScenes-shaped multi-sweep cloud: one base sweep (log-uniform ranges over the
    beams, uniform azimuth, ground plane, box-like clusters) replicated n_sweeps times with centimetre jitter --
    the sweeps of a mostly static scene overlap, which is what gives ~2-3 points per occupied voxel;
    intensity U(0,255); dt = sweep * 0.05 s; shuffled; range-filtered like the reference's PointsRangeFilter."""
    rng = np.random.default_rng(seed)
    r = np.asarray(pc_range, np.float32)
    half = float(min(r[3], r[4]))
    n_base = max(n_points // n_sweeps, 1)
    n_obj = n_base // 5
    n_bg = n_base - n_obj
    rad = np.exp(rng.uniform(np.log(1.0), np.log(half * 1.3), n_bg))
    az = rng.uniform(-np.pi, np.pi, n_bg)
    beam = rng.integers(0, n_beams, n_bg)
    elev = np.deg2rad(-30.0 + 40.0 * beam / max(n_beams - 1, 1))
    x, y = rad * np.cos(az), rad * np.sin(az)
    z = np.maximum(rad * np.tan(elev), r[2] + 0.3 + 0.02 * rng.standard_normal(n_bg))
    z = np.minimum(z, r[5] - 0.05)
    n_boxes = 40
    centers = rng.uniform(-half * 0.9, half * 0.9, (n_boxes, 2))
    which = rng.integers(0, n_boxes, n_obj)
    ox = centers[which, 0] + rng.normal(0, 0.8, n_obj)
    oy = centers[which, 1] + rng.normal(0, 0.4, n_obj)
    oz = r[2] + 0.5 + np.abs(rng.normal(0, 0.6, n_obj))
    base = np.stack([np.concatenate([x, ox]), np.concatenate([y, oy]), np.concatenate([z, oz])], 1)
    reps = -(-n_points // n_base)
    xyz = np.concatenate([base + rng.normal(0, 0.03, base.shape) for _ in range(reps)])[:n_points]
    sweep = np.repeat(np.arange(reps), n_base)[:n_points]
    pts = np.zeros((n_points, n_features), np.float32)
    pts[:, :3] = xyz
    if n_features > 3:
        pts[:, 3] = rng.uniform(0, 255, n_points)
    if n_features > 4:
        pts[:, 4] = (sweep % n_sweeps) * 0.05
    pts = pts[rng.permutation(n_points)]
    m = ((pts[:, 0] > r[0]) & (pts[:, 0] < r[3]) & (pts[:, 1] > r[1]) & (pts[:, 1] < r[4])
         & (pts[:, 2] > r[2]) & (pts[:, 2] < r[5]))
    return np.ascontiguousarray(pts[m])


def synth_cameras(n_cams=6, img_hw=(448, 800), seed=0):
    """Seeded pinhole rig: n cameras at yaw k*360/n, focal ~0.79 * W, 1.5 m up, looking outwards.
    Returns lidar2img [n, 4, 4] float32 (pixel = K [R|t] X_lidar, homogeneous, z forward)."""
    rng = np.random.default_rng(seed)
    H, W = img_hw
    f = 0.79 * W
    K = np.array([[f, 0, W / 2.0, 0], [0, f, H / 2.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    mats = []
    for k in range(n_cams):
        yaw = 2 * np.pi * k / n_cams + rng.normal(0, 0.01)
        # camera axes in the lidar frame: z forward (cos yaw, sin yaw, 0), x right, y down
        fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
        right = np.array([np.sin(yaw), -np.cos(yaw), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd])                     # lidar -> camera rotation
        t = -R @ np.array([0.3 * np.cos(yaw), 0.3 * np.sin(yaw), 1.5])
        E = np.eye(4)
        E[:3, :3], E[:3, 3] = R, t
        mats.append(K @ E)
    return np.stack(mats).astype(np.float32)
