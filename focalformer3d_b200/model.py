"""Host-side mirror of the reference's plugin modules for the per-scene forward path.

Registered ``type`` strings (same as the reference registry): ``FocalFormer3D``
(projects/mmdet3d_plugin/models/detectors/focalformer3d.py:26), ``FocalEncoder`` (models/necks/focal_encoder.py:89),
``FocalDecoder`` (models/dense_heads/focal_decoder.py:33), ``TransFusionBBoxCoder``
(core/bbox/coders/transfusion_bbox_coder.py:7) plus stand-ins for the upstream types the shipped configs name
(``HardSimpleVFE``, ``HardVFE``, ``DynamicSimpleVFE``, ``SparseEncoder``, ``SECOND``, ``SECONDFPN``, ``ResNet``, ``FPN``).  Every module is an ``nn.Module`` whose parameter
names reproduce the reference checkpoint keys, so ``load_state_dict`` / ``load_checkpoint`` work unchanged; the
arithmetic runs in libff3d.so (no PyTorch compute fallback: without the library the import of ``.ops`` fails).
"""
import math
import contextlib
import torch
from torch import nn

from .config import DETECTORS, NECKS, HEADS, BBOX_CODERS, VOXEL_ENCODERS, MIDDLE_ENCODERS, BACKBONES
from .synth import param_spec
from . import ops
from .ops import ACT_NONE, ACT_RELU, ACT_RELU6


# ------------------------------------------------------------------------------------------------ param tree
class ParamTree(nn.Module):
    """nn.Module whose (nested) parameter/buffer names are exactly the given spec keys."""

    def build_params(self, spec):
        for name, (shape, kind) in spec.items():
            parts = name.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, ParamTree())
                mod = mod._modules[p]
            leaf = parts[-1]
            if kind == "count":
                mod.register_buffer(leaf, torch.zeros((), dtype=torch.long))
            elif kind in ("bn_m", "bn_v"):
                mod.register_buffer(leaf, torch.zeros(shape) if kind == "bn_m" else torch.ones(shape))
            else:
                mod.register_parameter(leaf, nn.Parameter(torch.zeros(shape), requires_grad=False))

    def flat(self):
        return {k: v.detach() for k, v in self.state_dict().items()}


def sub_spec(spec, prefix):
    n = len(prefix) + 1
    return {k[n:]: v for k, v in spec.items() if k.startswith(prefix + ".")}


# ------------------------------------------------------------------------------------------------ weight packing
def _pad4(n):
    return (n + 3) // 4 * 4


def bn_scale_shift(sd, name, eps):
    s = sd[name + ".weight"].double() / torch.sqrt(sd[name + ".running_var"].double() + eps)
    b = sd[name + ".bias"].double() - sd[name + ".running_mean"].double() * s
    return s, b


def pack_taps(w_tco, dev, bn=None):
    """[taps, cin, cout] (float64/32) -> ops.PackedW: fp32 [taps, pad4(cin), pad4(cout)] for the SIMT kernel plus,
    when the shape is tensor-core tileable, the pre-swizzled hi/lo TF32 images for the tcgen05 kernel."""
    t, ci, co = w_tco.shape
    out = torch.zeros((t, _pad4(ci), _pad4(co)), dtype=torch.float32)
    out[:, :ci, :co] = w_tco.float()
    return ops.PackedW(out.contiguous(), dev, bn)


def pack_conv2d(w, scale=None, dev="cuda"):
    """nn.Conv2d weight [cout, cin, kh, kw] (optionally BN-scaled per cout) -> [kh*kw, cin, cout]."""
    w = w.double()
    if scale is not None:
        w = w * scale.view(-1, 1, 1, 1)
    co, ci, kh, kw = w.shape
    return pack_taps(w.permute(2, 3, 1, 0).reshape(kh * kw, ci, co), dev)


def pack_linear(w, scale=None, dev="cuda", bn=None):
    """nn.Linear / Conv1d(k=1) weight [out, in(,1)] -> [1, in, out]."""
    w = w.double().reshape(w.shape[0], -1)
    if scale is not None:
        w = w * scale.view(-1, 1)
    return pack_taps(w.t().unsqueeze(0), dev, bn)


def vec(v, dev, n=None):
    v = v.detach().float().cpu()
    if n is not None and v.numel() < n:
        v = torch.cat([v, torch.zeros(n - v.numel())])
    return v.contiguous().to(dev)


# ------------------------------------------------------------------------------------------------ components
@VOXEL_ENCODERS.register_module()
class HardSimpleVFE(ParamTree):
    """[upstream] mmdet3d HardSimpleVFE: the mean is fused into ff3d_voxelize_hard (mean_feats output)."""

    def __init__(self, num_features=4, spec=None, **kw):
        super().__init__()
        self.num_features = num_features


@VOXEL_ENCODERS.register_module()
class DynamicSimpleVFE(ParamTree):
    """[upstream] mmdet3d v0.17.1 DynamicSimpleVFE (DynamicScatter mean over ALL points of a voxel) as configured at
    DeformFormer3D_L_dynamic.py; computed inside the dynamic mode of the voxeliser (fp64 atomic sums)."""

    def __init__(self, voxel_size=None, point_cloud_range=None, spec=None, **kw):
        super().__init__()


@VOXEL_ENCODERS.register_module()
class HardVFE(ParamTree):
    """[upstream] mmdet3d v0.17.1 HardVFE as configured by the Waymo configs (FocalFormer3D_Waymo_L.py:141-152):
    no distance / cluster-centre / voxel-centre features, one VFELayer (Linear no-bias + BN1d + ReLU), max over points."""

    def __init__(self, in_channels=4, feat_channels=(), with_distance=False, with_cluster_center=False,
                 with_voxel_center=False, norm_cfg=None, spec=None, **kw):
        super().__init__()
        if with_distance or with_cluster_center or with_voxel_center or len(feat_channels) != 1:
            raise NotImplementedError("HardVFE: only the shipped single-layer, raw-feature variant is built")
        self.in_channels, self.out_channels = in_channels, feat_channels[0]
        self.eps = (norm_cfg or {}).get("eps", 1e-3)
        self.build_params(spec)
        self.pk = None

    def prepare(self, dev):
        sd = self.flat()
        s, b = bn_scale_shift(sd, "vfe_layers.0.norm", self.eps)
        w = sd["vfe_layers.0.linear.weight"].double() * s.view(-1, 1)            # [C, F]
        self.pk = (w.t().contiguous().float().to(dev), vec(b, dev))

    def forward(self, vox, max_points):
        return ops.vfe_hard(vox, self.pk[0], self.pk[1], self.out_channels, max_points, self.in_channels)


@MIDDLE_ENCODERS.register_module()
class SparseEncoder(ParamTree):
    """[upstream] mmdet3d v0.17.1 SparseEncoder (cfg FocalFormer3D_L.py:198-206) on the gather-GEMM kernel."""

    def __init__(self, in_channels, sparse_shape, output_channels=128, base_channels=16,
                 encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                 encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, (0, 1, 1)), (0, 0)), block_type="basicblock",
                 order=("conv", "norm", "act"), spec=None, **kw):
        super().__init__()
        assert block_type == "basicblock" and tuple(order) == ("conv", "norm", "act")
        self.in_channels, self.sparse_shape, self.output_channels = in_channels, tuple(sparse_shape), output_channels
        self.base_channels, self.encoder_channels, self.encoder_paddings = base_channels, encoder_channels, encoder_paddings
        self.build_params(spec)
        self.pk = None
        # row-capacity growth per strided conv (level-2 sites measured at 1.25-2.4x level 1; overflow raises)
        self.cap_growth = (3.0, 1.0, 1.0, 1.0)
        self._side, self._plan_keepalive = None, None

    def prepare(self, dev):
        sd = self.flat()
        eps = 1e-3

        def sp(wname, bnname):
            w = sd[wname].double()                           # [kD,kH,kW,cin,cout]
            s, b = bn_scale_shift(sd, bnname, eps)
            w = (w * s.view(1, 1, 1, 1, -1)).reshape(-1, w.shape[3], w.shape[4])
            return pack_taps(w, dev), vec(b, dev)

        pk = {"conv_input": sp("conv_input.0.weight", "conv_input.1")}
        for i, blocks in enumerate(self.encoder_channels):
            for j in range(len(blocks)):
                q = f"encoder_layers.encoder_layer{i + 1}.{j}"
                if j == len(blocks) - 1 and i != len(self.encoder_channels) - 1:
                    pk[q] = sp(f"{q}.0.weight", f"{q}.1")
                else:
                    pk[q + ".1"] = sp(f"{q}.conv1.weight", f"{q}.bn1")
                    pk[q + ".2"] = sp(f"{q}.conv2.weight", f"{q}.bn2")
        pk["conv_out"] = sp("conv_out.0.weight", "conv_out.1")
        self.pk = pk

    def _plan(self, vox, batch, overflow, bev_shape, streams):
        """Rulebooks of all 21 convs (hash tables, mask-ordered levels, neighbour maps, tile masks): coordinates only, no
        features.  Three phases: (A) the site sets of every level (each needs only its input level's coordinates);
        (B) per level, independently: tap-mask sort + SubM map; (C) per strided conv: its rulebook, once both of its levels
        are in final order.  With ``streams`` (one per level) the B / C chains of the levels run concurrently -- they are
        latency-bound small kernels -- and every rulebook gets an event the gather-GEMMs wait on.
        Returns dict(perm, levels=[(level, subm rulebook, down rulebook into it)], out=..., ready=[event per level], ...)."""
        main = torch.cuda.current_stream()
        ctx = (lambda k: torch.cuda.stream(streams[k % len(streams)])) if streams else (lambda k: contextlib.nullcontext())

        def event(k):
            if not streams:
                return None
            e = torch.cuda.Event()
            e.record(streams[k % len(streams)])
            return e

        def wait(k, e):
            if streams and e is not None:
                streams[k % len(streams)].wait_event(e)
        if streams:
            streams[0].wait_stream(main)
        n_lv = len(self.encoder_channels)
        Hb, Wb, ld = bev_shape
        Cc = self.output_channels
        with ctx(0):                                                     # ---- A: site sets
            cap1 = vox["coors"].shape[0]
            lv = ops.SparseLevel(vox["coors"], vox["n_dev"][:1], cap1, batch, self.sparse_shape)
            lv.build_hash()
            lvls, geo = [lv], []
            for i, blocks in enumerate(self.encoder_channels[:-1]):
                pad = self.encoder_paddings[i][len(blocks) - 1]
                p3 = tuple(pad) if isinstance(pad, (list, tuple)) else (pad,) * 3
                geo.append(((3, 3, 3), (2, 2, 2), p3, blocks[-1]))
                lvls.append(lvls[-1].create_down_level((3, 3, 3), (2, 2, 2), p3, int(lvls[-1].cap * self.cap_growth[i]), overflow))
            out_lvl = lvls[-1].create_down_level((3, 1, 1), (2, 1, 1), (0, 0, 0), lvls[-1].cap, overflow)
            assert ld == out_lvl.shape[0] * Cc and (Hb, Wb) == tuple(out_lvl.shape[1:])
            ev_sites = event(0)
        perm, subm, ev_sub = None, [None] * n_lv, [None] * n_lv
        for k in range(n_lv):                                            # ---- B: mask order + SubM map per level
            wait(k, ev_sites)
            with ctx(k):
                pk_ = lvls[k].sort_by_mask()
                if k == 0:
                    perm = pk_
                subm[k] = lvls[k].subm_map()
                ev_sub[k] = event(k)
        down, ready = [None] * n_lv, [None] * (n_lv + 1)
        ready[0] = ev_sub[0]
        for k in range(1, n_lv):                                         # ---- C: strided rulebooks
            wait(k, ev_sub[k - 1])
            with ctx(k):
                k3, s3, p3, ldy = geo[k - 1]
                down[k] = lvls[k - 1].down_rulebook(lvls[k], k3, s3, p3, ldy)
                ready[k] = event(k)
        with ctx(n_lv - 1):
            rb_out = lvls[-1].down_rulebook(out_lvl, (3, 1, 1), (2, 1, 1), (0, 0, 0), ld, bev=(Hb, Wb, Cc))
            ready[n_lv] = event(n_lv - 1)
        return dict(perm=perm, levels=[(lvls[k], subm[k], down[k]) for k in range(n_lv)], out=(out_lvl, rb_out), ready=ready)

    def forward(self, vox, batch, bev_out, overflow):
        """vox: dict from ops.voxelize; bev_out [B,H,W,2*C] zero-initialised NHWC buffer (channel = d*C + c).
        Every level is stored in tap-mask order (ops.SparseLevel.sort_by_mask): the voxeliser's first-appearance order
        (the reference's, checked by the parity tests) stays untouched in ``vox``; row order inside the sparse encoder is
        implementation-defined in spconv too and vanishes in the dense BEV scatter.

        Streams: the rulebooks depend on coordinates only, so their chains (hash inserts, probes, radix sorts -- latency /
        L2 bound small kernels, ~2 ms at bs = 4 when serialised) run on side streams, one per level, ahead of the
        gather-GEMMs of the main stream, which wait per level on an event.  Inside a CUDA-graph capture the side streams
        become parallel branches."""
        dev = vox["mean"].device
        assert bev_out.is_contiguous()
        bev_shape = (bev_out.shape[1], bev_out.shape[2], bev_out.shape[3])
        main = torch.cuda.current_stream()
        streams = None
        if ops.SPARSE_OVERLAP:
            if self._side is None or self._side[0].device != dev:
                self._side = [torch.cuda.Stream(device=dev) for _ in range(len(self.encoder_channels))]
            streams = self._side
        plan = self._plan(vox, batch, overflow, bev_shape, streams)

        def ready(k):
            if streams:
                main.wait_event(plan["ready"][k])
        # everything the side streams allocated stays referenced until the next forward (the caching allocator must not hand
        # those blocks out again while kernels of another stream still read them)
        self._plan_keepalive = plan

        def conv(x, rb, n_dev, wb, cap_out, act, res=None):
            """One sparse conv.  With the TMA / cp.async kernel every level lives in split form (ops.Split: a row of C
            channels = fp16 [hi | lo], written by the producing conv's epilogue, gathered with 16-byte cp.async copies);
            otherwise (FF3D_GEMM=tf32, FF3D_TMA=0) fp32 rows on the register-path kernel."""
            w, b = wb
            cout = w.shape[-1]
            if ops.tma_enabled() and (cout % 64 == 0 or isinstance(x, ops.Split)):
                ys = ops.Split.empty((cap_out,), cout, dev, zero_row=True)
                ops.sparse_conv(x, rb, n_dev, w, b, None, act=act, res=res, out_s=ys)
                return ys
            y = torch.empty((cap_out, cout), dtype=torch.float32, device=dev)
            ops.sparse_conv(x, rb, n_dev, w, b, y, act=act, res=res)
            return y

        fine = ops.SPARSE_MARKS
        if fine:
            ops.mark("sp:plan-issue")
        ready(0)
        if fine:
            ops.mark("sp:wait-level1-maps")
        lvl, subm, _ = plan["levels"][0]
        feats = ops.gather_rows(vox["mean"], plan["perm"], lvl.n_dev, vox["mean"].shape[1])
        if ops.tma_ok(self.pk["conv_input"][0], feats.shape[1], self.pk["conv_input"][0].shape[-1], sparse=True):
            feats = ops.split_rows(feats, n_dev=lvl.n_dev, zero_row=True)
        x = conv(feats, subm, lvl.n_dev, self.pk["conv_input"], lvl.cap, ACT_RELU)
        self.level_sizes = [lvl.n_dev]
        for i, blocks in enumerate(self.encoder_channels):
            for j, cout in enumerate(blocks):
                q = f"encoder_layers.encoder_layer{i + 1}.{j}"
                if j == len(blocks) - 1 and i != len(self.encoder_channels) - 1:
                    if fine:
                        ops.mark(f"sp:level{i + 1}-convs")
                    ready(i + 1)
                    if fine:
                        ops.mark(f"sp:wait-level{i + 2}-maps")
                    lvl, subm, rb = plan["levels"][i + 1]
                    x = conv(x, rb, lvl.n_dev, self.pk[q], lvl.cap, ACT_RELU)
                    self.level_sizes.append(lvl.n_dev)
                else:
                    t = conv(x, subm, lvl.n_dev, self.pk[q + ".1"], lvl.cap, ACT_RELU)
                    x = conv(t, subm, lvl.n_dev, self.pk[q + ".2"], lvl.cap, ACT_RELU, res=x)
        # conv_out: SparseConv3d k(3,1,1) s(2,1,1) p0 + BN + ReLU, scattered straight into the NHWC BEV grid
        if fine:
            ops.mark(f"sp:level{len(plan['levels'])}-convs")
        ready(len(plan["levels"]))
        nl, rb = plan["out"]
        self.level_sizes.append(nl.n_dev)
        w, b = self.pk["conv_out"]
        ops.sparse_conv(x, rb, nl.n_dev, w, b, bev_out, act=ACT_RELU, cout=self.output_channels)
        if streams:
            for st in streams:
                main.wait_stream(st)
        return bev_out


@BACKBONES.register_module()
class SECOND(ParamTree):
    """[upstream] mmdet3d SECOND (cfg FocalFormer3D_L.py:207-214)."""

    def __init__(self, in_channels=128, out_channels=(128, 128, 256), layer_nums=(3, 5, 5), layer_strides=(2, 2, 2),
                 norm_cfg=None, conv_cfg=None, spec=None, in_depth=1, **kw):
        super().__init__()
        self.in_channels, self.out_channels, self.layer_nums, self.layer_strides = in_channels, list(out_channels), list(layer_nums), list(layer_strides)
        self.eps = (norm_cfg or {}).get("eps", 1e-3)
        self.in_depth = in_depth
        self.build_params(spec)
        self.pk = None

    def prepare(self, dev):
        sd = self.flat()
        self.pk = []
        for i, n in enumerate(self.layer_nums):
            layers = []
            for l in range(n + 1):
                w = sd[f"blocks.{i}.{3 * l}.weight"]
                s, b = bn_scale_shift(sd, f"blocks.{i}.{3 * l + 1}", self.eps)
                if i == 0 and l == 0 and self.in_depth > 1:
                    # reference channel order after .dense().view is c*D + d; the BEV scatter writes d*C + c
                    co, ci, kh, kw = w.shape
                    D = self.in_depth
                    w = w.view(co, ci // D, D, kh, kw).permute(0, 2, 1, 3, 4).reshape(co, ci, kh, kw)
                layers.append((pack_conv2d(w, s, dev), vec(b, dev)))
            self.pk.append(layers)

    def forward(self, x):
        """x [B,H,W,C] fp32 NHWC.  Returns the stage outputs as (fp32 NHWC, ops.Split or None) pairs.  Inside a stage the
        activations live in split form only (TMA-fed convs); the stage outputs are stored in both forms (fp32 for the
        stride-2 / transposed-conv consumers and the tests, split for the TMA consumers)."""
        outs = []
        B, H, W, _ = x.shape
        dev = x.device
        xs = None                                                   # split form of x (when the next conv can take it)
        for i, layers in enumerate(self.pk):
            st = self.layer_strides[i]
            for l, (w, b) in enumerate(layers):
                s = st if l == 0 else 1
                Ho, Wo = ((H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1)
                cin, cout = w.shape[1], self.out_channels[i]
                last = l == len(layers) - 1
                use_tma = (s == 1 or (ops.TMA_STRIDED and xs is not None)) and ops.tma_ok(w, cin, cout)
                if use_tma and xs is None:
                    xs = ops.split_rows(x)                          # one elementwise pass (the BEV scatter output is fp32)
                nxt = layers[l + 1][0] if not last else None
                want_split = last or (nxt is not None and ops.tma_ok(nxt, cout, cout))
                ys = ops.Split.empty((B, Ho, Wo), cout, dev) if (want_split and ops.tma_enabled() and cout % 64 == 0) else None
                want_f32 = last or ys is None
                y = torch.empty((B, Ho, Wo, cout), dtype=torch.float32, device=dev) if want_f32 else None
                ops.conv2d(xs if use_tma else x, w, b, y, 3, stride=s, act=ACT_RELU, out_s=ys)
                x, xs, H, W = y, ys, Ho, Wo
            outs.append((x, xs))
        return outs


@NECKS.register_module()
class SECONDFPN(ParamTree):
    """[upstream] mmdet3d SECONDFPN (cfg FocalFormer3D_L.py:215-222): level outputs written into channel slices."""

    def __init__(self, in_channels=(128, 128, 256), out_channels=(256, 256, 256), upsample_strides=(1, 2, 4),
                 norm_cfg=None, upsample_cfg=None, conv_cfg=None, use_conv_for_no_stride=False, spec=None, **kw):
        super().__init__()
        self.in_channels, self.out_channels, self.strides = list(in_channels), list(out_channels), list(upsample_strides)
        self.use_conv = use_conv_for_no_stride
        self.eps = (norm_cfg or {}).get("eps", 1e-3)
        self.build_params(spec)
        self.pk = None

    def prepare(self, dev):
        sd = self.flat()
        self.pk = []
        for i, oc in enumerate(self.out_channels):
            st = self.strides[i]
            s, b = bn_scale_shift(sd, f"deblocks.{i}.1", self.eps)
            w = sd[f"deblocks.{i}.0.weight"].double()
            if st > 1 or not self.use_conv:
                # ConvTranspose2d [cin, cout, k, k], k == stride: one 1x1 GEMM per (dy, dx) output lattice
                taps = [pack_taps((w[:, :, dy, dx] * s.view(1, -1)).unsqueeze(0), dev) for dy in range(st) for dx in range(st)]
                self.pk.append(("deconv", st, taps, vec(b, dev)))
            else:
                self.pk.append(("conv", 1, pack_conv2d(w, s, dev), vec(b, dev)))

    def forward(self, xs, out, out_s=None):
        """xs: list of (fp32 NHWC, ops.Split or None) level pairs; out [B,H,W,sum(out_channels)] fp32; out_s: optional
        ops.Split of the same shape that also receives every level (for a TMA-fed consumer)."""
        c0 = 0
        for i, (kind, st, w, b) in enumerate(self.pk):
            oc = self.out_channels[i]
            view = out[..., c0:c0 + oc]
            view_s = out_s.slice(c0, c0 + oc) if out_s is not None else None
            x, x_s = xs[i]
            if kind == "conv":
                use_tma = x_s is not None and ops.tma_ok(w, w.shape[1], oc)
                ops.conv2d(x_s if use_tma else x, w, b, view, 1, act=ACT_RELU, out_s=view_s)
            else:
                for dy in range(st):
                    for dx in range(st):
                        wl = w[dy * st + dx]
                        use_tma = x_s is not None and ops.tma_ok(wl, wl.shape[1], oc)      # lattice GEMM on the TMA-fed kernel
                        ops.conv2d(x_s if use_tma else x, wl, b, view, 1, act=ACT_RELU, up=(st, dy, dx), out_s=view_s)
            c0 += oc
        return out


@BACKBONES.register_module()
class ResNet(ParamTree):
    """[upstream] mmdet 2.14 ResNet-50 (style 'pytorch', norm_eval) as configured at DeformFormer3D_C_R50.py
    img_backbone and called from focalformer3d.py:133-153.  BatchNorm folded; every conv is one implicit-GEMM launch
    with the ReLU / residual in its epilogue."""

    STAGES = ((64, 3), (128, 4), (256, 6), (512, 3))

    def __init__(self, depth=50, num_stages=4, out_indices=(0, 1, 2, 3), style="pytorch", spec=None, **kw):
        super().__init__()
        if depth != 50 or style != "pytorch" or num_stages != 4:
            raise NotImplementedError("ResNet: only depth 50, style 'pytorch' is built (DeformFormer3D_C_R50)")
        self.out_indices = tuple(out_indices)
        self.build_params(spec)
        self.pk = None

    def prepare(self, dev):
        sd = self.flat()

        def cb(conv, bn, pad_cin=None):
            w = sd[conv + ".weight"]
            s, b = bn_scale_shift(sd, bn, 1e-5)
            if pad_cin:
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_cin - w.shape[1], *w.shape[2:])], 1)
            return pack_conv2d(w, s, dev), vec(b, dev)
        pk = {"stem": cb("conv1", "bn1", pad_cin=8)}
        for li, (planes, blocks) in enumerate(self.STAGES, start=1):
            for b in range(blocks):
                q = f"layer{li}.{b}"
                pk[q] = [cb(f"{q}.conv{n}", f"{q}.bn{n}") for n in (1, 2, 3)]
                if b == 0:
                    pk[q].append(cb(f"{q}.downsample.0", f"{q}.downsample.1"))
        self.pk = pk

    def forward(self, x):
        """x [n, H, W, 8] NHWC (3 image channels + zero padding).  Returns the four stage outputs (NHWC)."""
        n, H, W, _ = x.shape
        dev = x.device

        def new(h, w, c):
            return torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
        Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        y = new(Ho, Wo, 64)
        ops.conv2d(x, *self.pk["stem"], y, 7, stride=2, pad=3, act=ACT_RELU)
        x = ops.maxpool3x3s2(y)
        outs = []
        for li, (planes, blocks) in enumerate(self.STAGES, start=1):
            for b in range(blocks):
                layers = self.pk[f"layer{li}.{b}"]
                st = 2 if (b == 0 and li > 1) else 1
                _, H, W, _ = x.shape
                Ho, Wo = (H + 2 - 3) // st + 1, (W + 2 - 3) // st + 1
                t1 = new(H, W, planes)
                ops.conv2d(x, *layers[0], t1, 1, act=ACT_RELU)
                t2 = new(Ho, Wo, planes)
                ops.conv2d(t1, *layers[1], t2, 3, stride=st, act=ACT_RELU)
                if b == 0:
                    idn = new(Ho, Wo, planes * 4)
                    ops.conv2d(x, *layers[3], idn, 1, stride=st, pad=0, act=ACT_NONE)
                else:
                    idn = x
                y = new(Ho, Wo, planes * 4)
                ops.conv2d(t2, *layers[2], y, 1, act=ACT_RELU, res=idn)
                x = y
            if li - 1 in self.out_indices:
                outs.append(x)
        return outs


@NECKS.register_module()
class FPN(ParamTree):
    """[upstream] mmdet 2.14 FPN (no norm / activation).  Only level 0 is consumed downstream
    (focalformer3d.py:186 passes img_feats[0]), so only fpn_convs[0] is evaluated; the other fpn_convs tensors are
    accepted from the checkpoint and unused."""

    def __init__(self, in_channels, out_channels, num_outs, spec=None, **kw):
        super().__init__()
        self.in_channels, self.out_channels = list(in_channels), out_channels
        self.build_params(spec)
        self.pk = None

    def prepare(self, dev):
        sd = self.flat()
        self.pk = dict(
            lat=[(pack_conv2d(sd[f"lateral_convs.{i}.conv.weight"], None, dev), vec(sd[f"lateral_convs.{i}.conv.bias"], dev))
                 for i in range(len(self.in_channels))],
            out0=(pack_conv2d(sd["fpn_convs.0.conv.weight"], None, dev), vec(sd["fpn_convs.0.conv.bias"], dev)))

    def forward(self, xs):
        lat = []
        for x, (w, b) in zip(xs, self.pk["lat"]):
            n, H, W, _ = x.shape
            y = torch.empty((n, H, W, self.out_channels), dtype=torch.float32, device=x.device)
            ops.conv2d(x, w, b, y, 1, act=ACT_NONE)
            lat.append(y)
        for i in range(len(lat) - 1, 0, -1):
            ops.upsample_add(lat[i - 1], lat[i])
        out = torch.empty_like(lat[0])
        ops.conv2d(lat[0], *self.pk["out0"], out, 3, act=ACT_NONE)
        return out


class LiftSplatShoot:
    """models/necks/lss.py:149-383 at test time: depthnet 1x1 -> fused soft-max / lift / splat kernel -> bevencode."""

    CAM_C, LD = 64, 128

    def __init__(self, sd, prefix, pc_range, grid, hidden, dev):
        import numpy as np
        f32 = np.float32
        self.frustum = sd[prefix + ".frustum"].detach().float().contiguous().to(dev)
        self.D, self.fH, self.fW = self.frustum.shape[:3]
        # gen_dx_bx (lss.py:77-82) in the fp32 arithmetic torch.Tensor([...]) gives the reference
        lows, highs = pc_range[:3], pc_range[3:]
        self.dx = [f32(grid)] * 3
        bx = [f32(lo + grid / 2.0) for lo in lows]
        self.lo = [f32(b) - f32(d) / f32(2.0) for b, d in zip(bx, self.dx)]
        self.nxyz = [int((hi - lo) / grid) for lo, hi in zip(lows, highs)]
        D, Cc = self.D, self.CAM_C
        w = sd[prefix + ".camencode.depthnet.weight"].double().reshape(D + Cc, -1)
        b = sd[prefix + ".camencode.depthnet.bias"].double()
        order = list(range(D, D + Cc)) + list(range(D))                       # context first (16-byte aligned), then depth
        wp = torch.zeros((self.LD, w.shape[1]), dtype=torch.float64)
        wp[:D + Cc] = w[order]
        self.depthnet = (pack_linear(wp, None, dev), vec(b[order], dev, self.LD))
        Z = self.nxyz[2]
        self.bevenc = []
        for i in range(4):
            wc = sd[f"{prefix}.bevencode.{3 * i}.weight"]
            s, sh = bn_scale_shift(sd, f"{prefix}.bevencode.{3 * i + 1}", 1e-5)
            if i == 0:
                # reference channel order after the s2c reshape is c*Z + z; the splat kernel writes z*64 + c
                co, ci, kh, kw = wc.shape
                wc = wc.view(co, Cc, Z, kh, kw).permute(0, 2, 1, 3, 4).reshape(co, ci, kh, kw)
            co = wc.shape[0]
            co_pad = -(-co // 128) * 128
            if co_pad != co and co > 128:
                # 832 output channels would tile as 13 x 64 and re-gather the A operand 13 times; 7 x 128 with 64 zero
                # channels is ~4x faster on the tensor-core kernel.  The next conv reads the [:co] channel slice.
                wc = wc.double() * s.view(-1, 1, 1, 1)
                wc = torch.cat([wc, wc.new_zeros(co_pad - co, *wc.shape[1:])], 0)
                self.bevenc.append((pack_conv2d(wc, None, dev), vec(sh, dev, co_pad), co))
            else:
                self.bevenc.append((pack_conv2d(wc, s, dev), vec(sh, dev), co))

    def __call__(self, feat, rots, trans, B, out):
        """feat [B*cams, fH, fW, 256]; rots [B*cams, 9]; trans [B*cams, 3]; out [B, ny, nx, hidden] view."""
        n_img, fH, fW, _ = feat.shape
        dev = feat.device
        dn = torch.empty((n_img, fH, fW, self.LD), dtype=torch.float32, device=dev)
        ops.conv2d(feat, *self.depthnet, dn, 1, act=ACT_NONE)
        nx, ny, nz = self.nxyz
        bev = torch.empty((B, ny, nx, nz * self.CAM_C), dtype=torch.float32, device=dev)
        ops.lss_splat(dn, self.frustum, rots, trans, bev, n_img // B, self.D, self.lo, self.dx)
        ops.mark("lss_lift_splat")
        tma = all(ops.tma_ok(w, (w.shape[1] if i == 0 else self.bevenc[i - 1][2]), w.shape[-1]) for i, (w, b, co) in enumerate(self.bevenc))
        if tma:
            # bevencode on the TMA-fed kernel: the splatted volume is split once, the three intermediate maps exist in split
            # form only; K = 9 x 832 = 7488 here, which is where the chunked accumulation of tmagemm.cu matters most
            xs = ops.split_rows(bev)
            for i, (w, b, co) in enumerate(self.bevenc):
                last = i == 3
                ys = None if last else ops.Split.empty((B, ny, nx), w.shape[-1], dev)
                ops.conv2d(xs, w, b, out if last else None, 3, act=ACT_RELU, out_s=ys)
                xs = None if last else ys.slice(0, co)
        else:
            x = bev
            for i, (w, b, co) in enumerate(self.bevenc):
                y = out if i == 3 else torch.empty((B, ny, nx, w.shape[-1]), dtype=torch.float32, device=dev)
                ops.conv2d(x, w, b, y, 3, act=ACT_RELU)
                x = y[..., :co]
        ops.mark("lss_bevencode")
        return out, dn, bev


class _IR:
    """torchvision InvertedResidual (focal_encoder.py:36-38) packed: optional 1x1 expand, dw3x3, 1x1 project."""

    def __init__(self, sd, name, expand, dev, eps=1e-5):
        i = 0
        self.expand = None
        if expand != 1:
            s, b = bn_scale_shift(sd, f"{name}.conv.0.1", eps)
            self.expand = (pack_conv2d(sd[f"{name}.conv.0.0.weight"], s, dev), vec(b, dev))
            i = 1
        s, b = bn_scale_shift(sd, f"{name}.conv.{i}.1", eps)
        wd = sd[f"{name}.conv.{i}.0.weight"].double() * s.view(-1, 1, 1, 1)          # [C,1,3,3]
        self.dw = (wd.reshape(wd.shape[0], 9).t().contiguous().float().to(dev), vec(b, dev))
        s, b = bn_scale_shift(sd, f"{name}.conv.{i + 2}", eps)
        self.proj = (pack_conv2d(sd[f"{name}.conv.{i + 1}.weight"], s, dev), vec(b, dev))

    def run(self, x, out, res=None, x_s=None, out_s=None):
        """x [B,H,W,cin] fp32 view (x_s: the same in split form for the TMA-fed expand conv); out fp32 view; out_s: Split that
        also receives the output.  The depthwise conv writes split rows when the project conv can take them."""
        B, H, W, _ = x.shape
        hid = self.dw[0].shape[1]
        if self.expand is not None:
            t = torch.empty((B, H, W, hid), dtype=torch.float32, device=x.device)
            use_tma = x_s is not None and ops.tma_ok(self.expand[0], self.expand[0].shape[1], hid)
            ops.conv2d(x_s if use_tma else x, self.expand[0], self.expand[1], t, 1, act=ACT_RELU6)
            x = t
        if ops.tma_ok(self.proj[0], hid, self.proj[0].shape[-1]):
            t2 = ops.dwconv3x3_split(x, self.dw[0], self.dw[1], act=ACT_RELU6)
        else:
            t2 = torch.empty((B, H, W, hid), dtype=torch.float32, device=x.device)
            ops.dwconv3x3(x, self.dw[0], self.dw[1], t2, act=ACT_RELU6)
        ops.conv2d(t2, self.proj[0], self.proj[1], out, 1, act=ACT_NONE, res=res, out_s=out_s)
        return out


@NECKS.register_module()
class FocalEncoder(ParamTree):
    """models/necks/focal_encoder.py:90-222.  Three built paths: LiDAR-only 'bevfusionmb2' (input_img=False,
    iterbev_wo_img=True; `forward`), camera-only Lift-Splat-Shoot (`forward_camera`), LiDAR + camera 'bevfusion' with
    cam_lss / iter_bev_cam (`forward_fusion`)."""

    def __init__(self, num_layers=2, in_channels_img=64, in_channels_pts=384, hidden_channel=128, bn_momentum=0.1,
                 bias="auto", iterbev="bevfusion", max_points_height=5, multistage_heatmap=False, input_img=True,
                 input_pts=True, iterbev_wo_img=False, extra_feat=False, iter_bev_cam=False, cam_lss=False,
                 newbevpool=False, pc_range=None, img_scale=None, spec=None, **kw):
        super().__init__()
        lss_ok = bool(input_img and cam_lss and cam_lss != "proj")
        self.camera_only = bool(lss_ok and not input_pts and not num_layers and not multistage_heatmap)
        # LiDAR + camera (FocalFormer3D_LC): camera BEV from cam_lss, 'bevfusion' blocks, iter_bev_cam (no I2P projection)
        self.fusion = bool(lss_ok and input_pts and iterbev == "bevfusion" and iter_bev_cam and not iterbev_wo_img
                           and multistage_heatmap and num_layers)
        lidar_only = (not input_img) and input_pts and iterbev == "bevfusionmb2" and iterbev_wo_img
        if not (self.camera_only or self.fusion or lidar_only):
            raise NotImplementedError("FocalEncoder: built paths are LiDAR-only 'bevfusionmb2', camera-only Lift-Splat-Shoot "
                                      "and LiDAR + camera 'bevfusion' with cam_lss / iter_bev_cam")
        self.pc_range, self.img_scale = pc_range, img_scale
        self.num_layers = num_layers or 0
        self.hidden = hidden_channel
        self.multistage_heatmap, self.extra_feat = multistage_heatmap, extra_feat
        self.build_params(spec)
        self.pk = None

    def prepare(self, dev):
        sd = self.flat()
        if self.camera_only:
            self.pk = {"lss": LiftSplatShoot(sd, "cam_lss", list(self.pc_range), 0.6, self.hidden, dev)}   # :129-131
            return
        pk = {"shared": (pack_conv2d(sd["shared_conv_pts.weight"], None, dev), vec(sd["shared_conv_pts.bias"], dev))}
        if self.fusion:
            pk["lss"] = LiftSplatShoot(sd, "cam_lss", list(self.pc_range), 0.6, self.hidden, dev)

            def cbn(name):                                   # ConvBNReLU (encoder_utils.py:10-33), BN folded
                s_, b_ = bn_scale_shift(sd, name + ".bn", 1e-5)
                return sd[name + ".conv.weight"].double() * s_.view(-1, 1, 1, 1), b_
            for i in range(self.num_layers):
                q = f"fusion_blocks.{i}"
                # first 1x1 of the query / key projections and the value projection read the same tensor: one GEMM
                ws, bs = zip(*[cbn(f"{q}.P_IML.{n}") for n in ("query_project.0", "key_project.0", "value_project")])
                blk = dict(qkv1=(pack_conv2d(torch.cat(ws, 0), None, dev), vec(torch.cat(bs), dev)))
                for n, key in (("query_project.1", "q2"), ("key_project.1", "k2"), ("P_out_proj", "outp"), ("P_integration", "integ")):
                    w_, b_ = cbn(f"{q}.{n}" if n.startswith("P_") else f"{q}.P_IML.{n}")
                    blk[key] = (pack_conv2d(w_, None, dev), vec(b_, dev))
                if i + 1 < self.num_layers:                  # the last block's camera-BEV update is never consumed
                    for n in (1, 2):
                        s_, b_ = bn_scale_shift(sd, f"{q}.iterimg_conv.0.bn{n}", 1e-5)
                        blk[f"img{n}"] = (pack_conv2d(sd[f"{q}.iterimg_conv.0.conv{n}.weight"], s_, dev), vec(b_, dev))
                pk[q] = blk
            if self.extra_feat:
                s, b = bn_scale_shift(sd, "extra_output.bn", 1e-5)
                pk["extra"] = (pack_conv2d(sd["extra_output.conv.weight"], s, dev), vec(b, dev))
            self.pk = pk
            return
        for i in range(self.num_layers):
            q = f"fusion_blocks.{i}"
            pk[q] = (_IR(sd, q + ".P_IML", 2, dev), _IR(sd, q + ".P_out_proj", 1, dev), _IR(sd, q + ".P_integration", 1, dev))
        if self.extra_feat:
            s, b = bn_scale_shift(sd, "extra_output.bn", 1e-5)
            pk["extra"] = (pack_conv2d(sd["extra_output.conv.weight"], s, dev), vec(b, dev))
        self.pk = pk

    @staticmethod
    def camera_rots_trans(img_metas, dev):
        """focal_encoder.py:178-194: per camera inverse(lidar2img) -> rotation [B*N, 9] and translation [B*N, 3].
        36 4x4 inverses: done on the host (LAPACK), uploaded once per call."""
        for m in img_metas:
            FocalEncoder._check_identity_metas(m)
        mats = torch.stack([torch.as_tensor(m["lidar2img"], dtype=torch.float32).reshape(-1, 4, 4) for m in img_metas])
        inv = torch.inverse(mats.cpu())
        rots = inv[..., :3, :3].reshape(-1, 9).contiguous().to(dev)
        trans = inv[..., :3, 3].reshape(-1, 3).contiguous().to(dev)
        return rots, trans

    @staticmethod
    def _check_identity_metas(m):
        """lss.py:239-267 get_geometry also un-applies img_metas['img_aug_matrix'] (post_rots / post_trans) and runs
        apply_3d_transformation (pcd_rotation, pcd_scale_factor, pcd_trans, flips) on the frustum points.  The shipped
        non-TTA test pipeline emits identities; anything else would silently give wrong camera-BEV features here."""
        import numpy as np

        def ident(v, ref):
            return v is None or np.allclose(np.asarray(v, dtype=np.float64), ref)
        bad = []
        if not ident(m.get("img_aug_matrix"), np.eye(4)):
            bad.append("img_aug_matrix")
        if not ident(m.get("pcd_rotation"), np.eye(3)):
            bad.append("pcd_rotation")
        if not ident(m.get("pcd_scale_factor"), 1.0):
            bad.append("pcd_scale_factor")
        if not ident(m.get("pcd_trans"), 0.0):
            bad.append("pcd_trans")
        for k in ("pcd_horizontal_flip", "pcd_vertical_flip", "flip"):
            if m.get(k):
                bad.append(k)
        if bad:
            raise NotImplementedError("Lift-Splat-Shoot geometry: non-identity augmentation metas %s are not applied "
                                      "(lss.py:239-267); only the shipped non-TTA test pipeline is built" % bad)

    def forward_camera(self, img_feat, img_metas, out):
        """img_feat [B*N, fH, fW, 256] (FPN level 0) -> out [B, ny, nx, hidden]; returned twice by the reference
        (focal_encoder.py:196-197: conv_feat and decoder feature are the same tensor)."""
        rots, trans = self.camera_rots_trans(img_metas, img_feat.device)
        return self.pk["lss"](img_feat, rots, trans, len(img_metas), out)

    def forward_fusion(self, pts_feats, img_feat, img_metas, extra_out, neck_s=None):
        """focal_encoder.py:171-219 with both towers.  pts_feats [B,H,W,512] (SECONDFPN), img_feat [B*N,fH,fW,256] (FPN
        level 0).  Concatenations are channel slices of two ping-pong buffers per layer: catA = [camera BEV | P2P],
        catB = [P_Aug | LiDAR BEV]; every producer writes straight into its slice.
        Returns (conv_feat, [stage features], extra, camera BEV of layer 0)."""
        B, H, W, _ = pts_feats.shape
        dev, hc, pk = pts_feats.device, self.hidden, self.pk
        new = lambda c: torch.empty((B, H, W, c), dtype=torch.float32, device=dev)
        catA, catB = new(2 * hc), new(2 * hc)
        rots, trans = self.camera_rots_trans(img_metas, dev)
        pk["lss"](img_feat, rots, trans, B, catA[..., :hc])                                        # :196 camera BEV
        img_bev0 = catA[..., :hc]
        conv_feat = new(hc)
        sw = pk["shared"][0]
        shared_in = neck_s if (neck_s is not None and ops.tma_ok(sw, sw.shape[1], hc)) else pts_feats
        ops.conv2d(shared_in, *pk["shared"], conv_feat, 3, act=ACT_NONE)                           # :204
        catB[..., hc:].copy_(conv_feat)                                                            # :207 (.clone())
        stages = []
        for i in range(self.num_layers):
            blk = pk[f"fusion_blocks.{i}"]
            lidar = catB[..., hc:]
            qkv = new(3 * hc)
            ops.conv2d(lidar, *blk["qkv1"], qkv, 1, act=ACT_RELU)                                  # encoder_utils.py:156-158
            q2, k2 = new(hc), new(hc)
            ops.conv2d(qkv[..., :hc], *blk["q2"], q2, 1, act=ACT_RELU)
            ops.conv2d(qkv[..., hc:2 * hc], *blk["k2"], k2, 1, act=ACT_RELU)
            ops.local_attention(q2, k2, qkv[..., 2 * hc:], catA[..., hc:], 9)                      # :160-162 -> P2P
            ops.mark("locatt")
            ops.conv2d(catA, *blk["outp"], catB[..., :hc], 1, act=ACT_NONE)                        # focal_encoder.py:73
            last = i == self.num_layers - 1
            nxtB = new(hc) if last else new(2 * hc)
            new_feat = nxtB if last else nxtB[..., hc:]
            ops.conv2d(catB, *blk["integ"], new_feat, 1, act=ACT_NONE)                             # :74
            stages.append(new_feat)
            if not last:
                nxtA = new(2 * hc)
                t = new(hc)
                ops.conv2d(catA[..., :hc], *blk["img1"], t, 3, act=ACT_RELU)                        # :85 BasicBlock
                ops.conv2d(t, *blk["img2"], nxtA[..., :hc], 3, act=ACT_RELU, res=catA[..., :hc])
                if i == 0:
                    img_bev0 = catA[..., :hc]
                catA, catB = nxtA, nxtB
        extra = None
        if self.extra_feat:
            extra = extra_out if extra_out is not None else new(hc)
            ops.conv2d(stages[-1], *pk["extra"], extra, 3, act=ACT_NONE)                           # :218-219
        return conv_feat, stages, extra, img_bev0

    def forward(self, pts_feats, extra_out=None, neck_s=None):
        """pts_feats [B,H,W,512] NHWC (neck_s: the same in split form for the TMA-fed shared conv).
        Returns (conv_feat view, [stage feature views...], extra view)."""
        B, H, W, _ = pts_feats.shape
        dev, hc = pts_feats.device, self.hidden
        cat1 = torch.empty((B, H, W, 2 * hc), dtype=torch.float32, device=dev)
        feat = cat1[..., :hc]
        sw = self.pk["shared"][0]
        shared_in = neck_s if (neck_s is not None and ops.tma_ok(sw, sw.shape[1], hc)) else pts_feats
        tma = ops.tma_enabled()
        feat_s = ops.Split.empty((B, H, W), hc, dev) if tma else None      # split copies feed the TMA-fed 3x3 / 1x1 consumers
        ops.conv2d(shared_in, sw, self.pk["shared"][1], feat, 3, act=ACT_NONE, out_s=feat_s)       # :204
        conv_feat = feat
        self.split_feats = dict(conv_feat=feat_s, stages=[])
        stages = []
        for i in range(self.num_layers):
            iml, outp, integ = self.pk[f"fusion_blocks.{i}"]
            cat2 = torch.empty((B, H, W, 2 * hc), dtype=torch.float32, device=dev)
            cat2[..., hc:].copy_(feat)                                       # cat((P_Aug, lidar_feat)) :78
            iml.run(feat, cat1[..., hc:], res=feat, x_s=feat_s)              # P2P = P_IML(lidar) :76 -> cat((lidar, P2P))
            outp.run(cat1, cat2[..., :hc])                                   # :77
            last = i == self.num_layers - 1
            nxt = torch.empty((B, H, W, hc if last else 2 * hc), dtype=torch.float32, device=dev)
            new_feat = nxt[..., :hc]
            new_s = ops.Split.empty((B, H, W), hc, dev) if tma else None
            integ.run(cat2, new_feat, out_s=new_s)                           # :78
            stages.append(new_feat)
            self.split_feats["stages"].append(new_s)
            feat, feat_s, cat1 = new_feat, new_s, nxt
        extra = None
        if not stages and extra_out is not None:
            # DeformFormer3D (num_layers=None): the shared-conv output itself is the decoder's level-0 feature
            # (focal_encoder.py:220-222 returns it twice; focal_decoder.py:541-545,812)
            extra_out.copy_(conv_feat)
            extra = extra_out
        if self.extra_feat and stages:
            extra = extra_out if extra_out is not None else torch.empty((B, H, W, hc), dtype=torch.float32, device=dev)
            ew = self.pk["extra"][0]
            last_s = self.split_feats["stages"][-1]
            ops.conv2d(last_s if (last_s is not None and ops.tma_ok(ew, hc, hc)) else stages[-1], ew, self.pk["extra"][1], extra, 3,
                       act=ACT_NONE)                                                               # :218-219
        return conv_feat, stages, extra


@BBOX_CODERS.register_module()
class TransFusionBBoxCoder:
    """core/bbox/coders/transfusion_bbox_coder.py:8-22 (parameters only; decode runs in ff3d_box_decode)."""

    def __init__(self, pc_range, out_size_factor, voxel_size, post_center_range=None, score_threshold=None,
                 code_size=8, **kw):
        self.pc_range, self.out_size_factor, self.voxel_size = list(pc_range), out_size_factor, list(voxel_size)
        self.post_center_range, self.score_threshold, self.code_size = post_center_range, score_threshold, code_size
        if score_threshold:
            raise NotImplementedError("score_threshold > 0 is not used by any shipped config")

    @property
    def cell(self):
        return (self.out_size_factor * self.voxel_size[0], self.out_size_factor * self.voxel_size[1])


@HEADS.register_module()
class FocalDecoder(ParamTree):
    """models/dense_heads/focal_decoder.py:34 -- eval forward (:522-992) + get_bboxes (:1313-1413)."""

    def __init__(self, num_proposals=128, hidden_channel=128, hidden_channel_roi=512, num_classes=4,
                 num_decoder_layers=1, num_heads=8, initialize_by_heatmap=False, nms_kernel_size=1,
                 common_heads=dict(), num_heatmap_convs=2, bias="auto", train_cfg=None, test_cfg=None, bbox_coder=None,
                 multiscale=False, multistage_heatmap=False, reuse_first_heatmap=False, extra_feat=False,
                 heatmap_box=False, bevpos=False, input_img=True, iterbev_wo_img=False, mask_heatmap_mode="poscls",
                 roi_feats=0, roi_dropout_rate=0.0, roi_expand_ratio=1.0, roi_based_reg=False, classaware_reg=False,
                 boxpos=None, decoder_cfg=None, spec=None, loss_cls=dict(type='GaussianFocalLoss', reduction='mean'),
                 **unused):
        super().__init__()
        if not (initialize_by_heatmap and multiscale and bevpos and mask_heatmap_mode == "poscls") \
                or heatmap_box or boxpos is not None:
            raise NotImplementedError("FocalDecoder: only the shipped LiDAR head variants are built")
        self.classaware_reg = bool(classaware_reg)
        if self.classaware_reg and "vel" in common_heads:
            # focal_decoder.py:317-319 widens every head by num_classes but :940-943 class-selects only
            # center/height/dim/rot: a class-aware 'vel' head keeps nc*2 channels in the reference.  Not reproduced.
            raise NotImplementedError("FocalDecoder: classaware_reg with a 'vel' head is not built (no shipped config)")
        if not loss_cls.get("use_sigmoid", False):
            # focal_decoder.py:164-166 appends a background class in that case; no shipped config uses it
            raise NotImplementedError("FocalDecoder: only loss_cls.use_sigmoid=True heads are built")
        stages = (multistage_heatmap or 0) + (1 if reuse_first_heatmap else 0)
        if stages >= 1 and not extra_feat:
            raise NotImplementedError("FocalDecoder: multi-stage heads without extra_feat are not used by any shipped config")
        if stages < 1 and not (input_img or iterbev_wo_img):
            raise NotImplementedError("FocalDecoder: single-stage head without heatmap_head_img")
        self.single_stage = stages < 1            # DeformFormer3D: focal_decoder.py:539-586
        self.num_classes, self.num_proposals, self.hc = num_classes, num_proposals, hidden_channel
        self.num_decoder_layers, self.num_heads = num_decoder_layers, num_heads
        self.nms_kernel_size, self.test_cfg = nms_kernel_size, test_cfg
        self.nms_type = (test_cfg or {}).get("nms_type")
        if self.nms_type not in (None, "circle", "rotate"):
            raise NotImplementedError(f"FocalDecoder: test_cfg.nms_type={self.nms_type!r} (focal_decoder.py:1352-1385)")
        self.stages, self.reuse_first = stages, reuse_first_heatmap
        self.roi_feats, self.roi_based_reg = roi_feats, roi_based_reg
        self.roi_expand_ratio = [roi_expand_ratio] * num_decoder_layers if isinstance(roi_expand_ratio, float) else list(roi_expand_ratio)
        self.roi_step = 4 if roi_dropout_rate > 1e-4 else 3
        self.common_heads = dict(common_heads)
        self.bbox_coder = BBOX_CODERS.build(bbox_coder)
        tl = decoder_cfg["transformerlayers"]
        self.n_layers = decoder_cfg["num_layers"]
        self.ffn_ch = tl["feedforward_channels"]
        m = tl["attn_cfgs"][1]
        self.n_levels, self.n_points = m["num_levels"], m["num_points"]
        assert tl["attn_cfgs"][0]["num_heads"] == num_heads and m["num_heads"] == num_heads
        ds = test_cfg["dataset"]
        self.exempt = (8, 9) if ds == "nuScenes" else (1, 2) if ds == "Waymo" else (1, 0)
        # hard-coded in the reference (focal_decoder.py:903-906), NOT the config range
        self.roi_range = (-54.0, -54.0, 54.0, 54.0) if ds == "nuScenes" else (-75.2, -75.2, 75.2, 75.2)
        self.has_vel = "vel" in self.common_heads
        self.build_params(spec)
        self.pk = None

    # ---- weights
    def prepare(self, dev):
        sd = self.flat()
        hc, nc = self.hc, self.num_classes
        pk = {}

        def heat(name):
            s, b = bn_scale_shift(sd, f"{name}.0.bn", 1e-5)
            w2 = sd[f"{name}.1.weight"]
            co = (w2.shape[0] + 15) // 16 * 16        # zero-pad cout (10 -> 16) so the conv tiles on the tensor cores
            w2p = torch.zeros((co,) + tuple(w2.shape[1:]), dtype=w2.dtype)
            w2p[:w2.shape[0]] = w2.detach().cpu()
            return (pack_conv2d(sd[f"{name}.0.conv.weight"], s, dev), vec(b, dev),
                    pack_conv2d(w2p, None, dev), vec(sd[f"{name}.1.bias"], dev, co))
        pk["heat"] = []
        if self.single_stage:
            pk["heat"] = [heat("heatmap_head"), heat("heatmap_head_img")]
        for i in range(self.stages):
            pk["heat"].append(heat("heatmap_head") if (i == 0 and self.reuse_first) else heat(f"heatmap_head_img.{i}"))
        if not self.single_stage and not self.reuse_first:
            # focal_decoder.py:588,663-664: heatmap_head(conv_feat) is still evaluated and returned as the first
            # 'dense_heatmap' entry (a training-loss output; it feeds no proposal)
            pk["heat_first"] = heat("heatmap_head")
        pk["cls_w"] = sd["class_encoding.weight"].reshape(hc, nc).t().contiguous().float().to(dev)      # [C, Cf]
        pk["cls_b"] = vec(sd["class_encoding.bias"], dev)
        for n in ("dconv", "dconv2"):
            s, b = bn_scale_shift(sd, f"{n}.bn", 1e-5)
            pk[n] = (pack_conv2d(sd[f"{n}.conv.weight"], s, dev), vec(b, dev))
        if self.roi_feats:
            g2, L = self.roi_feats ** 2, self.n_levels
            w0 = sd["roi_mlp.0.weight"].double()                           # [512, (lvl, c, pt)] focal_decoder.py:919
            w0 = w0.view(w0.shape[0], L, hc, g2).permute(0, 1, 3, 2).reshape(w0.shape[0], -1)   # -> (lvl, pt, c)
            ws = [w0, sd[f"roi_mlp.{self.roi_step}.weight"].double(), sd[f"roi_mlp.{2 * self.roi_step}.weight"].double()]
            pk["roi"] = []
            for i, w in enumerate(ws):
                s, b = bn_scale_shift(sd, f"roi_mlp.{i * self.roi_step + 1}", 1e-5)
                # (measured: a 64-wide N tile for roi_mlp.0 -- more CTAs for its ~2400 rows -- is 1.5x SLOWER:
                #  the 180 MB gathered-ROI operand is then re-read 8x instead of 4x; kept on the default 128)
                pk["roi"].append((pack_linear(w, s, dev), vec(b, dev)))
        pk["dim_t"] = (10000 ** (2 * (torch.arange(128, dtype=torch.float32) // 2) / 128)).to(dev)
        pk["stage"] = []
        for i in range(self.num_decoder_layers):
            st = {"pos": [(pack_linear(sd[f"pos_embed_learned.{i}.layers.{j}.weight"], None, dev),
                           vec(sd[f"pos_embed_learned.{i}.layers.{j}.bias"], dev)) for j in range(2)]}
            layers, vw, vb = [], [], []
            for j in range(self.n_layers):
                q = f"decoder.{i}.layers.{j}"
                ipw, ipb = sd[f"{q}.attentions.0.attn.in_proj_weight"], sd[f"{q}.attentions.0.attn.in_proj_bias"]
                lay = dict(
                    qk=(pack_linear(ipw[:2 * hc], None, dev), vec(ipb[:2 * hc], dev)),
                    v=(pack_linear(ipw[2 * hc:], None, dev), vec(ipb[2 * hc:], dev)),
                    o=(pack_linear(sd[f"{q}.attentions.0.attn.out_proj.weight"], None, dev),
                       vec(sd[f"{q}.attentions.0.attn.out_proj.bias"], dev)),
                    oa=(pack_linear(torch.cat([sd[f"{q}.attentions.1.sampling_offsets.weight"],
                                               sd[f"{q}.attentions.1.attention_weights.weight"]]), None, dev),
                        vec(torch.cat([sd[f"{q}.attentions.1.sampling_offsets.bias"],
                                       sd[f"{q}.attentions.1.attention_weights.bias"]]), dev)),
                    op=(pack_linear(sd[f"{q}.attentions.1.output_proj.weight"], None, dev),
                        vec(sd[f"{q}.attentions.1.output_proj.bias"], dev)),
                    f1=(pack_linear(sd[f"{q}.ffns.0.layers.0.0.weight"], None, dev), vec(sd[f"{q}.ffns.0.layers.0.0.bias"], dev)),
                    f2=(pack_linear(sd[f"{q}.ffns.0.layers.1.weight"], None, dev), vec(sd[f"{q}.ffns.0.layers.1.bias"], dev)),
                    ln=[(vec(sd[f"{q}.norms.{n}.weight"], dev), vec(sd[f"{q}.norms.{n}.bias"], dev)) for n in range(3)])
                layers.append(lay)
                vw.append(sd[f"{q}.attentions.1.value_proj.weight"])
                vb.append(sd[f"{q}.attentions.1.value_proj.bias"])
            st["layers"] = layers
            st["vproj"] = (pack_linear(torch.cat(vw), None, dev), vec(torch.cat(vb), dev))   # 3 layers batched: N = 3*hc
            # prediction heads: 6 x (Conv1d hc->64 + BN + ReLU, Conv1d 64->k): one GEMM + one block-diagonal GEMM
            names = list(self.common_heads.keys()) + ["heatmap"]
            mult = nc if self.classaware_reg else 1                 # :317-319 one regression set per class
            ks = [self.common_heads[n][0] * mult for n in self.common_heads] + [nc]
            w1, b1 = [], []
            w2 = torch.zeros((sum(ks), 64 * len(names)), dtype=torch.float64)
            b2, r0 = [], 0
            for hi, (n, k) in enumerate(zip(names, ks)):
                q = f"prediction_heads.{i}.{n}"
                s, b = bn_scale_shift(sd, f"{q}.0.bn", 1e-5)
                w1.append(sd[f"{q}.0.conv.weight"].double().reshape(64, hc) * s.view(-1, 1))
                b1.append(b)
                w2[r0:r0 + k, hi * 64:(hi + 1) * 64] = sd[f"{q}.1.weight"].double().reshape(k, 64)
                b2.append(sd[f"{q}.1.bias"].double())
                r0 += k
            st["head1"] = (pack_linear(torch.cat(w1), None, dev), vec(torch.cat(b1), dev))
            st["head2"] = (pack_linear(w2, None, dev), vec(torch.cat(b2), dev, _pad4(sum(ks))))
            st["pred_dim"] = sum(ks)
            # the whole stage as one launch (csrc/decstage.cu): weights in MMA fragment order
            n_oa = (self.num_heads * self.n_levels * self.n_points * 3 + 15) // 16 * 16
            n_h1 = (64 * len(names) + 127) // 128 * 128
            n_pred = (sum(ks) + 15) // 16 * 16
            if (ops.FUSED_DECODER and hc == 128 and self.num_heads == 8 and 1 <= self.n_layers <= 4 and self.ffn_ch % 256 == 0
                    and self.n_levels <= 4 and 32 * (n_oa + 4) * 4 <= 51200 and n_h1 <= 384):
                fl = []
                for j in range(self.n_layers):
                    q = f"decoder.{i}.layers.{j}"
                    woa = torch.cat([sd[f"{q}.attentions.1.sampling_offsets.weight"], sd[f"{q}.attentions.1.attention_weights.weight"]])
                    boa = torch.cat([sd[f"{q}.attentions.1.sampling_offsets.bias"], sd[f"{q}.attentions.1.attention_weights.bias"]])
                    fl.append(dict(
                        w_qkv=ops.pack_frag(sd[f"{q}.attentions.0.attn.in_proj_weight"]).to(dev),
                        b_qkv=vec(sd[f"{q}.attentions.0.attn.in_proj_bias"], dev),
                        w_o=ops.pack_frag(sd[f"{q}.attentions.0.attn.out_proj.weight"]).to(dev),
                        b_o=vec(sd[f"{q}.attentions.0.attn.out_proj.bias"], dev),
                        w_oa=ops.pack_frag(woa, n_pad=n_oa).to(dev), b_oa=vec(boa, dev, n_oa),
                        w_op=ops.pack_frag(sd[f"{q}.attentions.1.output_proj.weight"]).to(dev),
                        b_op=vec(sd[f"{q}.attentions.1.output_proj.bias"], dev),
                        w_f1=ops.pack_frag(sd[f"{q}.ffns.0.layers.0.0.weight"]).to(dev), b_f1=vec(sd[f"{q}.ffns.0.layers.0.0.bias"], dev),
                        w_f2=ops.pack_frag(sd[f"{q}.ffns.0.layers.1.weight"]).to(dev), b_f2=vec(sd[f"{q}.ffns.0.layers.1.bias"], dev),
                        ln=layers[j]["ln"]))
                st["fused"] = dict(layers=fl, heads=self.num_heads, ffn=self.ffn_ch, n_oa=n_oa,
                                   w_h1=ops.pack_frag(torch.cat(w1), n_pad=n_h1).to(dev), b_h1=vec(torch.cat(b1), dev, n_h1), n_h1=n_h1,
                                   w_h2=ops.pack_frag(w2, n_pad=n_pred, k_pad=n_h1).to(dev), b_h2=vec(torch.cat(b2), dev, n_pred),
                                   n_pred=n_pred)
            pk["stage"].append(st)
        self.pk = pk
        self._bev_pos_cache = {}

    def _bev_pos_embed(self, i, geom, W0, H0, dev):
        """pos_embed_learned[i](sine(bev_pos / (W,H))) for all pyramid cells (focal_decoder.py:883-885): input
        independent, computed once per (stage, geometry) with the same kernels and cached."""
        key = (i, tuple(geom.shapes))
        if key not in self._bev_pos_cache:
            pos = []
            for lvl, (h, w) in enumerate(geom.shapes):
                scale = float(2 ** lvl)
                ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
                pos.append(torch.stack([(xs + 0.5) * scale, (ys + 0.5) * scale], -1).reshape(-1, 2))
            pos = torch.cat(pos).contiguous().to(dev)
            self._bev_pos_cache[key] = self._pos_mlp(i, pos, W0, H0)
        return self._bev_pos_cache[key]

    def _pos_mlp(self, i, pos, W0, H0):
        st = self.pk["stage"][i]
        sine = ops.sine_embed(pos, W0, H0, self.pk["dim_t"])
        h = ops.linear(sine, st["pos"][0][0], st["pos"][0][1], act=ACT_RELU)
        return ops.linear(h, st["pos"][1][0], st["pos"][1][1])

    # ---- forward
    def forward(self, conv_feat, stage_feats, ms_value, geom, split_feats=None):
        """conv_feat / stage_feats: NHWC views [B,H,W,hc]; ms_value [B, n_tokens, hc] with level 0 (= extra feature)
        already written; split_feats: optional dict(conv_feat=Split, stages=[Split...]) -- split copies of the same maps
        for the TMA-fed heat-map convs.  Returns the reference's head dict (channel-major tensors) plus raw token-major
        state."""
        pk = self.pk
        B, H, W, hc = conv_feat.shape
        dev = conv_feat.device
        nc, k = self.num_classes, self.num_proposals
        n_sel = 1 if self.single_stage else self.stages
        nq = k * n_sel
        feats = ([conv_feat] if (self.reuse_first or self.single_stage) else []) + list(stage_feats)
        assert len(feats) == n_sel
        sf = split_feats or {}
        feats_s = ([sf.get("conv_feat")] if (self.reuse_first or self.single_stage) else []) + list(sf.get("stages") or [])
        feats_s = feats_s + [None] * (len(feats) - len(feats_s))
        # --- HIP stages (:588-791)
        acc_mask = torch.ones((B, nc, H, W), dtype=torch.float32, device=dev)
        q_feat = torch.empty((B * nq, hc), dtype=torch.float32, device=dev)
        q_pos = torch.empty((B * nq, 2), dtype=torch.float32, device=dev)
        q_score = torch.empty((B * nq, nc), dtype=torch.float32, device=dev)
        q_label = torch.empty((B * nq,), dtype=torch.int32, device=dev)
        dense_heatmaps, nms_heats, tops = [], [], []
        def heat_logits(hw, x, x_s=None):
            """ConvModule 3x3 + BN + ReLU, Conv 3x3 -> class logits (focal_decoder.py:202-221).  With a split copy of the
            input both convs run TMA-fed and the intermediate map exists in split form only."""
            w1, b1, w2, b2 = hw
            lg = torch.empty((B, H, W, w2.shape[-1]), dtype=torch.float32, device=dev)       # padded channels = 0
            if x_s is not None and ops.tma_ok(w1, hc, hc) and ops.tma_ok(w2, hc, w2.shape[-1]):
                t_s = ops.Split.empty((B, H, W), hc, dev)
                ops.conv2d(x_s, w1, b1, None, 3, act=ACT_RELU, out_s=t_s)
                ops.conv2d(t_s, w2, b2, lg, 3, act=ACT_NONE)
                return lg
            t = torch.empty((B, H, W, hc), dtype=torch.float32, device=dev)
            ops.conv2d(x, w1, b1, t, 3, act=ACT_RELU)
            ops.conv2d(t, w2, b2, lg, 3, act=ACT_NONE)
            return lg

        if "heat_first" in pk:
            dense_heatmaps.append(heat_logits(pk["heat_first"], conv_feat, sf.get("conv_feat")))
        for s in range(n_sel):
            logits2 = None
            if self.single_stage:      # heatmap = (sigmoid(heatmap_head(x)) + sigmoid(heatmap_head_img(x))) / 2   :547-549
                logits = heat_logits(pk["heat"][0], feats[0], feats_s[0])
                logits2 = heat_logits(pk["heat"][1], feats[0], feats_s[0])
                dense_heatmaps.append(logits)
            else:
                logits = heat_logits(pk["heat"][s], feats[s], feats_s[s])
            nms_heat = torch.empty((B, nc, H, W), dtype=torch.float32, device=dev)
            top = torch.empty((B, k), dtype=torch.int32, device=dev)
            ops.hip_stage(logits, acc_mask, nms_heat, feats[s], pk["cls_w"], pk["cls_b"], k, self.nms_kernel_size,
                          self.exempt, s * k, nq, top, q_feat, q_pos, q_score, q_label, logits2=logits2)
            dense_heatmaps.append(logits if logits2 is None else logits2)
            nms_heats.append(nms_heat)
            tops.append(top)
            ops.mark(f"hip_stage{s}")
        # --- multi-scale pyramid (:810-823): dconv / dconv2 write levels 1, 2 of the value buffer
        lv = [ms_value[:, geom.starts[l]:geom.starts[l] + h * w].view(B, h, w, hc) for l, (h, w) in enumerate(geom.shapes)]
        ops.conv2d(lv[0], pk["dconv"][0], pk["dconv"][1], lv[1], 3, stride=2, act=ACT_RELU)
        ops.conv2d(lv[1], pk["dconv2"][0], pk["dconv2"][1], lv[2], 3, stride=2, act=ACT_RELU)
        ops.mark("pyramid")
        # --- decoder stages (:826-958)
        preds, prev = [], None
        x = q_feat
        heads, d = self.num_heads, hc // self.num_heads
        LP = self.n_levels * self.n_points
        n_off = heads * LP * 2
        for i in range(self.num_decoder_layers):
            st = pk["stage"][i]
            qpe = self._pos_mlp(i, q_pos, W, H)                                           # :869-872
            bev_pe = self._bev_pos_embed(i, geom, W, H, dev)
            if ops.tma_ok(st["vproj"][0], hc, st["vproj"][0].shape[-1]):
                # value + positional embedding written once in split form; the 3 layers' value_proj = one TMA-fed GEMM
                vproj = ops.linear(ops.add_bcast_rows_split(ms_value, bev_pe), st["vproj"][0], st["vproj"][1])     # :886
            else:
                valin = torch.empty_like(ms_value)
                ops.add_bcast_rows(ms_value, bev_pe, valin)                                # :886
                vproj = ops.linear(valin.view(B * geom.n_tokens, hc), st["vproj"][0], st["vproj"][1])
            vproj = vproj.view(B, geom.n_tokens, -1)
            ops.mark(f"pos+value_proj{i}")
            if self.roi_feats and prev is not None:                                        # :890-922
                g = self.roi_feats
                K_roi = self.n_levels * g * g * hc
                if ops.tma_ok(pk["roi"][0][0], K_roi, pk["roi"][0][0].shape[-1]):
                    roi = ops.roi_sample_split(prev, ms_value, geom, hc, g, self.roi_expand_ratio[i], self.bbox_coder.cell,
                                               self.bbox_coder.pc_range, self.roi_range, B, nq)
                else:
                    roi = torch.empty((B * nq, K_roi), dtype=torch.float32, device=dev)
                    ops.roi_sample(prev, ms_value, geom, hc, g, self.roi_expand_ratio[i], self.bbox_coder.cell,
                                   self.bbox_coder.pc_range, self.roi_range, roi, B, nq)
                h1 = ops.linear(roi, pk["roi"][0][0], pk["roi"][0][1], act=ACT_RELU)
                h2 = ops.linear(h1, pk["roi"][1][0], pk["roi"][1][1], act=ACT_RELU)
                x = ops.linear(h2, pk["roi"][2][0], pk["roi"][2][1], act=ACT_RELU, res=x, res_after_act=True)
                ops.mark(f"roi{i}")
            fused = st.get("fused") if ops.FUSED_DECODER else None
            if fused is not None:
                # the stage's layers and prediction heads in one launch: query block resident in shared memory
                x, pred = ops.decoder_stage(x, qpe, q_pos, W, H, vproj, geom, self.n_points, fused, B, nq, st["pred_dim"])
            for j in range(self.n_layers if fused is None else 0):                         # [upstream] decoder layer
                lay = st["layers"][j]
                qk = ops.linear(x, lay["qk"][0], lay["qk"][1], x2=qpe)
                v = ops.linear(x, lay["v"][0], lay["v"][1])
                att = torch.empty((B * nq, hc), dtype=torch.float32, device=dev)
                ops.mha_core(qk[:, :hc], qk[:, hc:], v, att, B, nq, heads, d)
                y = ops.linear(att, lay["o"][0], lay["o"][1], res=x)
                x1 = ops.layernorm(y, *lay["ln"][0])
                oa = ops.linear(x1, lay["oa"][0], lay["oa"][1], x2=qpe)
                samp = torch.empty((B * nq, hc), dtype=torch.float32, device=dev)
                ops.msda(vproj, j * hc, geom, self.n_points, q_pos, W, H, oa[:, :n_off], oa[:, n_off:], samp, B, nq, heads, d)
                y = ops.linear(samp, lay["op"][0], lay["op"][1], res=x1)
                x2_ = ops.layernorm(y, *lay["ln"][1])
                f = ops.linear(x2_, lay["f1"][0], lay["f1"][1], act=ACT_RELU)
                y = ops.linear(f, lay["f2"][0], lay["f2"][1], res=x2_)
                x = ops.layernorm(y, *lay["ln"][2])
            if fused is None:
                hh = ops.linear(x, st["head1"][0], st["head1"][1], act=ACT_RELU)           # :939
                pred = ops.linear(hh, st["head2"][0], st["head2"][1], cout=st["pred_dim"])
            if self.classaware_reg:                                                        # :940-943
                gk = [self.common_heads[n][0] for n in self.common_heads]
                pred = ops.class_select(pred, q_label, gk, nc, nc, _pad4(sum(gk) + nc))
            q_pos = q_pos.clone()
            ops.head_update(pred, q_pos, prev if (self.roi_based_reg and prev is not None) else None)   # :945-957
            preds.append(pred)
            prev = pred
            ops.mark(f"decoder{i}")
        self._state = dict(B=B, nq=nq, q_score=q_score, q_label=q_label, last=preds[-1])
        return self._pack(preds, q_score, q_label, dense_heatmaps, nms_heats, tops, B, nq, q_feat)

    def _pack(self, preds, q_score, q_label, dense_heatmaps, nms_heats, tops, B, nq, q_feat0):
        """Reference output layout (:960-992): per-key [B, k, n_stage*nq] channel-major tensors."""
        cols, c0 = {}, 0
        for n, (kdim, _) in self.common_heads.items():
            cols[n] = (c0, c0 + kdim)
            c0 += kdim
        cols["heatmap"] = (c0, c0 + self.num_classes)
        res = {}
        for n, (a, b) in cols.items():
            res[n] = torch.cat([p[:, a:b].reshape(B, nq, b - a).transpose(1, 2) for p in preds], dim=-1)
        res["query_heatmap_score"] = q_score.view(B, nq, -1).transpose(1, 2)
        nc = self.num_classes
        res["dense_heatmap"] = [l[..., :nc].permute(0, 3, 1, 2) for l in dense_heatmaps]
        res["query_labels"] = q_label.view(B, nq).long()
        res["_nms_heatmap"] = nms_heats
        res["_top_proposals"] = tops
        res["_query_feat0"] = q_feat0.view(B, nq, -1)
        self._cls_col = cols["heatmap"][0]
        return res

    def get_bboxes(self):
        """get_bboxes (:1313-1413) with nms_type=None, generalised over the batch.  Returns device tensors:
        boxes [B,nq,code-1], scores [B,nq], labels [B,nq], keep [B,nq]."""
        st = self._state
        B, nq = st["B"], st["nq"]
        dev = st["last"].device
        code = 9 if self.has_vel else 7
        boxes = torch.empty((B * nq, code), dtype=torch.float32, device=dev)
        scores = torch.empty((B * nq,), dtype=torch.float32, device=dev)
        labels = torch.empty((B * nq,), dtype=torch.int32, device=dev)
        keep = torch.empty((B * nq,), dtype=torch.uint8, device=dev)
        bc = self.bbox_coder
        ops.box_decode(st["last"], self._cls_col, self.has_vel, st["q_score"], st["q_label"], self.num_classes, bc.cell,
                       bc.pc_range, bc.post_center_range, boxes, scores, labels, keep)
        boxes, scores, labels, keep = boxes.view(B, nq, code), scores.view(B, nq), labels.view(B, nq), keep.view(B, nq)
        if self.nms_type is not None:                                              # focal_decoder.py:1333-1385
            if self.test_cfg["dataset"] == "nuScenes":
                tasks = [(list(range(8)), -1.0), ([8], 0.175), ([9], 0.175)]
            elif self.test_cfg["dataset"] == "Waymo":
                tasks = [([0], 0.7), ([1], 0.7), ([2], 0.7)]
            else:
                raise NotImplementedError("FocalDecoder.get_bboxes: NMS tasks are defined for nuScenes and Waymo only")
            keep = ops.nms_tasks(boxes, scores, labels, keep, tasks, self.nms_type, self.test_cfg.get("pre_maxsize"),
                                 self.test_cfg.get("post_maxsize"))
        return boxes, scores, labels, keep


@DETECTORS.register_module()
class FocalFormer3D(nn.Module):
    """models/detectors/focalformer3d.py:27 -- inference forward (simple_test :321-332) on libff3d.so: LiDAR-only,
    camera-only (input_pts=False) and LiDAR + camera configs."""

    def __init__(self, pts_voxel_layer=None, pts_voxel_encoder=None, pts_middle_encoder=None, pts_backbone=None,
                 pts_neck=None, imgpts_neck=None, pts_bbox_head=None, train_cfg=None, test_cfg=None, input_img=True,
                 input_pts=True, img_backbone=None, img_neck=None, **unused):
        super().__init__()
        self.input_img, self.input_pts = bool(input_img), bool(input_pts)
        cfg = dict(pts_voxel_layer=pts_voxel_layer, pts_voxel_encoder=pts_voxel_encoder,
                   pts_middle_encoder=pts_middle_encoder, pts_backbone=pts_backbone, pts_neck=pts_neck,
                   imgpts_neck=imgpts_neck, pts_bbox_head=pts_bbox_head, test_cfg=test_cfg, input_img=input_img,
                   input_pts=input_pts, img_backbone=img_backbone, img_neck=img_neck)
        spec = param_spec(cfg)
        tcfg = test_cfg["pts"] if (test_cfg and "pts" in test_cfg) else test_cfg
        self._prepared_on = None
        if self.input_img:
            # image tower (focalformer3d.py:133-153); camera-only (DeformFormer3D_C_R50) builds no LiDAR tower
            self.img_backbone = BACKBONES.build(img_backbone, spec=sub_spec(spec, "img_backbone"))
            self.img_neck = NECKS.build(img_neck, spec=sub_spec(spec, "img_neck"))
        if self.input_img and not self.input_pts:
            self.imgpts_neck = NECKS.build(imgpts_neck, spec=sub_spec(spec, "imgpts_neck"), input_img=True)
            self.pts_bbox_head = HEADS.build(pts_bbox_head, spec=sub_spec(spec, "pts_bbox_head"), test_cfg=tcfg)
            return
        self.voxel_cfg = dict(pts_voxel_layer)
        self.pts_voxel_encoder = VOXEL_ENCODERS.build(pts_voxel_encoder, spec=sub_spec(spec, "pts_voxel_encoder"))
        self.pts_middle_encoder = MIDDLE_ENCODERS.build(pts_middle_encoder, spec=sub_spec(spec, "pts_middle_encoder"))
        depth = self._bev_depth(self.pts_middle_encoder)
        self.pts_backbone = BACKBONES.build(pts_backbone, spec=sub_spec(spec, "pts_backbone"), in_depth=depth)
        self.pts_neck = NECKS.build(pts_neck, spec=sub_spec(spec, "pts_neck"))
        self.imgpts_neck = NECKS.build(imgpts_neck, spec=sub_spec(spec, "imgpts_neck"))
        self.pts_bbox_head = HEADS.build(pts_bbox_head, spec=sub_spec(spec, "pts_bbox_head"), test_cfg=tcfg)

    @staticmethod
    def _bev_depth(me):
        d = me.sparse_shape[0]
        for i, blocks in enumerate(me.encoder_channels[:-1]):
            pad = me.encoder_paddings[i][len(blocks) - 1]
            pz = pad[0] if isinstance(pad, (list, tuple)) else pad
            d = (d + 2 * pz - 3) // 2 + 1
        return (d - 3) // 2 + 1

    def prepare(self, device="cuda"):
        """Fold BatchNorm, pack weights into kernel layouts on the device (call after load_state_dict)."""
        dev = torch.device(device)
        for m in self.children():
            if hasattr(m, "prepare"):
                m.prepare(dev)
        self._prepared_on = dev
        return self

    # ---- the hot path
    @torch.no_grad()
    def forward_camera(self, img, img_metas, keep_stages=False):
        """img [B, N, 3, H, W] float32 (already normalised, as the test pipeline hands it to simple_test);
        img_metas: list[B] of dict(lidar2img=[N, 4, 4]).  focalformer3d.py:133-153,186 + focal_encoder.py:171-197."""
        dev = self._prepared_on
        B, N, Cc, H, W = img.shape
        ops.gemm_flag(dev).zero_()
        ops.mark("start")
        x = ops.nchw_to_nhwc(img.to(dev, torch.float32).reshape(B * N, Cc, H, W).contiguous(), 8)
        feats = self.img_backbone(x)
        ops.mark("img_backbone")
        f0 = self.img_neck(feats)
        ops.mark("img_neck")
        head = self.pts_bbox_head
        lss = self.imgpts_neck.pk["lss"]
        nx, ny, _ = lss.nxyz
        geom = ops.LevelGeom([(ny >> l, nx >> l) for l in range(head.n_levels)])
        ms_value = torch.empty((B, geom.n_tokens, head.hc), dtype=torch.float32, device=dev)
        bev_feat = torch.empty((B, ny, nx, head.hc), dtype=torch.float32, device=dev)
        _, dn, bev = self.imgpts_neck.forward_camera(f0, img_metas, bev_feat)
        # the same tensor is the heatmap input and level 0 of the decoder's value pyramid (focal_encoder.py:196-197)
        ms_value[:, :ny * nx].view(B, ny, nx, head.hc).copy_(bev_feat)
        ops.mark("lift_splat+bevencode")
        res = head(bev_feat, [], ms_value, geom)
        det = head.get_bboxes()
        ops.mark("heads+decode")
        self._overflow = torch.zeros((1,), dtype=torch.int32, device=dev)
        stages = None
        if keep_stages:
            stages = dict(img_nhwc=x, backbone=feats, img_feat=f0, depthnet=dn, bev=bev, conv_feat=bev_feat,
                          ms_value=ms_value)
        return res, det, stages

    @torch.no_grad()
    def forward_raw(self, points, keep_stages=False, img=None, img_metas=None):
        """points: list[B] of CUDA float32 [Ni, F].  Returns (head dict, (boxes, scores, labels, keep), stages)."""
        if self._prepared_on is None:
            raise RuntimeError("call model.prepare(device) after loading weights")
        if self.input_img:
            if img is None or img_metas is None:
                raise ValueError("camera config: forward_raw needs img [B,N,3,H,W] and img_metas with 'lidar2img'")
            if not self.input_pts:
                return self.forward_camera(img, img_metas, keep_stages)
        dev = self._prepared_on
        if isinstance(points, tuple):
            # (concatenated [sum N_i, F] device tensor, host row offsets [B+1]): the CUDA-graph runner's static input buffer
            allp, offs = points
            B = len(offs) - 1
            sizes = [offs[b + 1] - offs[b] for b in range(B)]
        else:
            B = len(points)
            offs = [0]
            for p in points:
                offs.append(offs[-1] + int(p.shape[0]))
            sizes = [int(p.shape[0]) for p in points]
            allp = torch.cat([p.to(dev, torch.float32) for p in points], 0).contiguous() if B > 1 else points[0].to(dev, torch.float32).contiguous()
        vc = self.voxel_cfg
        mv = vc["max_voxels"]
        mv = mv[1] if isinstance(mv, (tuple, list)) else mv
        n_max = max(max(sizes), 1)
        # focalformer3d.py:80,159-163: a 'Dynamic*' voxel encoder switches to dynamic voxelisation (no caps)
        dynamic = isinstance(self.pts_voxel_encoder, DynamicSimpleVFE)
        mv = n_max if (dynamic or mv <= 0) else min(mv, n_max)
        ops.gemm_flag(dev).zero_()
        ops.mark("start")
        hard_vfe = isinstance(self.pts_voxel_encoder, HardVFE)
        vox = ops.voxelize(allp, offs, vc["voxel_size"], vc["point_cloud_range"], -1 if dynamic else vc["max_num_points"], mv,
                           mean_ld=8, want_voxels=(keep_stages and not dynamic) or hard_vfe)
        if hard_vfe:
            vox["mean"] = self.pts_voxel_encoder(vox, vc["max_num_points"])      # [cap, 64] learned voxel features
        ops.mark("voxelize+vfe")
        me = self.pts_middle_encoder
        depth = self._bev_depth(me)
        H, W = me.sparse_shape[1] // 8, me.sparse_shape[2] // 8
        bev = torch.zeros((B, H, W, depth * me.output_channels), dtype=torch.float32, device=dev)
        overflow = torch.zeros((1,), dtype=torch.int32, device=dev)
        me(vox, B, bev, overflow)
        ops.mark("sparse_encoder")
        xs = self.pts_backbone(bev)
        ops.mark("second")
        c_neck = sum(self.pts_neck.out_channels)
        neck = torch.empty((B, H, W, c_neck), dtype=torch.float32, device=dev)
        neck_s = ops.Split.empty((B, H, W), c_neck, dev) if (ops.tma_enabled() and c_neck % 64 == 0) else None
        self.pts_neck(xs, neck, neck_s)
        ops.mark("secondfpn")
        head = self.pts_bbox_head
        geom = ops.LevelGeom([(H >> l, W >> l) for l in range(head.n_levels)])
        ms_value = torch.empty((B, geom.n_tokens, head.hc), dtype=torch.float32, device=dev)
        extra_view = ms_value[:, :H * W].view(B, H, W, head.hc)
        cam = None
        if self.input_img:
            # LiDAR + camera (FocalFormer3D_LC): image tower (focalformer3d.py:133-153), then the fused encoder
            Bi, N, Cc, Hi, Wi = img.shape
            x = ops.nchw_to_nhwc(img.to(dev, torch.float32).reshape(Bi * N, Cc, Hi, Wi).contiguous(), 8)
            feats = self.img_backbone(x)
            ops.mark("img_backbone")
            f0 = self.img_neck(feats)
            ops.mark("img_neck")
            conv_feat, stage_feats, extra, img_bev = self.imgpts_neck.forward_fusion(neck, f0, img_metas, extra_view, neck_s=neck_s)
            cam = dict(img_backbone=feats, img_feat=f0, img_bev=img_bev)
        else:
            conv_feat, stage_feats, extra = self.imgpts_neck(neck, extra_out=extra_view, neck_s=neck_s)
        ops.mark("focal_encoder")
        res = head(conv_feat, stage_feats, ms_value, geom, split_feats=getattr(self.imgpts_neck, "split_feats", None))
        det = head.get_bboxes()
        ops.mark("heads+decode")
        self._overflow = overflow
        stages = None
        if keep_stages:
            stages = dict(vox=vox, bev=bev, backbone=[a for a, _ in xs], neck=neck, conv_feat=conv_feat, stage_feats=stage_feats,
                          extra=extra, ms_value=ms_value, overflow=overflow, level_sizes=me.level_sizes, cam=cam)
        return res, det, stages

    def check_flags(self):
        """Device-side failure flags of the last forward_raw (one small D2H copy): sparse-level capacity overflow and
        fp16-range saturation of a GEMM operand.  Either one invalidates the result: fail loudly."""
        flags = torch.cat([self._overflow.view(-1)[:1], ops.gemm_flag(self._prepared_on).view(-1)]).cpu().tolist()
        if flags[0]:
            raise RuntimeError("sparse encoder level capacity exceeded; raise SparseEncoder.cap_growth")
        if flags[1]:
            raise RuntimeError("an activation left the fp16 range (|a| > 65504) in the fp16 hi/lo GEMM: the result is "
                               "invalid; rerun with FF3D_GEMM=tf32 (TF32 hi/lo operand split, no range limit)")

    def simple_test(self, points, img_metas=None, img=None, rescale=False):
        """Reference signature (focalformer3d.py:321): list of dict(pts_bbox=dict(boxes_3d, scores_3d, labels_3d)) on CPU."""
        _, (boxes, scores, labels, keep), _ = self.forward_raw(points, img=img, img_metas=img_metas)
        self.check_flags()
        out = []
        for b in range(boxes.shape[0]):
            m = keep[b].bool()
            bx, sc, lb = boxes[b][m], scores[b][m], labels[b][m]
            if bx.shape[0] > 200:                                     # focal_decoder.py:1395-1400
                inds = sc.argsort(descending=True, stable=True)[:200]
                bx, sc, lb = bx[inds], sc[inds], lb[inds]
            out.append(dict(pts_bbox=dict(boxes_3d=bx.cpu(), scores_3d=sc.cpu(), labels_3d=lb.cpu())))
        return out

    def forward(self, return_loss=False, rescale=True, points=None, img_metas=None, img=None, **kw):
        """[upstream] Base3DDetector.forward -> forward_test -> simple_test(points[0], img_metas[0], img[0])."""
        if return_loss:
            raise NotImplementedError("training is outside the hot path built here (SURVEY.md 8f row 4)")
        return self.simple_test(points[0], img_metas[0] if img_metas else None, img[0] if img else None, rescale)


def build_model(model_cfg, test_cfg=None):
    """tools/test.py:202-203 build_model(cfg.model, test_cfg=cfg.get('test_cfg'))."""
    cfg = dict(model_cfg)
    if test_cfg is not None and cfg.get("test_cfg") is None:
        cfg["test_cfg"] = test_cfg
    return DETECTORS.build(cfg)
