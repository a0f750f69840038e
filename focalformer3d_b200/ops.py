"""Operator-level host API over the C ABI (include/ff3d.h).  torch is used for device memory and the stream only.

Feature maps are NHWC views: a tensor whose last dim is contiguous; the row stride (``ld``) and batch stride are
read from the view, so channel slices of wider buffers (concat-free ``torch.cat`` replacements) work directly.
"""
import ctypes as C
import torch

from . import lib as L
from .lib import lib, check, GemmDesc, ACT_NONE, ACT_RELU, ACT_RELU6, GEMM_ROWS, GEMM_CONV2D, GEMM_SPARSE  # noqa: F401

launch_count = 0   # number of ff3d kernel-launching C calls (bench.py reports kernel launches from this)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _chk_f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise L.Ff3dError(f"{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    if t.stride(-1) != 1:
        raise L.Ff3dError(f"{name}: last dim must be contiguous")


def _count(n=1):
    global launch_count
    launch_count += n


class Profiler:
    """Optional CUDA-event instrumentation (off in the timed benchmark region): per-op (label, ms, flops, bytes)
    records and stage markers for the per-stage ms the headline metric asks for."""

    def __init__(self):
        self.enabled = False
        self.keep_records = True
        self.records = []      # (label, start, end, flops, bytes, n_dev_or_None, meta)
        self.marks = []        # (name, event)

    def start(self, records=True):
        self.enabled, self.keep_records, self.records, self.marks = True, records, [], []

    def stop(self):
        self.enabled = False

    def mark(self, name):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def begin(self):
        if not self.enabled or not self.keep_records:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, start, label, flops, nbytes, n_dev=None, meta=None):
        if start is None:
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.records.append((label, start, e, flops, nbytes, n_dev, meta))

    def stage_ms(self):
        out = {}
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out


prof = Profiler()
mark = prof.mark

# FF3D_TC=0 forces the fp32 SIMT kernel everywhere (debugging aid; both are our own sm_100a kernels)
import os as _os
USE_TC = _os.environ.get("FF3D_TC", "1") != "0"


def tc_weight_images(w, bn=None):
    """[taps, cin, cout] fp32 (cin in {8,16} or a multiple of 32) -> [n_tiles, n_stages, 2, bn, 32] fp32: per
    (N tile, pipeline K-step) the hi and lo TF32 parts of the weights as 128B-swizzled K-major smem images
    (row n = output channel, 16-byte chunk j stored at chunk j ^ (n % 8)): one cp.async.bulk per stage."""
    taps, cin, cout = w.shape
    default_bn = lib.ff3d_tcgemm_ntile(cin, cout)
    if default_bn <= 0:
        return None, 0
    bn = bn or default_bn
    assert cout % bn == 0 and bn in (16, 32, 64, 128)
    n_stages = lib.ff3d_tcgemm_stages(cin, taps)
    if cin >= 32:
        kmat = w.reshape(taps * cin, cout)                                    # stage s = rows [32 s, 32 s + 32)
    else:
        tps = 32 // cin
        kmat = torch.zeros((n_stages * tps, cin, cout), dtype=torch.float32)
        kmat[:taps] = w
        kmat = kmat.reshape(n_stages * 32, cout)
    def rn_tf32(t):          # cvt.rna.tf32.f32: round the magnitude to 10 mantissa bits, ties away from zero
        return ((t.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    hi = rn_tf32(kmat)
    lo = rn_tf32(kmat - hi)
    n_tiles = cout // bn
    n_idx = torch.arange(bn)
    j_idx = torch.arange(8)
    dst_chunk = (j_idx[None, :] ^ (n_idx[:, None] & 7))                       # [bn, 8]
    imgs = torch.empty((n_tiles, n_stages, 2, bn, 32), dtype=torch.float32)
    for part, src in enumerate((hi, lo)):
        blk = src.view(n_stages, 8, 4, n_tiles, bn).permute(3, 0, 4, 1, 2)    # [tile, stage, n, j, e]
        out = torch.empty((n_tiles, n_stages, bn, 8, 4), dtype=torch.float32)
        out.scatter_(3, dst_chunk[None, None, :, :, None].expand(n_tiles, n_stages, bn, 8, 4), blk)
        imgs[:, :, part] = out.view(n_tiles, n_stages, bn, 32)
    return imgs.contiguous(), bn


LO_SCALE = 2048.0          # fp16 split: w = hi + lo / 2048


def split_f16(t):
    """fp32 tensor -> (hi, lo) fp16 tensors with t ~= hi + lo / 2048 (saturating at the fp16 range)."""
    hi = t.clamp(-65504.0, 65504.0).half()
    lo = ((t - hi.float()) * LO_SCALE).clamp(-65504.0, 65504.0).half()
    return hi, lo


def tc_weight_images_f16(w, bn=None):
    """[taps, cin, cout] fp32 (cin in {8,16,32} or a multiple of 64) -> ([n_tiles, n_stages, 2, bn, 64] fp16, bn): per
    (N tile, pipeline K-step of 64) the hi and lo * 2^11 fp16 parts of the weights as 128B-swizzled K-major smem images
    (row n = output channel, 16-byte chunk j = 8 halves stored at chunk j ^ (n % 8)): one cp.async.bulk per stage."""
    taps, cin, cout = w.shape
    default_bn = lib.ff3d_tcgemm_f16_ntile(cin, cout)
    if default_bn <= 0:
        return None, 0
    bn = bn or default_bn
    assert cout % bn == 0 and bn in (16, 32, 64, 128)
    n_stages = lib.ff3d_tcgemm_f16_stages(cin, taps)
    if cin >= 64:
        kmat = w.reshape(taps * cin, cout)                          # stage s = rows [64 s, 64 s + 64)
    else:
        tps = 64 // cin
        kmat = torch.zeros((n_stages * tps, cin, cout), dtype=torch.float32)
        kmat[:taps] = w
        kmat = kmat.reshape(n_stages * 64, cout)
    hi, lo = split_f16(kmat.float())
    n_tiles = cout // bn
    n_idx, j_idx = torch.arange(bn), torch.arange(8)
    dst_chunk = j_idx[None, :] ^ (n_idx[:, None] & 7)              # [bn, 8]
    imgs = torch.empty((n_tiles, n_stages, 2, bn, 64), dtype=torch.float16)
    for part, src in enumerate((hi, lo)):
        blk = src.view(n_stages, 8, 8, n_tiles, bn).permute(3, 0, 4, 1, 2)      # [tile, stage, n, chunk j, 8 halves]
        out = torch.empty((n_tiles, n_stages, bn, 8, 8), dtype=torch.float16)
        out.scatter_(3, dst_chunk[None, None, :, :, None].expand(n_tiles, n_stages, bn, 8, 8), blk)
        imgs[:, :, part] = out.view(n_tiles, n_stages, bn, 64)
    return imgs.contiguous(), bn


def unpack_images_f16(imgs, taps, cin, cout):
    """Inverse of tc_weight_images_f16 (test helper): -> (hi, lo) as [taps, cin, cout] fp32."""
    n_tiles, n_stages, _, bn, _ = imgs.shape
    n_idx, j_idx = torch.arange(bn), torch.arange(8)
    src_chunk = j_idx[None, :] ^ (n_idx[:, None] & 7)
    parts = []
    for part in range(2):
        sw = imgs[:, :, part].reshape(n_tiles, n_stages, bn, 8, 8)
        un = sw.gather(3, src_chunk[None, None, :, :, None].expand(n_tiles, n_stages, bn, 8, 8))
        kmat = un.permute(1, 3, 4, 0, 2).reshape(n_stages * 64, n_tiles * bn).float()       # [K, cout]
        parts.append(kmat.reshape(taps, cin, cout) if cin >= 64 else kmat.reshape(-1, cin, cout)[:taps])
    return parts[0], parts[1]


# Operand format of the tensor-core GEMM: "f16" = fp16 hi/lo split on kind::f16 (default: twice the MMA rate, half the
# shared-memory bytes per product, and -- measured at full size against an fp64 run of the oracle -- HALF the error of
# the TF32 split, because the accumulator is truncated once per 16 instead of once per 8 products); "tf32" = TF32 hi/lo
# split (no range limit).  Activations beyond +-65504 raise a device flag in f16 mode (gemm_flag) and fail the call.
GEMM_KIND = _os.environ.get("FF3D_GEMM", "f16")
if GEMM_KIND not in ("f16", "tf32"):
    raise L.Ff3dError(f"FF3D_GEMM={GEMM_KIND!r}: expected 'f16' or 'tf32'")

_flags = {}


def gemm_flag(dev=None):
    """Device int32[1] raised by the f16 kernels when an activation saturated the fp16 range."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else torch.device(dev)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _flags:
        _flags[key] = torch.zeros((1,), dtype=torch.int32, device=dev)
    return _flags[key]


class PackedW:
    """Device weights of one GEMM-like layer: ``w`` [taps, cin, ldw] (SIMT kernel) and the tcgen05 kernel's pre-swizzled
    hi/lo images in the selected operand format (``kind`` 'f16' / 'tf32'; None = not tensor-core tileable)."""

    def __init__(self, w_cpu, dev, bn=None, kind=None):
        self.w = w_cpu.to(dev)
        kind = kind or GEMM_KIND
        img, self.bn, self.kind = None, 0, None
        if kind == "f16":
            img, self.bn = tc_weight_images_f16(w_cpu, bn)
            if img is not None:
                self.kind = "f16"
        if self.kind is None:                                   # 'tf32' requested, or a cin only the 32-wide K step tiles
            img, self.bn = tc_weight_images(w_cpu, bn)
            if img is not None:
                self.kind = "tf32"
        self.img = img.to(dev) if img is not None else None

    @property
    def shape(self):
        return self.w.shape


# rows below which a plain linear layer goes to the fp32 SIMT kernel: a persistent tcgen05 CTA pays ~10 us of fixed
# cost (TMEM alloc, barrier init, pipeline fill) that a [<= few thousand rows] x 128 GEMM never amortises
TC_MIN_ROWS = int(_os.environ.get("FF3D_TC_MIN_ROWS", "0"))


def _gemm(d, w, what):
    """Dispatch one implicit-GEMM launch: tcgen05 split-precision kernel when the layer is tensor-core tileable."""
    small = d.mode == GEMM_ROWS and d.M < TC_MIN_ROWS and d.cin <= 1024
    if USE_TC and w.img is not None and not small:
        if w.kind == "f16":
            check(lib.ff3d_tcgemm_f16(C.byref(d), _ptr(w.img), w.bn, _ptr(gemm_flag(w.img.device)), _stream()),
                  f"ff3d_tcgemm_f16({what})")
        else:
            check(lib.ff3d_tcgemm_bn(C.byref(d), _ptr(w.img), w.bn, _stream()), f"ff3d_tcgemm({what})")
    else:
        check(lib.ff3d_igemm(C.byref(d), _stream()), f"ff3d_igemm({what})")


class Split:
    """Activation rows in split form for the TMA-fed GEMM (csrc/tmagemm.cu): ``t`` [..., 2*C] fp16, channels [0, C) = hi,
    [C, 2C) = lo * 2^11 (value = hi + lo / 2048) -- 4C bytes per row, the footprint of fp32.  ``c0:c1`` selects a channel
    slice (concat-free views); ``zero_row`` = index of an all-zero row kept after the last data row (sparse levels)."""

    def __init__(self, t, C, c0=0, c1=None, zero_row=-1):
        self.t, self.C, self.c0, self.c1, self.zero_row = t, C, c0, (C if c1 is None else c1), zero_row

    @staticmethod
    def empty(shape, C, dev, zero_row=False):
        """shape = leading dims (rows,) or (B, H, W); with zero_row one extra all-zero row is appended to a 2-D buffer."""
        shape = tuple(shape)
        if zero_row:
            assert len(shape) == 1
            t = torch.empty((shape[0] + 1, 2 * C), dtype=torch.float16, device=dev)
            if GATHER4:                        # only the tile::gather4 engine reads it (cp.async zero-fills in flight)
                t[shape[0]].zero_()
            return Split(t, C, zero_row=shape[0])
        return Split(torch.empty(shape + (2 * C,), dtype=torch.float16, device=dev), C)

    def slice(self, c0, c1):
        return Split(self.t, self.C, self.c0 + c0, self.c0 + c1, self.zero_row)

    @property
    def channels(self):
        return self.c1 - self.c0

    @property
    def ptr(self):
        return self.t.data_ptr() + 2 * self.c0

    @property
    def ld(self):
        return 2 * self.C


def split_rows(x, n_dev=None, zero_row=False, out=None):
    """fp32 rows [..., C] (last dim contiguous, uniform row stride) -> Split of the same leading shape."""
    _chk_f32(x, "split_rows.x")
    Cc = x.shape[-1]
    lead = tuple(x.shape[:-1])
    rows = 1
    for v in lead:
        rows *= v
    ldx = x.stride(-2) if x.dim() > 1 else Cc
    if out is None:
        out = Split.empty((rows,), Cc, x.device, zero_row=True) if zero_row else Split.empty(lead, Cc, x.device)
    check(lib.ff3d_split_rows(_ptr(x), ldx, _ptr(n_dev), rows, Cc, C.c_void_p(out.ptr), out.ld, out.C, _ptr(gemm_flag(x.device)),
                              _stream()), "ff3d_split_rows")
    _count()
    return out


def unsplit_rows(xs, rows, n_dev=None):
    """Split (2-D, or any leading shape flattened to ``rows`` rows) -> fp32 [rows, channels]."""
    out = torch.empty((rows, xs.channels), dtype=torch.float32, device=xs.t.device)
    check(lib.ff3d_unsplit_rows(C.c_void_p(xs.ptr), xs.ld, xs.C, _ptr(n_dev), rows, xs.channels, _ptr(out), out.stride(0),
                                _stream()), "ff3d_unsplit_rows")
    _count()
    return out


USE_TMA = _os.environ.get("FF3D_TMA", "1") != "0"     # FF3D_TMA=0: never take the TMA-fed kernel (debugging aid)
GATHER4 = _os.environ.get("FF3D_SPARSE_GATHER", "cpasync") == "tma"    # sparse rows by TMA tile::gather4 instead of cp.async


def _attach_split(d, xs=None, res_s=None, out_s=None, xs_rows=0):
    if xs is not None:
        d.xs, d.ldxs, d.xs_lo, d.xs_rows, d.zero_row = xs.ptr, xs.ld, xs.C, xs_rows, xs.zero_row
    if res_s is not None:
        d.res_s, d.ldres_s, d.res_s_lo = res_s.ptr, res_s.ld, res_s.C
    if out_s is not None:
        d.ys, d.ldys, d.ys_lo = out_s.ptr, out_s.ld, out_s.C


def _gemm_any(d, w, what, x_is_split):
    """TMA-fed kernel when the A rows are in split form and the layer qualifies, else the register-path kernels."""
    if x_is_split:
        if not (USE_TC and USE_TMA and w.kind == "f16" and lib.ff3d_tmagemm_supported(C.byref(d))):
            raise L.Ff3dError(f"{what}: split input given but the layer is not TMA-tileable (cin={d.cin} cout={d.cout})")
        bn = w.bn if w.bn in (16, 32, 64, 128) else 0
        check(lib.ff3d_tmagemm(C.byref(d), _ptr(w.img), bn, _ptr(gemm_flag(w.img.device)), _stream()), f"ff3d_tmagemm({what})")
    else:
        _gemm(d, w, what)


def tma_enabled():
    """Is the TMA-fed kernel (split activations) in use at all?  (fp16 operand format, not disabled by FF3D_TMA / FF3D_TC)"""
    return bool(USE_TC and USE_TMA and GEMM_KIND == "f16")


def tma_ok(w, cin, cout, sparse=False):
    """Can a layer with these weights take split A rows (ff3d_tmagemm)?  Dense layers: cin a multiple of 64, cout 64 or a
    multiple of 128; sparse layers (cp.async gather) also the narrow levels cin 8 / 16 / 32, cout 16 / 32."""
    if not (USE_TC and USE_TMA and w.img is not None and w.kind == "f16"):
        return False
    wide = cin % 64 == 0 and cin >= 64 and (cout % 128 == 0 or cout in (16, 32, 64)) and w.bn in (16, 32, 64, 128)
    if sparse:
        return bool(wide or ((cin in (8, 16, 32) or (cin % 64 == 0 and cin >= 64))
                             and (cout in (16, 32, 64) or cout % 128 == 0) and w.bn in (16, 32, 64, 128)))
    return bool(wide)


# --------------------------------------------------------------------------------------------------------------
def linear(x, w, bias=None, out=None, act=ACT_NONE, res=None, x2=None, cout=None, res_after_act=False, out_s=None,
           want_out=True):
    """y = act(x @ w + bias (+ res)).  x [M, cin] fp32 row view (ld from stride) or a Split; w packed [1, cin, ldw].
    ``out_s``: Split that also receives the output rows; ``want_out=False`` skips the fp32 output (TMA kernel only)."""
    split_in = isinstance(x, Split)
    if split_in:
        M, cin = x.t.shape[0] - (1 if x.zero_row >= 0 else 0), x.channels
        dev = x.t.device
    else:
        _chk_f32(x, "linear.x")
        M, cin = x.shape
        dev = x.device
    ldw = w.shape[-1]
    cout = cout or ldw
    if out is None and want_out:
        out = torch.empty((M, cout), device=dev, dtype=torch.float32)
    d = GemmDesc()
    d.mode, d.M, d.cin, d.cout, d.taps = GEMM_ROWS, M, cin, cout, 1
    if split_in:
        if x2 is not None:
            raise L.Ff3dError("linear: x2 is not supported with a split input")
        _attach_split(d, xs=x, xs_rows=M)
    else:
        d.x, d.ldx, d.x2 = x.data_ptr(), x.stride(0), (x2.data_ptr() if x2 is not None else None)
        if x2 is not None and x2.stride(0) != x.stride(0):
            raise L.Ff3dError("linear: x2 must share x's row stride")
    d.w, d.ldw, d.bias = w.w.data_ptr(), ldw, (bias.data_ptr() if bias is not None else None)
    if isinstance(res, Split):
        _attach_split(d, res_s=res)
    else:
        d.res, d.ldres = (res.data_ptr(), res.stride(0)) if res is not None else (None, 0)
    if out is not None:
        d.y, d.ldy = out.data_ptr(), out.stride(0)
    d.act = act
    d.res_after_act = 1 if res_after_act else 0
    _attach_split(d, out_s=out_s)
    t0 = prof.begin()
    _gemm_any(d, w, "rows", split_in)
    prof.end(t0, f"linear[{cin}x{cout}]", 2.0 * M * cin * cout, 4.0 * (M * cin + cin * cout + M * cout))
    _count()
    return out


def _nhwc_geom(t, name):
    _chk_f32(t, name)
    B, H, W, Cc = t.shape
    ld = t.stride(2)
    if t.stride(1) != W * ld or t.stride(0) % ld != 0:
        raise L.Ff3dError(f"{name}: not an NHWC view (strides {t.stride()})")
    return B, H, W, Cc, ld, t.stride(0) // ld


def _split_nhwc_geom(xs, name):
    t = xs.t
    if t.dim() != 4 or not t.is_contiguous():
        raise L.Ff3dError(f"{name}: a split NHWC activation is a contiguous [B, H, W, 2C] fp16 buffer")
    B, H, W, _ = t.shape
    return B, H, W, xs.channels, H * W


def conv2d(x, w, bias, out, k, stride=1, pad=None, act=ACT_NONE, res=None, up=None, out_s=None):
    """NHWC conv: x [B,H,W,cin] fp32 view or a Split ([B,H,W,2C] fp16), w packed [k*k, cin, ldw], out [B,Ho*u,Wo*u,cout] view
    (None with a Split input when only the split output ``out_s`` is wanted).
    ``up=(u, dy, dx)`` writes onto the (oy*u+dy, ox*u+dx) lattice (transposed conv with kernel == stride)."""
    split_in = isinstance(x, Split)
    if split_in:
        B, H, W, cin, xbs = _split_nhwc_geom(x, "conv2d.x")
        ldx = 0
    else:
        B, H, W, cin, ldx, xbs = _nhwc_geom(x, "conv2d.x")
    pad = (k // 2) if pad is None else pad
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    u, dy, dx = up if up else (1, 0, 0)
    d = GemmDesc()
    if out is not None:
        Bo, Hy, Wy, cout, ldy, ybs = _nhwc_geom(out, "conv2d.out")
        if (Bo, Hy, Wy) != (B, Ho * u, Wo * u):
            raise L.Ff3dError(f"conv2d: out shape {tuple(out.shape)} != expected {(B, Ho * u, Wo * u, cout)}")
        d.y, d.ldy = out.data_ptr(), ldy
    else:
        if out_s is None:
            raise L.Ff3dError("conv2d: no output")
        cout, ybs = out_s.channels, Ho * u * Wo * u
    d.mode, d.M, d.cin, d.cout, d.taps = GEMM_CONV2D, B * Ho * Wo, cin, cout, k * k
    if split_in:
        _attach_split(d, xs=x, xs_rows=B * H * W)
    else:
        d.x, d.ldx = x.data_ptr(), ldx
    d.w, d.ldw, d.bias = w.w.data_ptr(), w.shape[-1], (bias.data_ptr() if bias is not None else None)
    if isinstance(res, Split):
        _attach_split(d, res_s=res)
    elif res is not None:
        rB, rH, rW, rC, ldr, rbs = _nhwc_geom(res, "conv2d.res")
        if rbs != rH * rW or (rB, rH, rW) != (B, Ho, Wo) or u != 1:
            raise L.Ff3dError("conv2d: residual must be a batch-dense NHWC view of the output shape")
        d.res, d.ldres = res.data_ptr(), ldr
    d.act = act
    d.B, d.H, d.W, d.Ho, d.Wo, d.kh, d.kw, d.stride, d.pad = B, H, W, Ho, Wo, k, k, stride, pad
    d.x_bstride, d.y_bstride, d.y_row0 = xbs, ybs, 0
    d.ux, d.uy, d.dx, d.dy = u, u, dx, dy
    if out_s is not None:
        if out_s.t.dim() != 4 or tuple(out_s.t.shape[:3]) != (B, Ho * u, Wo * u) or (out is not None and ybs != Ho * u * Wo * u):
            raise L.Ff3dError("conv2d: the split output is a batch-dense [B, Ho*u, Wo*u, 2C] buffer (same rows as out)")
        _attach_split(d, out_s=out_s)
    t0 = prof.begin()
    _gemm_any(d, w, "conv2d", split_in)
    Mo = B * Ho * Wo
    prof.end(t0, f"conv{k}x{k}s{stride}[{cin}->{cout}@{Ho}]", 2.0 * Mo * k * k * cin * cout,
             4.0 * (B * H * W * cin + k * k * cin * cout + Mo * cout))
    _count()
    return out


def sparse_conv(x, rb, n_dev, w, bias, out, act=ACT_RELU, res=None, cout=None, out_s=None):
    """Rulebook gather-GEMM: x [cap_in, cin] fp32 rows or a Split (with its all-zero row), rb a Rulebook (nbr [taps, cap_out],
    per-tile tap masks, optional row map), out [cap_out, cout] fp32 rows (None: split output only) or an arbitrary buffer
    when the rulebook carries element offsets; ``out_s``: Split receiving the output rows; res: fp32 rows or a Split."""
    split_in = isinstance(x, Split)
    nbr = rb.nbr
    taps, cap = nbr.shape
    if split_in:
        if x.zero_row < 0:
            raise L.Ff3dError("sparse_conv: a split input needs its all-zero row (Split.empty(..., zero_row=True))")
        cin, x_rows = x.channels, x.t.shape[0]
    else:
        _chk_f32(x, "sparse_conv.x")
        cin, x_rows = x.shape[1], x.shape[0]
    cout = cout or (out_s.channels if out is None else w.shape[-1])
    d = GemmDesc()
    d.mode, d.M, d.m_dev, d.cin, d.cout, d.taps = GEMM_SPARSE, cap, n_dev.data_ptr(), cin, cout, taps
    if split_in:
        _attach_split(d, xs=x, xs_rows=x_rows)
    else:
        d.x, d.ldx = x.data_ptr(), x.stride(0)
    d.w, d.ldw, d.bias = w.w.data_ptr(), w.shape[-1], (bias.data_ptr() if bias is not None else None)
    if isinstance(res, Split):
        _attach_split(d, res_s=res)
    else:
        d.res, d.ldres = (res.data_ptr(), res.stride(0)) if res is not None else (None, 0)
    if out is not None:
        d.y, d.ldy = out.data_ptr(), (out.stride(0) if rb.y_off is None else 0)
    d.act = act
    d.nbr, d.nbr_stride = nbr.data_ptr(), nbr.stride(0)
    d.y_off = rb.y_off.data_ptr() if rb.y_off is not None else None
    d.y_row = rb.y_row.data_ptr() if rb.y_row is not None else None
    d.tile_mask = rb.tile_mask.data_ptr() if rb.tile_mask is not None else None
    _attach_split(d, out_s=out_s)
    t0 = prof.begin()
    _gemm_any(d, w, "sparse", split_in)
    prof.end(t0, f"spconv[{taps}t {cin}->{cout}]", None, None, n_dev, dict(nbr=nbr, cin=cin, cout=cout, taps=taps,
                                                                            x_rows=x_rows, tile_mask=rb.tile_mask))
    _count()
    return out


def dwconv3x3(x, w, bias, out, act=ACT_RELU6):
    B, H, W, Cc, ldx, xbs = _nhwc_geom(x, "dwconv.x")
    _, _, _, _, ldy, ybs = _nhwc_geom(out, "dwconv.out")
    if xbs != H * W or ybs != H * W:
        raise L.Ff3dError("dwconv3x3: batch-dense views required")
    check(lib.ff3d_dwconv3x3(_ptr(x), ldx, _ptr(w), _ptr(bias), _ptr(out), ldy, B, H, W, Cc, act, _stream()),
          "ff3d_dwconv3x3")
    _count()
    return out


def dwconv3x3_split(x, w, bias, act=ACT_RELU6):
    """dwconv3x3 whose output goes straight into split form ([B,H,W,2C] fp16) for a TMA-fed 1x1 conv."""
    B, H, W, Cc, ldx, xbs = _nhwc_geom(x, "dwconv.x")
    if xbs != H * W:
        raise L.Ff3dError("dwconv3x3_split: batch-dense view required")
    out = Split.empty((B, H, W), Cc, x.device)
    check(lib.ff3d_dwconv3x3_split(_ptr(x), ldx, _ptr(w), _ptr(bias), C.c_void_p(out.ptr), B, H, W, Cc, act,
                                   _ptr(gemm_flag(x.device)), _stream()), "ff3d_dwconv3x3_split")
    _count()
    return out


def layernorm(x, gamma, beta, out=None, eps=1e-5):
    _chk_f32(x, "layernorm.x")
    assert x.is_contiguous()
    out = torch.empty_like(x) if out is None else out
    check(lib.ff3d_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), x.shape[0], x.shape[1], eps, _stream()),
          "ff3d_layernorm")
    _count()
    return out


def local_attention(q, k, v, out, ksize):
    """9x9 local-context attention on NHWC views: out[b,y,x,:] = sum_k softmax_k(q.k_nb / sqrt(C)) v_nb."""
    B, H, W, Cc, ldq, qbs = _nhwc_geom(q, "local_attention.q")
    _, _, _, _, ldk, kbs = _nhwc_geom(k, "local_attention.k")
    _, _, _, _, ldv, vbs = _nhwc_geom(v, "local_attention.v")
    _, _, _, _, ldo, obs = _nhwc_geom(out, "local_attention.out")
    if not (qbs == kbs == vbs == obs == H * W):
        raise L.Ff3dError("local_attention: batch-dense views required")
    t0 = prof.begin()
    check(lib.ff3d_local_attention(_ptr(q), ldq, _ptr(k), ldk, _ptr(v), ldv, _ptr(out), ldo, B, H, W, Cc, ksize, _stream()),
          "ff3d_local_attention")
    prof.end(t0, f"local_attention[k{ksize} C{Cc}@{H}]", 4.0 * B * H * W * ksize * ksize * Cc, 16.0 * B * H * W * Cc)
    _count()
    return out


def add_bcast_rows(a, p, out):
    """out[b] = a[b] + p for a [B, rows, C] contiguous, p [rows, C]."""
    assert a.is_contiguous() and p.is_contiguous() and out.is_contiguous()
    B, rows, Cc = a.shape
    check(lib.ff3d_add_bcast_rows(_ptr(a), _ptr(p), _ptr(out), B, rows, Cc, _stream()), "ff3d_add_bcast_rows")
    _count()
    return out


# --------------------------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------------------------
# camera branch
def nchw_to_nhwc(img, ld):
    """[n, C, H, W] fp32 CUDA -> [n, H, W, ld] (zero channel padding)."""
    _chk_f32(img, "nchw_to_nhwc.img")
    n, Cc, H, W = img.shape
    out = torch.empty((n, H, W, ld), dtype=torch.float32, device=img.device)
    t0 = prof.begin()
    check(lib.ff3d_nchw_to_nhwc(_ptr(img), _ptr(out), n, Cc, H, W, ld, _stream()), "ff3d_nchw_to_nhwc")
    prof.end(t0, "nchw_to_nhwc", 0.0, 4.0 * n * H * W * (Cc + ld))
    _count()
    return out


def maxpool3x3s2(x):
    _chk_f32(x, "maxpool.x")
    n, H, W, Cc = x.shape
    out = torch.empty((n, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc), dtype=torch.float32, device=x.device)
    t0 = prof.begin()
    check(lib.ff3d_maxpool3x3s2(_ptr(x), _ptr(out), n, H, W, Cc, _stream()), "ff3d_maxpool3x3s2")
    prof.end(t0, "maxpool3x3s2", 0.0, 4.0 * (x.numel() + out.numel()))
    _count()
    return out


def upsample_add(dst, src):
    """dst [n,Hd,Wd,C] += nearest-upsampled src [n,Hs,Ws,C] (in place)."""
    _chk_f32(dst, "upsample_add.dst")
    _chk_f32(src, "upsample_add.src")
    n, Hd, Wd, Cc = dst.shape
    _, Hs, Ws, _ = src.shape
    t0 = prof.begin()
    check(lib.ff3d_upsample_add(_ptr(dst), _ptr(src), n, Hd, Wd, Hs, Ws, Cc, _stream()), "ff3d_upsample_add")
    prof.end(t0, "upsample_add", 0.0, 4.0 * (2 * dst.numel() + src.numel()))
    _count()
    return dst


def lss_splat(dn, frustum, rots, trans, bev, cams, D, lo, dx):
    """dn [B*cams, fH, fW, ld] (64 context | D depth logits); bev [B, ny, nx, nz*64] is zeroed and filled."""
    _chk_f32(dn, "lss_splat.dn")
    _chk_f32(bev, "lss_splat.bev")
    n_img, fH, fW, ld = dn.shape
    B, ny, nx, cz = bev.shape
    lo3 = (C.c_float * 3)(*[float(v) for v in lo])
    dx3 = (C.c_float * 3)(*[float(v) for v in dx])
    t0 = prof.begin()
    check(lib.ff3d_lss_splat(_ptr(dn), ld, _ptr(frustum), _ptr(rots), _ptr(trans), _ptr(bev), B, cams, D, fH, fW, lo3, dx3,
                             nx, ny, cz // 64, _stream()), "ff3d_lss_splat")
    prof.end(t0, "lss_splat", 2.0 * n_img * fH * fW * D * 64, 4.0 * (dn.numel() + bev.numel() + n_img * fH * fW * D * 64))
    _count(2)
    return bev


def voxelize(points, batch_offsets, voxel_size, pc_range, max_points, max_voxels, mean_ld=8, want_voxels=False):
    """points [N, F] (samples concatenated), batch_offsets python list [B+1].
    Returns dict(coors [cap,4], num_points [cap], mean [cap, mean_ld], n_dev [1+B] (total, per sample), voxels?)."""
    _chk_f32(points, "voxelize.points")
    assert points.is_contiguous()
    N, F = points.shape
    B = len(batch_offsets) - 1
    cap = B * max_voxels
    dev = points.device
    ws_bytes = lib.ff3d_voxelize_workspace_bytes(N, B, max_voxels, max_points)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    coors = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    nump = torch.empty((cap,), dtype=torch.int32, device=dev)
    mean = torch.empty((cap, mean_ld), dtype=torch.float32, device=dev)
    n_dev = torch.empty((1 + B,), dtype=torch.int32, device=dev)
    if max_points <= 0 and want_voxels:
        raise L.Ff3dError("voxelize: dynamic mode (max_points <= 0) returns per-voxel means only")
    voxels = torch.empty((cap, max_points, F), dtype=torch.float32, device=dev) if want_voxels else None
    check(lib.ff3d_voxelize_hard(_ptr(points), N, F, L.int_array(batch_offsets), B, L.float_array(voxel_size),
                                 L.float_array(pc_range), max_points, max_voxels, _ptr(voxels), _ptr(coors),
                                 _ptr(nump), _ptr(mean), mean_ld, _ptr(n_dev), _ptr(ws), ws_bytes, _stream()),
          "ff3d_voxelize_hard")
    _count(9)
    return dict(coors=coors, num_points=nump, mean=mean, n_dev=n_dev, voxels=voxels)


def vfe_hard(vox, w, b, Cout, max_points, n_feat):
    """HardVFE: vox = dict from voxelize(want_voxels=True); returns [cap, Cout] voxel features."""
    cap = vox["coors"].shape[0]
    out = torch.empty((cap, Cout), dtype=torch.float32, device=vox["coors"].device)
    check(lib.ff3d_vfe_hard(_ptr(vox["voxels"]), _ptr(vox["num_points"]), _ptr(vox["n_dev"]), cap, max_points, n_feat,
                            _ptr(w), _ptr(b), _ptr(out), Cout, Cout, _stream()), "ff3d_vfe_hard")
    _count()
    return out


def next_pow2(v):
    p = 1
    while p < v:
        p <<= 1
    return p


class Rulebook:
    """Output-stationary rulebook of one sparse conv in mask-sorted tile order: ``nbr`` [taps, cap] (input row feeding
    tile position j through tap t, -1 = none), ``tile_mask`` [ceil(cap/128)] (OR of the tap masks of each 128-row tile:
    the gather-GEMM skips the other taps), ``y_row`` (output ROW of tile position j: strided convs run over their own
    mask order and write the new level's rows through it) or ``y_off`` (element offset into the NHWC BEV grid)."""

    def __init__(self, nbr, tile_mask, y_off=None, y_row=None):
        self.nbr, self.tile_mask, self.y_off, self.y_row = nbr, tile_mask, y_off, y_row


SUBM_K, SUBM_S, SUBM_P = (3, 3, 3), (1, 1, 1), (1, 1, 1)


class SparseLevel:
    """One resolution level of the sparse encoder: coordinates, device row count, hash (coordinate -> row)."""

    def __init__(self, coors, n_dev, cap, batch, shape):
        self.coors, self.n_dev, self.cap, self.batch, self.shape = coors, n_dev, cap, batch, tuple(shape)
        self.hsize = next_pow2(2 * cap)
        self.hkeys = torch.empty((self.hsize,), dtype=torch.int32, device=coors.device)
        self.hvals = torch.empty((self.hsize,), dtype=torch.int32, device=coors.device)
        self.subm = None

    def build_hash(self):
        D, H, W = self.shape
        check(lib.ff3d_sp_hash_build(_ptr(self.coors), _ptr(self.n_dev), self.cap, self.batch, D, H, W,
                                     _ptr(self.hkeys), _ptr(self.hvals), self.hsize, _stream()), "ff3d_sp_hash_build")
        _count(2)

    def _sorted_perm(self, out_coors, n_out, cap_out, k3, s3, p3):
        """(perm, nbr_u): perm[j] = row (of out_coors) at sorted tile position j, ordered by the tap mask of the conv
        (k3, s3, p3) whose INPUT level is self; nbr_u [kvol, cap_out] = the probed input rows in the CURRENT order."""
        D, H, W = self.shape
        dev = out_coors.device
        kvol = k3[0] * k3[1] * k3[2]
        keys = torch.empty((cap_out,), dtype=torch.int32, device=dev)
        nbr_u = torch.empty((kvol, cap_out), dtype=torch.int32, device=dev)
        check(lib.ff3d_sp_tap_keys(_ptr(out_coors), _ptr(n_out), cap_out, D, H, W, _ptr(self.hkeys), _ptr(self.hvals),
                                   self.hsize, L.int_array(k3), L.int_array(s3), L.int_array(p3), _ptr(keys), _ptr(nbr_u),
                                   _stream()), "ff3d_sp_tap_keys")
        ws_bytes = lib.ff3d_sort_workspace_bytes(cap_out)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        skeys = torch.empty_like(keys)
        perm = torch.empty((cap_out,), dtype=torch.int32, device=dev)
        check(lib.ff3d_sort_pairs(_ptr(keys), None, _ptr(n_out), cap_out, kvol, _ptr(skeys), _ptr(perm), _ptr(ws), ws_bytes,
                                  _stream()), "ff3d_sort_pairs")
        _count(1 + 3 * ((kvol + 8) // 9))
        return perm, nbr_u

    def _rulebook(self, out_coors, perm, n_out, cap_out, k3, s3, p3, y_mode=0, ldy=0, bev=(0, 0, 0), nbr_u=None, inv=None):
        """Rulebook in tile order.  With ``nbr_u`` (rows kept by the tap-key pass) the neighbour map is a re-ordering
        (ff3d_sp_nbr_permute); without, the hash is probed (ff3d_sp_nbr_build: levels that were never sorted)."""
        D, H, W = self.shape
        dev = out_coors.device
        kvol = k3[0] * k3[1] * k3[2]
        nbr = torch.empty((kvol, cap_out), dtype=torch.int32, device=dev)
        tile_mask = torch.empty(((cap_out + 127) // 128,), dtype=torch.int32, device=dev)
        y_off = torch.empty((cap_out,), dtype=torch.int32, device=dev) if y_mode else None
        if nbr_u is not None:
            check(lib.ff3d_sp_nbr_permute(_ptr(nbr_u), _ptr(perm), _ptr(inv), _ptr(out_coors), _ptr(n_out), cap_out, kvol,
                                          _ptr(nbr), _ptr(tile_mask), _ptr(y_off), y_mode, ldy, bev[0], bev[1], bev[2], _stream()),
                  "ff3d_sp_nbr_permute")
        else:
            check(lib.ff3d_sp_nbr_build(_ptr(out_coors), _ptr(perm), _ptr(n_out), cap_out, D, H, W, _ptr(self.hkeys),
                                        _ptr(self.hvals), self.hsize, L.int_array(k3), L.int_array(s3), L.int_array(p3),
                                        _ptr(nbr), _ptr(tile_mask), _ptr(y_off), y_mode, ldy, bev[0], bev[1], bev[2], _stream()),
                  "ff3d_sp_nbr_build")
        _count()
        return Rulebook(nbr, tile_mask, y_off if y_mode == 2 else None, y_off if y_mode == 1 else None)

    def sort_by_mask(self):
        """Re-store the level in SubM tap-mask order (all SubM convs of the level then run on homogeneous tiles).
        Returns perm (new row i <- old row perm[i]) so that the caller can move row data stored in the old order."""
        D, H, W = self.shape
        perm, nbr_u = self._sorted_perm(self.coors, self.n_dev, self.cap, SUBM_K, SUBM_S, SUBM_P)
        coors = torch.empty_like(self.coors)
        inv = torch.empty((self.cap,), dtype=torch.int32, device=self.coors.device)
        check(lib.ff3d_sp_level_permute(_ptr(self.coors), _ptr(perm), _ptr(self.n_dev), self.cap, D, H, W, _ptr(coors),
                                        _ptr(self.hkeys), _ptr(self.hvals), self.hsize, _ptr(inv), _stream()),
              "ff3d_sp_level_permute")
        _count()
        self._unsorted_coors = self.coors            # stays referenced: kernels of other streams may still read it
        self.coors = coors
        self.subm = None
        self._subm_src = (nbr_u, perm, inv)          # probed rows in the old order: subm_map() re-orders them
        return perm

    def subm_map(self):
        """SubM k=3 rulebook of the level (rows already in mask order: no row map)."""
        if self.subm is None:
            src = getattr(self, "_subm_src", None)
            if src is not None:
                nbr_u, perm, inv = src
                self.subm = self._rulebook(self.coors, perm, self.n_dev, self.cap, SUBM_K, SUBM_S, SUBM_P, nbr_u=nbr_u, inv=inv)
                self._subm_src = None
            else:
                self.subm = self._rulebook(self.coors, None, self.n_dev, self.cap, SUBM_K, SUBM_S, SUBM_P)
        return self.subm

    def create_down_level(self, k3, s3, p3, cap_out, overflow):
        """Site-creation half of a SparseConv3d (k3, s3, p3): the output level (coordinates, count, hash) in hash-slot order.
        Needs only this level's coordinates (in any order), not its mask sort."""
        D, H, W = self.shape
        oshape = tuple((self.shape[i] + 2 * p3[i] - k3[i]) // s3[i] + 1 for i in range(3))
        cells = self.batch * oshape[0] * oshape[1] * oshape[2]
        cap_out = int(min(cap_out, cells))
        dev = self.coors.device
        coors_o = torch.empty((cap_out, 4), dtype=torch.int32, device=dev)
        n_o = torch.empty((1,), dtype=torch.int32, device=dev)
        lvl = SparseLevel(coors_o, n_o, cap_out, self.batch, oshape)
        scratch = torch.empty((lib.ff3d_sp_down_sites_scratch_ints(lvl.hsize),), dtype=torch.int32, device=dev)
        check(lib.ff3d_sp_down_sites(_ptr(self.coors), _ptr(self.n_dev), self.cap, self.batch, D, H, W, L.int_array(k3),
                                     L.int_array(s3), L.int_array(p3), _ptr(coors_o), _ptr(n_o), cap_out, oshape[0],
                                     oshape[1], oshape[2], _ptr(lvl.hkeys), _ptr(lvl.hvals), lvl.hsize, _ptr(overflow),
                                     _ptr(scratch), _stream()), "ff3d_sp_down_sites")
        _count(5)
        return lvl

    def down_rulebook(self, lvl, k3, s3, p3, ldy, bev=None):
        """Rulebook of the SparseConv3d (k3, s3, p3) from this level into ``lvl`` (both in their FINAL row order): the conv
        runs over its own mask order and writes through ``y_off`` (row * ldy, or -- ``bev=(H, W, C)`` -- straight into the
        NHWC BEV grid)."""
        perm, nbr_u = self._sorted_perm(lvl.coors, lvl.n_dev, lvl.cap, k3, s3, p3)
        if bev is not None:
            return self._rulebook(lvl.coors, perm, lvl.n_dev, lvl.cap, k3, s3, p3, y_mode=2, ldy=ldy, bev=bev, nbr_u=nbr_u)
        return self._rulebook(lvl.coors, perm, lvl.n_dev, lvl.cap, k3, s3, p3, y_mode=1, ldy=ldy, nbr_u=nbr_u)

    def downsample(self, k3, s3, p3, cap_out, overflow, ldy, sort_level=True, bev=None):
        """SparseConv3d (k3, s3, p3): creates the output level and the conv's rulebook.  The new level is stored in ITS
        SubM mask order (``sort_level``).  Returns (level, Rulebook)."""
        lvl = self.create_down_level(k3, s3, p3, cap_out, overflow)
        if sort_level:
            lvl.sort_by_mask()
        return lvl, self.down_rulebook(lvl, k3, s3, p3, ldy, bev)


def gather_rows(src, perm, n_dev, cols):
    """dst[i, :cols] = src[perm[i], :cols] for the first *n_dev rows."""
    _chk_f32(src, "gather_rows.src")
    cap = perm.shape[0]
    dst = torch.empty((cap, src.shape[1]), dtype=torch.float32, device=src.device)
    check(lib.ff3d_sp_gather_rows(_ptr(src), src.stride(0), _ptr(perm), _ptr(n_dev), cap, _ptr(dst), dst.stride(0), cols,
                                  _stream()), "ff3d_sp_gather_rows")
    _count()
    return dst


# --------------------------------------------------------------------------------------------------------------
def hip_stage(logits, acc_mask, nms_heat, feat, cls_w, cls_b, k, nms_kernel, exempt, q0, nq_total, top_idx,
              query_feat, query_pos, query_score, query_label, logits2=None):
    B, H, W, _, ldl, _ = _nhwc_geom(logits, "hip.logits")
    ldl2 = _nhwc_geom(logits2, "hip.logits2")[4] if logits2 is not None else 0
    _, _, _, Cf, ldf, fbs = _nhwc_geom(feat, "hip.feat")
    if fbs != H * W:
        raise L.Ff3dError("hip_stage: batch-dense feature view required")
    Cc = acc_mask.shape[1]
    ws_bytes = lib.ff3d_hip_workspace_bytes(B, Cc, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
    check(lib.ff3d_hip_stage(_ptr(logits), ldl, _ptr(logits2), ldl2, _ptr(acc_mask), _ptr(nms_heat), _ptr(feat), ldf, Cf, _ptr(cls_w),
                             _ptr(cls_b), B, Cc, H, W, k, nms_kernel, exempt[0], exempt[1], q0, nq_total, _ptr(top_idx),
                             _ptr(query_feat), _ptr(query_pos), _ptr(query_score), _ptr(query_label), _ptr(ws),
                             ws_bytes, _stream()), "ff3d_hip_stage")
    _count(4)


def sine_embed(pos, w, h, dim_t, out=None):
    rows = pos.shape[0]
    out = torch.empty((rows, 256), dtype=torch.float32, device=pos.device) if out is None else out
    check(lib.ff3d_sine_embed(_ptr(pos), float(w), float(h), _ptr(dim_t), _ptr(out), rows, _stream()), "ff3d_sine_embed")
    _count()
    return out


def mha_core(q, k, v, out, B, Nq, heads, d):
    check(lib.ff3d_mha_core(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out), out.stride(0),
                            B, Nq, heads, d, _stream()), "ff3d_mha_core")
    _count()
    return out


class LevelGeom:
    def __init__(self, shapes):
        self.shapes = [(int(h), int(w)) for h, w in shapes]
        self.starts, s = [], 0
        for h, w in self.shapes:
            self.starts.append(s)
            s += h * w
        self.n_tokens = s
        self.h = L.int_array([h for h, _ in self.shapes])
        self.w = L.int_array([w for _, w in self.shapes])
        self.s = L.int_array(self.starts)
        self.L = len(self.shapes)


def msda(value, v_col0, geom, P, ref, ref_w, ref_h, offs, attw, out, B, Nq, heads, d):
    """value [B, n_tokens, ldv] contiguous; offs/attw row views of one [B*Nq, *] buffer."""
    ldv = value.stride(1)
    check(lib.ff3d_msda(_ptr(value), ldv, v_col0, value.stride(0) // ldv, geom.h, geom.w, geom.s, geom.L, P, _ptr(ref),
                        float(ref_w), float(ref_h), _ptr(offs), offs.stride(0), _ptr(attw), attw.stride(0), _ptr(out),
                        B, Nq, heads, d, _stream()), "ff3d_msda")
    _count()
    return out


def roi_sample(query_box, value, geom, Cc, g, expand, cell, origin, roi_range, out, B, Nq):
    ldv = value.stride(1)
    check(lib.ff3d_roi_sample(_ptr(query_box), query_box.stride(0), _ptr(value), ldv, value.stride(0) // ldv, geom.h,
                              geom.w, geom.s, geom.L, Cc, g, float(expand), float(cell[0]), float(cell[1]),
                              float(origin[0]), float(origin[1]), L.float_array(roi_range), _ptr(out), B, Nq, _stream()),
          "ff3d_roi_sample")
    _count()
    return out


def roi_sample_split(query_box, value, geom, Cc, g, expand, cell, origin, roi_range, B, Nq):
    """roi_sample with the [B*Nq, L*g*g*C] operand written straight in split form (fp16 [hi | lo]) for the TMA-fed
    roi_mlp.0 GEMM: same bytes as the fp32 operand, no conversion inside the GEMM."""
    ldv = value.stride(1)
    K = geom.L * g * g * Cc
    out = Split.empty((B * Nq,), K, value.device)
    check(lib.ff3d_roi_sample_split(_ptr(query_box), query_box.stride(0), _ptr(value), ldv, value.stride(0) // ldv, geom.h,
                                    geom.w, geom.s, geom.L, Cc, g, float(expand), float(cell[0]), float(cell[1]),
                                    float(origin[0]), float(origin[1]), L.float_array(roi_range), C.c_void_p(out.ptr), B, Nq,
                                    _ptr(gemm_flag(value.device)), _stream()), "ff3d_roi_sample_split")
    _count()
    return out


def add_bcast_rows_split(a, p):
    """Split of (a[b] + p) for a [B, rows, C] contiguous, p [rows, C]: [B*rows, 2C] fp16 [hi | lo]."""
    assert a.is_contiguous() and p.is_contiguous()
    B, rows, Cc = a.shape
    out = Split.empty((B * rows,), Cc, a.device)
    check(lib.ff3d_add_bcast_rows_split(_ptr(a), _ptr(p), C.c_void_p(out.ptr), B, rows, Cc, _ptr(gemm_flag(a.device)), _stream()),
          "ff3d_add_bcast_rows_split")
    _count()
    return out


def pack_frag(w, n_pad=None, k_pad=None):
    """nn.Linear-style weight [N, K] (y = x @ w.T) -> int32 [N/8, K/16, 32, 4]: the B operand of mma.sync m16n8k16 in
    FRAGMENT order with fp16 hi / lo parts (csrc/decstage.cu).  Lane (g, t) of n-tile j, k-step s holds
    b0 = w[8j + g, 16s + 2t : +2], b1 = w[8j + g, 16s + 8 + 2t : +2]; the four words are (hi b0, hi b1, lo b0, lo b1)."""
    w = w.detach().double().cpu()
    N, K = w.shape
    n_pad = n_pad or (N + 15) // 16 * 16
    k_pad = k_pad or (K + 15) // 16 * 16
    wp = torch.zeros((n_pad, k_pad), dtype=torch.float64)
    wp[:N, :K] = w
    hi, lo = split_f16(wp.float())
    parts = []
    for src in (hi, lo):
        v = src.view(n_pad // 8, 8, k_pad // 16, 2, 4, 2)               # [j, g, s, b, t, pair]
        parts.append(v.permute(0, 2, 1, 4, 3, 5))                       # [j, s, g, t, b, pair]
    out = torch.stack(parts, dim=4).contiguous()                        # [j, s, g, t, (hi, lo), b, pair]
    return out.view(n_pad // 8, k_pad // 16, 32, 4, 2).view(torch.int32).reshape(n_pad // 8, k_pad // 16, 32, 4)


def unpack_frag(frag, N, K):
    """Inverse of pack_frag (test helper): -> (hi, lo) fp32 [N, K]."""
    nj, ns = frag.shape[0], frag.shape[1]
    v = frag.contiguous().view(torch.float16).view(nj, ns, 8, 4, 2, 2, 2)      # [j, s, g, t, (hi, lo), b, pair]
    out = []
    for part in range(2):
        x = v[:, :, :, :, part].permute(0, 2, 1, 4, 3, 5).reshape(nj * 8, ns * 16)   # [j, g, s, b, t, pair]
        out.append(x.float()[:N, :K])
    return out[0], out[1]


FUSED_DECODER = _os.environ.get("FF3D_FUSED_DECODER", "1") != "0"
# sparse encoder: build the rulebooks on a side stream, concurrently with the gather-GEMMs of earlier levels
SPARSE_OVERLAP = _os.environ.get("FF3D_SPARSE_OVERLAP", "1") != "0"
# stride-2 3x3 convs on the TMA-fed kernel (tensor map with a traversal stride on W / H) when their input exists in split form
TMA_STRIDED = _os.environ.get("FF3D_TMA_STRIDED", "1") != "0"
# FF3D_SPARSE_MARKS=1: per-level stage markers inside the sparse encoder (main-stream waits for the side streams included)
SPARSE_MARKS = _os.environ.get("FF3D_SPARSE_MARKS", "0") == "1"


def decoder_stage(x, qpe, q_pos, ref_w, ref_h, value, geom, n_points, stage_w, B, nq, pred_cols):
    """One decoder stage (all its layers + prediction heads) in one launch (csrc/decstage.cu).  ``stage_w`` = dict from
    model.FocalDecoder.prepare: layers = [dict(w_qkv, b_qkv, ...)], w_h1, b_h1, n_h1, w_h2, b_h2, n_pred, ffn.
    Returns (x_out [B*nq, 128], pred [B*nq, pred_cols])."""
    _chk_f32(x, "decoder_stage.x"); _chk_f32(qpe, "decoder_stage.qpe"); _chk_f32(q_pos, "decoder_stage.q_pos")
    assert x.is_contiguous() and qpe.is_contiguous() and q_pos.is_contiguous() and value.is_contiguous()
    dev = x.device
    layers = stage_w["layers"]
    d = L.DecoderStageDesc()
    d.B, d.nq, d.n_layers, d.hidden, d.heads = B, nq, len(layers), x.shape[1], stage_w["heads"]
    d.n_levels, d.n_points, d.ffn = geom.L, n_points, stage_w["ffn"]
    for i, (h, w) in enumerate(geom.shapes):
        d.lvl_h[i], d.lvl_w[i], d.lvl_start[i] = h, w, geom.starts[i]
    d.x_in, d.qpe, d.q_pos = x.data_ptr(), qpe.data_ptr(), q_pos.data_ptr()
    d.ref_w, d.ref_h = float(ref_w), float(ref_h)
    ldv = value.stride(1)
    d.value, d.ldv, d.v_bstride = value.data_ptr(), ldv, value.stride(0) // ldv
    for i, lay in enumerate(layers):
        lw = d.layers[i]
        for k in ("w_qkv", "b_qkv", "w_o", "b_o", "w_oa", "b_oa", "w_op", "b_op", "w_f1", "b_f1", "w_f2", "b_f2"):
            setattr(lw, k, lay[k].data_ptr())
        for n in range(3):
            lw.ln_gamma[n], lw.ln_beta[n] = lay["ln"][n][0].data_ptr(), lay["ln"][n][1].data_ptr()
    d.w_h1, d.b_h1, d.n_h1 = stage_w["w_h1"].data_ptr(), stage_w["b_h1"].data_ptr(), stage_w["n_h1"]
    d.w_h2, d.b_h2, d.n_pred = stage_w["w_h2"].data_ptr(), stage_w["b_h2"].data_ptr(), stage_w["n_pred"]
    x_out = torch.empty_like(x)
    pred = torch.empty((B * nq, pred_cols), dtype=torch.float32, device=dev)
    d.x_out, d.pred, d.ld_pred, d.pred_cols = x_out.data_ptr(), pred.data_ptr(), pred.stride(0), pred_cols
    ws_bytes = lib.ff3d_decoder_stage_workspace_bytes(B, nq, len(layers))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws_bytes
    d.overflow_dev = gemm_flag(dev).data_ptr()
    t0 = prof.begin()
    check(lib.ff3d_decoder_stage(C.byref(d), _stream()), "ff3d_decoder_stage")
    M = B * nq
    ffn = stage_w["ffn"]
    flops = len(layers) * (2.0 * M * 128 * (384 + 128 + stage_w["n_oa"] + 128 + 2 * ffn) + 4.0 * B * nq * nq * 128) \
        + 2.0 * M * (128 * stage_w["n_h1"] + stage_w["n_h1"] * stage_w["n_pred"])
    prof.end(t0, f"decoder_stage[{len(layers)}L x {M}]", flops, 4.0 * M * 128 * 4)
    _count(2)
    return x_out, pred


def head_update(pred, query_pos, prev):
    check(lib.ff3d_head_update(_ptr(pred), pred.stride(0), _ptr(query_pos), _ptr(prev),
                               prev.stride(0) if prev is not None else 0, pred.shape[0], _stream()), "ff3d_head_update")
    _count()


def class_select(full, label, group_k, tail, num_classes, out_cols_ld):
    """full [rows, ld] = [nc*k0 | nc*k1 | ... | tail] -> [rows, out_cols_ld] keeping each row's own class set."""
    rows = full.shape[0]
    out = torch.zeros((rows, out_cols_ld), dtype=torch.float32, device=full.device)
    gk = (C.c_int * len(group_k))(*group_k)
    check(lib.ff3d_class_select(_ptr(full), full.stride(0), _ptr(label), gk, len(group_k), tail, num_classes, _ptr(out),
                                out.stride(0), rows, _stream()), "ff3d_class_select")
    _count()
    return out


def box_decode(pred, cls_col, has_vel, query_score, query_label, Cc, cell, origin, post_range, boxes, scores, labels,
               keep):
    check(lib.ff3d_box_decode(_ptr(pred), pred.stride(0), cls_col, 1 if has_vel else 0, _ptr(query_score),
                              _ptr(query_label), pred.shape[0], Cc, float(cell[0]), float(cell[1]), float(origin[0]),
                              float(origin[1]), L.float_array(post_range), _ptr(boxes), _ptr(scores), _ptr(labels),
                              _ptr(keep), _stream()), "ff3d_box_decode")
    _count()


# --------------------------------------------------------------------------------------------------------------
# output side
def nms_tasks(boxes, scores, labels, keep, tasks, nms_type, pre_max=0, post_max=0):
    """Per-task circle / rotated NMS of get_bboxes (focal_decoder.py:1333-1393).  boxes [B,nq,code], scores / labels /
    keep [B,nq]; tasks = [(class indices, radius), ...].  Returns keep_out [B,nq] uint8."""
    B, nq, ld = boxes.shape
    out = torch.empty((B, nq), dtype=torch.uint8, device=boxes.device)
    masks = (C.c_uint * len(tasks))(*[sum(1 << c for c in idx) for idx, _ in tasks])
    radii = L.float_array([r for _, r in tasks])
    mode = {"circle": 0, "rotate": 1}[nms_type]
    if mode == 0:
        post_max = 83                                   # [upstream] circle_nms(dets, thresh, post_max_size=83)
    check(lib.ff3d_nms_tasks(_ptr(boxes), ld, _ptr(scores), _ptr(labels), _ptr(keep), B, nq, len(tasks), masks, radii, mode,
                             int(pre_max or 0), int(post_max or 0), _ptr(out), _stream()), "ff3d_nms_tasks")
    _count(2)
    return out


def boxes_iou_bev(a, b):
    """Rotated BEV IoU of (x, y, z, dx, dy, dz, yaw, ...) rows: [n, m]."""
    n, m = a.shape[0], b.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=a.device)
    check(lib.ff3d_boxes_iou_bev(_ptr(a), a.stride(0), n, _ptr(b), b.stride(0), m, _ptr(out), _stream()), "ff3d_boxes_iou_bev")
    _count()
    return out


def merge_aug_bboxes_3d(aug_results, img_metas, nms_thr=0.1, max_num=500, vote_iou_thresh=0.65, dev="cuda"):
    """Test-time-augmentation merge (core/post_processing/merge_augs.py:14-184, the live branch): map every augmented
    result back (flips, scale), per-class rotated NMS (thr 0.1), IoU-weighted box voting (>= 0.65), best max_num by score.
    aug_results: list of dict(boxes_3d [n, 7|9], scores_3d [n], labels_3d [n]); img_metas: list of dict(pcd_scale_factor,
    pcd_horizontal_flip, pcd_vertical_flip).  Returns the merged dict (device tensors)."""
    boxes = []
    for r, m in zip(aug_results, img_metas):
        b = r["boxes_3d"].to(dev, torch.float32).contiguous().clone()
        check(lib.ff3d_boxes_map_back(_ptr(b), b.stride(0), b.shape[1], b.shape[0], float(m.get("pcd_scale_factor", 1.0)),
                                      int(bool(m.get("pcd_horizontal_flip", False))), int(bool(m.get("pcd_vertical_flip", False))),
                                      _stream()), "ff3d_boxes_map_back")
        _count()
        boxes.append(b)
    boxes = torch.cat(boxes)
    scores = torch.cat([r["scores_3d"].to(dev, torch.float32) for r in aug_results]).contiguous()
    labels = torch.cat([r["labels_3d"].to(dev) for r in aug_results]).int().contiguous()
    n, dim = boxes.shape
    if n == 0:
        return dict(boxes_3d=boxes, scores_3d=scores, labels_3d=labels)
    if n > 1024:
        raise L.Ff3dError(f"merge_aug_bboxes_3d: {n} boxes (<= 1024 supported)")
    n_cls = int(labels.max().item()) + 1                      # merge_augs.py:136 (a host decision in the reference too)
    keep_in = torch.ones((1, n), dtype=torch.uint8, device=dev)
    keep = nms_tasks(boxes.view(1, n, dim), scores.view(1, n), labels.view(1, n), keep_in,
                     [([c], nms_thr) for c in range(min(n_cls, 8))], "rotate")
    if n_cls > 8:
        raise L.Ff3dError("merge_aug_bboxes_3d: more than 8 classes")
    sel = keep[0].bool()
    sb, ss, sl = boxes[sel].contiguous(), scores[sel], labels[sel]
    iou = boxes_iou_bev(sb, boxes)
    iou = iou * (sl[:, None] == labels[None, :]).float()      # voting runs per class (merge_augs.py:137-148)
    voted = torch.empty((sb.shape[0], dim), dtype=torch.float32, device=dev)
    check(lib.ff3d_box_voting(_ptr(iou), sb.shape[0], n, _ptr(boxes), boxes.stride(0), dim, float(vote_iou_thresh), _ptr(voted),
                              _stream()), "ff3d_box_voting")
    _count()
    # per class the reference keeps NMS order (score descending); then a global score sort (merge_augs.py:176-178)
    order = ss.argsort(descending=True, stable=True)[:min(max_num, n)]
    return dict(boxes_3d=voted[order], scores_3d=ss[order], labels_3d=sl[order])
