"""EXPERIMENTAL host side of the fp16 hi/lo GEMM operand format (round-2 work item 1).
Not used by the product path; pairs with focalformer3d_b200/csrc/experimental/tcgemm_f16.cu (compile-checked only).

Weight split:  w = hi + 2^-11 * lo,  hi = rn_f16(w),  lo = rn_f16((w - hi) * 2^11)   (|w| << 65504 for any trained layer).
Image layout per (N tile, pipeline K step of 64): [hi | lo] x [bn rows (output channel)] x [64 halves (K)], K-major rows
of 128 bytes with the 128-byte swizzle tcgen05 expects (16-byte chunk j of row n stored at chunk j ^ (n % 8)) -- one
cp.async.bulk per stage, exactly like the TF32 images of ops.tc_weight_images."""
import torch

LO_SCALE = 2048.0


def f16_ntile(cin, cout):
    if not (cin in (8, 16, 32) or (cin >= 64 and cin % 64 == 0)):
        return 0
    if cout % 128 == 0:
        return 128
    return cout if cout in (64, 32, 16) else 0


def f16_stages(cin, taps):
    if cin >= 64:
        return taps * (cin // 64)
    tps = 64 // cin
    return (taps + tps - 1) // tps


def split_f16(t):
    """fp32 tensor -> (hi, lo) fp16 tensors with t ~= hi + lo / 2048."""
    hi = t.clamp(-65504.0, 65504.0).half()
    lo = ((t - hi.float()) * LO_SCALE).clamp(-65504.0, 65504.0).half()
    return hi, lo


def tc_weight_images_f16(w, bn=None):
    """[taps, cin, cout] fp32 (CPU) -> ([n_tiles, n_stages, 2, bn, 64] fp16, bn)."""
    taps, cin, cout = w.shape
    default_bn = f16_ntile(cin, cout)
    if default_bn <= 0:
        return None, 0
    bn = bn or default_bn
    assert cout % bn == 0 and bn in (16, 32, 64, 128)
    n_stages = f16_stages(cin, taps)
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
        if cin >= 64:
            parts.append(kmat.reshape(taps, cin, cout))
        else:
            parts.append(kmat.reshape(-1, cin, cout)[:taps])
    return parts[0], parts[1]


# ------------------------------------------------------------------------------------------------------------------
# Opt-in test harness (round 2): route every tensor-core GEMM of focalformer3d_b200.ops through the experimental fp16
# kernel, so that the WHOLE existing GPU test suite (kernel parity + end to end) validates it:
#     make -C focalformer3d_b200/csrc experimental && FF3D_EXPERIMENTAL_F16=1 python -m pytest tests -m gpu
# Never active by default; nothing in the product path imports this.
def enable(ops_module=None):
    import ctypes as C
    import os
    from . import lib as L
    ops = ops_module
    if ops is None:
        from . import ops
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libff3d_x.so")
    if not os.path.exists(path):
        raise RuntimeError("libff3d_x.so missing: make -C focalformer3d_b200/csrc experimental")
    xlib = C.CDLL(path)
    xlib.ff3d_x_tcgemm_f16.restype = C.c_int
    xlib.ff3d_x_tcgemm_f16.argtypes = [C.POINTER(L.GemmDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    xlib.ff3d_last_error.restype = C.c_char_p
    state = {"overflow": None, "calls": 0, "fallbacks": 0}
    orig = ops._gemm

    def gemm_f16(d, w, what):
        taps, cin, cout = w.w.shape
        if f16_ntile(cin, cout) <= 0 or w.img is None:
            state["fallbacks"] += 1
            return orig(d, w, what)
        if getattr(w, "img16", None) is None:
            bn = w.bn if (w.bn and cout % w.bn == 0) else None
            img, w.bn16 = tc_weight_images_f16(w.w.detach().float().cpu(), bn)
            w.img16 = img.to(w.w.device)
        if state["overflow"] is None:
            state["overflow"] = torch.zeros(1, dtype=torch.int32, device=w.w.device)
        rc = xlib.ff3d_x_tcgemm_f16(C.byref(d), C.c_void_p(w.img16.data_ptr()), w.bn16,
                                    C.c_void_p(state["overflow"].data_ptr()), ops._stream())
        if rc != 0:
            raise L.Ff3dError(f"ff3d_x_tcgemm_f16({what}): {xlib.ff3d_last_error().decode()}")
        state["calls"] += 1

    ops._gemm = gemm_f16
    return state
